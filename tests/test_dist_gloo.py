"""world_size-2 gloo test of the N>1 path's host logic (sharding, payload gather,
stream-order reassembly) on CPU: each rank drives the emulator build of the library
on its round-robin shard, rank 0 assembles the stream and compares it with the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "sim", "libknzsim.so")

WORKER = r'''
import os, sys, ctypes
import numpy as np, torch, torch.distributed as dist
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200"), os.path.join(ROOT, "tests")]
import synth
from kanzi_b200 import Context, E_IDS
from kanzi_b200 import sharded
from oracle.oracle import Oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
BS = 65536
nblocks = 6
data = synth.synth_compressible(nblocks * BS, 41)
my = sharded.shard_blocks(nblocks, rank, world)
ctx = Context(0, BS, 2, lib_path=sys.argv[2])   # batch 2 < blocks per rank: exercises batching too
ostride = (BS + BS // 4 + 4096 + 255) // 256 * 256
d_in = torch.from_numpy(data).view(nblocks, BS)[my].contiguous()
d_blk = torch.zeros((len(my), ostride), dtype=torch.uint8)
d_bits = torch.zeros(len(my), dtype=torch.int64)
tt, et = ctx.transform_type("BWT+RANK+ZRLT"), E_IDS["ANS0"]
sharded.encode_shard(ctx, tt, et, BS, d_in, [BS] * len(my), BS, d_blk, d_bits)
res = sharded.gather_blocks(d_blk, d_bits, rank, world)
d_out = torch.zeros((len(my), BS), dtype=torch.uint8)
lens = sharded.decode_shard(ctx, tt, et, BS, d_blk, d_bits.numpy().astype(np.uint64), d_out)
assert (lens == BS).all() and torch.equal(d_out, d_in), "shard round trip failed"
if rank == 0:
    blk, bits = res
    hdr = np.zeros(32, dtype=np.uint8)
    hb = ctx.lib.knz_stream_header(ctypes.c_uint64(tt), et, BS, ctypes.c_int64(data.size), hdr.ctypes.data_as(ctypes.c_void_p))
    stream = torch.zeros(data.size + data.size // 4 + 65536, dtype=torch.uint8)
    end = sharded.assemble_stream(ctx, blk, bits, stream, 8 * hb)
    total = (end + 8 + 7) // 8
    got = stream[:total].numpy().copy()
    got[:hb] = hdr[:hb]
    want = Oracle().stream_compress(data, "BWT+RANK+ZRLT", "ANS0", BS)
    assert got.size == want.size and np.array_equal(got, want), (got.size, want.size)
    print("GLOO_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_stream_assembly(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim"), "-j8"], stdout=subprocess.DEVNULL)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29591")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29591", str(script), ROOT, SIM],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "GLOO_OK" in out.stdout
