"""world_size-2 (and 3) gloo tests of the N>1 path on CPU: every rank drives the emulator build of
the library through the multi-GPU C ABI (knz_compress_dist / knz_decompress_dist /
knz_dist_encode_dev / knz_dist_decode_dev, csrc/dist.cu) with torch.distributed(gloo) serving the
three collectives; rank 0 compares the assembled stream with the oracle, every rank checks the
blocks it decoded."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "sim", "libknzsim.so")

WORKER = r'''
import os, sys, ctypes
import numpy as np, torch, torch.distributed as dist
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200"), os.path.join(ROOT, "tests")]
import synth
from kanzi_b200 import Context, E_IDS, _ptr
from oracle.oracle import Oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
BS = 65536
ctx = Context(0, BS, 2, lib_path=sys.argv[2])   # batch 2 < blocks per rank: exercises sub-batching too
ctx.dist_init(rank, world)
oracle = Oracle()
for nbytes, tname, ename in ((7 * BS + 1234, "BWT+RANK+ZRLT", "ANS0"), (5 * BS + 9, "NONE", "HUFFMAN"),
                             (2 * BS, "BWT+SRT+ZRLT", "FPAQ"), (BS // 2, "ZRLT", "ANS1")):
    data = synth.synth_compressible(nbytes, 41)
    # ---- stream level, host buffers
    comp = ctx.compress_dist(data, tname, ename, BS)
    want = oracle.stream_compress(data, tname, ename, BS)
    if rank == 0:
        assert comp.size == want.size and np.array_equal(comp, want), (tname, ename, comp.size, want.size)
    else:
        assert comp.size == 0
    out = np.full(nbytes, 0xEE, dtype=np.uint8)
    back = ctx.decompress_dist(want, nbytes, out=out)
    assert back.size == nbytes
    nblocks = (nbytes + BS - 1) // BS
    for i in range(nblocks):
        lo, hi = i * BS, min((i + 1) * BS, nbytes)
        if i % world == rank:
            assert np.array_equal(out[lo:hi], data[lo:hi]), (tname, ename, "own block", i)
        else:
            assert (out[lo:hi] == 0xEE).all(), (tname, ename, "foreign block touched", i)
    # ---- device-resident legs (the emulator's device memory is host memory)
    my = list(range(rank, nblocks, world))
    lens = np.array([min(BS, nbytes - i * BS) for i in my], dtype=np.int32)
    d_in = np.zeros((max(len(my), 1), BS), dtype=np.uint8)
    for k, i in enumerate(my):
        d_in[k, : lens[k]] = data[i * BS: i * BS + lens[k]]
    tt, et = ctx.transform_type(tname), E_IDS[ename]
    hdr = np.zeros(32, dtype=np.uint8)
    hb = ctx.lib.knz_stream_header(ctypes.c_uint64(tt), et, BS, ctypes.c_int64(nbytes), _ptr(hdr))
    stream = np.zeros(want.size + 4096, dtype=np.uint8)
    bits = np.zeros(nblocks, dtype=np.uint64)
    end = ctypes.c_uint64(0)
    rc = ctx.lib.knz_dist_encode_dev(ctx.h, tt, et, BS, _ptr(d_in), BS, _ptr(lens), len(my), nblocks, min(BS, nbytes),
                                     _ptr(stream), stream.size, 8 * hb, _ptr(bits), ctypes.byref(end))
    assert rc == 0, (rc, ctx.lib.knz_last_error(ctx.h))
    if rank == 0:
        total = (end.value + 8 + 7) // 8
        got = stream[:total].copy()
        got[:hb] = hdr[:hb]
        assert got.size == want.size and np.array_equal(got, want), (tname, ename, "dev stream")
    d_out = np.zeros((max(len(my), 1), BS), dtype=np.uint8)
    ol = np.zeros(max(len(my), 1), dtype=np.int32)
    rc = ctx.lib.knz_dist_decode_dev(ctx.h, tt, et, BS, _ptr(stream), want.size, 8 * hb, _ptr(bits), nblocks, _ptr(d_out),
                                     BS, _ptr(ol))
    assert rc == 0, (rc, ctx.lib.knz_last_error(ctx.h))
    for k, i in enumerate(my):
        assert ol[k] == lens[k] and np.array_equal(d_out[k, : lens[k]], d_in[k, : lens[k]]), (tname, ename, "dev block", i)
# ---- block checksums (XXHash32 / XXHash64) through the sharded path
for ck in (32, 64):
    nbytes = 5 * BS + 77
    data = synth.synth_compressible(nbytes, 43)
    ctx.set_checksum(ck)
    comp = ctx.compress_dist(data, "BWT+RANK+ZRLT", "ANS0", BS)
    want = oracle.stream_compress(data, "BWT+RANK+ZRLT", "ANS0", BS, checksum=ck)
    if rank == 0:
        assert comp.size == want.size and np.array_equal(comp, want), ("checksum", ck)
    ctx.set_checksum(0)  # the stream header says which checksum to verify
    out = np.zeros(nbytes, dtype=np.uint8)
    ctx.decompress_dist(want, nbytes, out=out)
    for i in range(rank, (nbytes + BS - 1) // BS, world):
        assert np.array_equal(out[i * BS: (i + 1) * BS], data[i * BS: (i + 1) * BS]), ("checksum", ck, i)
# ---- sequences with leading host stages (levels 3 and 5 of the reference): pinned against the reference itself
from oracle.oracle import Ref
ref = Ref.load()
if ref is not None:
    from test_pre_stages import mixed_stream
    data = mixed_stream(BS)
    nbytes = data.size
    for tname, ename, ck in (("TEXT+UTF+PACK+MM+LZX", "HUFFMAN", 0), ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0", 32)):
        ctx.set_checksum(ck)
        comp = ctx.compress_dist(data, tname, ename, BS)
        ctx.set_checksum(0)
        want = ref.stream_compress(data, tname, ename, BS, 1, ck)
        if rank == 0:
            assert comp.size == want.size and np.array_equal(comp, want), ("host stages", tname, ename)
        out = np.full(nbytes, 0xEE, dtype=np.uint8)
        ctx.decompress_dist(want, nbytes, out=out)
        for i in range((nbytes + BS - 1) // BS):
            lo, hi = i * BS, min((i + 1) * BS, nbytes)
            if i % world == rank:
                assert np.array_equal(out[lo:hi], data[lo:hi]), ("host stages", tname, "own block", i)
            else:
                assert (out[lo:hi] == 0xEE).all(), ("host stages", tname, "foreign block touched", i)
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_multi_rank_c_abi(tmp_path, world):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim"), "-j8"], stdout=subprocess.DEVNULL)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29590 + world)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", port, str(script), ROOT, SIM],
                         env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "GLOO_OK" in out.stdout
