"""Profiling target for ncu (not a test): one warm-up + one measured encode/decode of
64 x 4 MiB blocks of the headline workload through the device-resident block API."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np, torch
import synth
from kanzi_b200 import Context, E_IDS, sharded

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
BLOCK = 4 << 20
data = synth.synth_compressible(nb * BLOCK, 2)
ctx = Context(0, BLOCK, nb)
dev = torch.device("cuda", 0)
d_in = torch.from_numpy(data).view(nb, BLOCK).to(dev)
ostride = (BLOCK + BLOCK // 4 + 4096 + 255) // 256 * 256
d_blk = torch.zeros((nb, ostride), dtype=torch.uint8, device=dev)
d_bits = torch.zeros(nb, dtype=torch.int64, device=dev)
d_dec = torch.empty((nb, BLOCK), dtype=torch.uint8, device=dev)
tt, et = ctx.transform_type("BWT+RANK+ZRLT"), E_IDS["ANS0"]
for _ in range(reps):
    sharded.encode_shard(ctx, tt, et, BLOCK, d_in, [BLOCK] * nb, BLOCK, d_blk, d_bits)
    e = ctx.timings()
    sharded.decode_shard(ctx, tt, et, BLOCK, d_blk, d_bits.cpu().numpy().astype(np.uint64), d_dec)
    d = ctx.timings()
torch.cuda.synchronize()
assert torch.equal(d_dec, d_in)
print("enc", e)
print("dec", d)
