"""Timing probe (not a test): entropy kernels alone on the raw workload (BASELINE config 4
shape: -t NONE -e {ANS0,HUFFMAN}, 4 MiB blocks) and on the post-transform bytes of config 2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np, torch
import synth
from kanzi_b200 import Context, E_IDS, sharded

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["ANS0"]
BLOCK = 4 << 20
data = synth.synth_compressible(nb * BLOCK, 4)
ctx = Context(0, BLOCK, nb)
dev = torch.device("cuda", 0)
d_in = torch.from_numpy(data).view(nb, BLOCK).to(dev)
ostride = (BLOCK + BLOCK // 4 + 4096 + 255) // 256 * 256
d_blk = torch.zeros((nb, ostride), dtype=torch.uint8, device=dev)
d_bits = torch.zeros(nb, dtype=torch.int64, device=dev)
d_dec = torch.empty((nb, BLOCK), dtype=torch.uint8, device=dev)
for tname in ("NONE",):
    tt = ctx.transform_type(tname)
    for ename in names:
        et = E_IDS[ename]
        best_e, best_d = 1e9, 1e9
        for _ in range(4):
            sharded.encode_shard(ctx, tt, et, BLOCK, d_in, [BLOCK] * nb, BLOCK, d_blk, d_bits)
            e = ctx.timings()
            d = {"ans_dec_kernel": 0.0, "entropy": 0.0}
            if not os.environ.get('KNZ_ANS_STOP'):
                sharded.decode_shard(ctx, tt, et, BLOCK, d_blk, d_bits.cpu().numpy().astype(np.uint64), d_dec)
                d = ctx.timings()
            best_e = min(best_e, e["ans_enc_kernel"]); best_d = min(best_d, d["ans_dec_kernel"])
        torch.cuda.synchronize()
        if not os.environ.get('KNZ_ANS_STOP'):
            assert torch.equal(d_dec, d_in)
        ebytes = int((d_bits.sum().item() + 7) // 8)
        alg = nb * BLOCK + ebytes
        print(f"{tname}/{ename}: m={nb*BLOCK} e={ebytes} enc_kernel {best_e:.3f} ms = {alg/best_e/1e6:.1f} GB/s | "
              f"dec_kernel {best_d:.3f} ms = {alg/max(best_d,1e-9)/1e6:.1f} GB/s | stage enc {e['entropy']:.3f} dec {d['entropy']:.3f}")
