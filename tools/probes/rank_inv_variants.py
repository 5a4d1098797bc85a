"""Probe: device time of the inverse RANK stage for the kernel variant selected by KNZ_SBRT_INV
(1 = default, 2 = rank-0 steps branched over the shuffles).  Usage: python tools/probes/rank_inv_variants.py [blocks]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kanzi-cpp_b200"))
import synth  # noqa: E402
from kanzi_b200 import Context  # noqa: E402

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bs = 4 << 20
data = synth.synth_compressible(blocks * bs, 2)
ctx = Context(0, bs, blocks)
comp = ctx.compress(data, "BWT+RANK+ZRLT", "ANS0", bs)
enc = ctx.timings()
res = []
for it in range(3):
    back = ctx.decompress(comp, data.size)
    t = ctx.timings()
    res.append((round(t["rank"], 2), round(t["total"], 2)))
assert np.array_equal(back, data)
print(f"KNZ_SBRT_INV={os.environ.get('KNZ_SBRT_INV', '1')} blocks={blocks} encode_total={enc['total']:.1f} "
      f"decode (rank_ms, total_ms) x3 = {res}", flush=True)
