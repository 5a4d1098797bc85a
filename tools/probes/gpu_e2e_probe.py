"""Experiment (not a test): wall-clock of knz_compress / knz_decompress on pinned host buffers."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np, torch
import synth
from kanzi_b200 import Context

BLOCK = 4 << 20
size = 1 << 30
data = synth.synth_compressible(size, 2)
ctx = Context(0, BLOCK, 256)
host = torch.from_numpy(data).pin_memory().numpy()
oc = torch.empty(size + size // 4 + (1 << 20), dtype=torch.uint8).pin_memory().numpy()
op = torch.empty(size, dtype=torch.uint8).pin_memory().numpy()
comp = ctx.compress(host, "BWT+RANK+ZRLT", "ANS0", BLOCK, out=oc)
back = ctx.decompress(comp, size, out=op)
assert np.array_equal(back, data)
for _ in range(3):
    t0 = time.perf_counter(); comp = ctx.compress(host, "BWT+RANK+ZRLT", "ANS0", BLOCK, out=oc); t1 = time.perf_counter()
    te = ctx.timings()
    back = ctx.decompress(comp, size, out=op); t2 = time.perf_counter()
    td = ctx.timings()
    print("compress %.1f ms (stages %.1f)  decompress %.1f ms (bwt %.1f rank %.1f ent %.1f total %.1f)" % (
        (t1 - t0) * 1e3, te["total"], (t2 - t1) * 1e3, td["bwt"], td["rank"], td["entropy"], td["total"]))
