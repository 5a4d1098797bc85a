"""Experiment (not a test): RANK-inverse cycles per step by rank class, one 4 MiB block."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np
from kanzi_b200 import Context

N = 4 << 20
rng = np.random.default_rng(5)
def mix(pz, lo, hi):
    a = rng.integers(lo, hi + 1, N).astype(np.uint8)
    a[rng.random(N) < pz] = 0
    return a
cases = {
    "all zero": np.zeros(N, np.uint8),
    "all r=1": np.ones(N, np.uint8),
    "r in 1..5": mix(0, 1, 5),
    "r in 0..5": mix(0, 0, 5),
    "r in 6..31": mix(0, 6, 31),
    "r in 1..31": mix(0, 1, 31),
    "r in 32..255": mix(0, 32, 255),
    "r in 0..255": mix(0, 0, 255),
    "zero word runs + 1..7": None,
}
a = mix(0, 1, 7).reshape(-1, 4)
runs = (np.arange(a.shape[0]) // 16) % 2 == 0
a[runs] = 0
cases["zero word runs + 1..7"] = a.reshape(-1)
ctx = Context(0, N, 1)
mhz = float(os.environ.get("SM_MHZ", "1965"))
for name, d in cases.items():
    for _ in range(2):
        out, ok = ctx.transform_inverse("RANK", d, N)
    ms = ctx.timings()["rank"]
    print("%-24s %8.2f ms  %6.1f cycles/step" % (name, ms, ms * 1e-3 * mhz * 1e6 / N))
