"""Scratch probe (GPU box): ANS1 per 4 MiB block against the reference on the config-4 data."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "kanzi-cpp_b200")):
    sys.path.insert(0, p)
import synth
from kanzi_b200 import Context
from oracle.oracle import Ref
ref = Ref.load()
ctx = Context(0, 4 << 20, 4)
data = synth.synth_compressible(256 << 20, 4)
bs = 4 << 20
for k in range(0, 64):
    blk = data[k * bs:(k + 1) * bs]
    want, wbits = ref.entropy_encode("ANS1", blk)
    try:
        got, bits = ctx.entropy_encode("ANS1", blk)
        same = bits == wbits and np.array_equal(got, want)
    except Exception as e:
        same = f"enc exc {e}"
    try:
        dec = ctx.entropy_decode("ANS1", want, wbits, blk.size)
        dok = bool(np.array_equal(dec, blk))
        if not dok:
            d = np.nonzero(dec != blk)[0]
            dok = f"first wrong byte {int(d[0])} of {d.size}"
    except Exception as e:
        dok = f"dec exc {e}"
    if same is not True or dok is not True:
        print("block", k, "enc same", same, "dec", dok, "bits", wbits, flush=True)
print("done")
