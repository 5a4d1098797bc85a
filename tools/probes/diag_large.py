"""Scratch probe (GPU box): which stage of BWT+RANK+ZRLT / ANS0 diverges from the reference at 32 MiB blocks."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "kanzi-cpp_b200")):
    sys.path.insert(0, p)
import synth
from kanzi_b200 import Context
from oracle.oracle import Ref

ref = Ref.load()
ctx = Context(0, 32 << 20, 4)
for n in (24 << 20, (32 << 20) - 16, 32 << 20):
    data = synth.synth_compressible(n, 82)
    cur = data
    for t in ("BWT", "RANK", "ZRLT"):
        want, fl, ok = ref.sequence_forward(t, cur, out_cap=cur.size + 1024)
        got, applied = ctx.transform_forward(t, cur, cur.size + 1024)
        same = got.size == want.size and np.array_equal(got, want)
        d = np.nonzero(got[:min(got.size, want.size)] != want[:min(got.size, want.size)])[0]
        print(n, t, "applied", applied, "ref flags", fl, "sizes", got.size, want.size, "same", same,
              "first diff", int(d[0]) if d.size else None, "ndiff", int(d.size), flush=True)
        if applied:
            back, ok2 = ctx.transform_inverse(t, want, cur.size + 64)
            print("   inverse ok", ok2, bool(back.size == cur.size and np.array_equal(back, cur)), flush=True)
        cur = want
    enc, bits = ctx.entropy_encode("ANS0", cur)
    wenc, wbits = ref.entropy_encode("ANS0", cur)
    print(n, "ANS0 bits", bits, wbits, "same", bool(bits == wbits and np.array_equal(enc, wenc)), flush=True)
