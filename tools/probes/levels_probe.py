"""Probe: the reference's levels that start with host stages, end to end through knz_compress / knz_decompress
(host buffers), next to the unmodified reference with all host threads on the same box.
Usage: python tools/probes/levels_probe.py [MiB]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kanzi-cpp_b200"))
import synth  # noqa: E402
from kanzi_b200 import Context  # noqa: E402
from oracle.oracle import Ref  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
bs = 4 << 20
data = synth.synth_silesia(mib << 20, 3)
ref = Ref.load()
jobs = min(64, os.cpu_count() or 1)
ctx = Context(0, bs, 64)
for level, tname, ename in ((2, "DNA+LZ", "HUFFMAN"), (3, "TEXT+UTF+PACK+MM+LZX", "HUFFMAN"), (5, "TEXT+UTF+BWT+RANK+ZRLT", "ANS0")):
    ctx.compress(data[: 8 * bs], tname, ename, bs)  # warm-up (lazy allocations)
    t0 = time.time()
    comp = ctx.compress(data, tname, ename, bs)
    t1 = time.time()
    enc_t = ctx.timings()
    back = ctx.decompress(comp, data.size)
    t2 = time.time()
    dec_t = ctx.timings()
    assert np.array_equal(back, data)
    line = (f"level {level} ({tname} / {ename}) {mib} MiB: {comp.size} bytes; GPU path encode {t1 - t0:.3f} s (device "
            f"{enc_t['total']:.0f} ms), decode {t2 - t1:.3f} s (device {dec_t['total']:.0f} ms)")
    if ref is not None:
        t3 = time.time()
        want = ref.stream_compress(data, tname, ename, bs, jobs, 0)
        t4 = time.time()
        rb, rc = ref.stream_decompress(want, data.size, jobs)
        t5 = time.time()
        same = want.size == comp.size and np.array_equal(want, comp)
        line += f"; reference jobs={jobs}: encode {t4 - t3:.3f} s, decode {t5 - t4:.3f} s; streams identical: {same}"
    print(line, flush=True)
# skipBlocks on data that is half incompressible
mixed = data.copy()
mixed[: (mib << 19)] = synth.synth_incompressible(mib << 19, 5)
ctx.set_skip_blocks(True)
t0 = time.time()
comp = ctx.compress(mixed, "BWT+RANK+ZRLT", "ANS0", bs)
t1 = time.time()
ctx.set_skip_blocks(False)
print(f"skipBlocks, {mib} MiB half incompressible: {comp.size} bytes, encode {t1 - t0:.3f} s", flush=True)
