"""compute-sanitizer target (not a test): small round trips through every kernel family."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np
import synth
from kanzi_b200 import Context

ctx = Context(0, 1 << 18, 4)
for name, data in (("comp", synth.synth_compressible(700000, 3)), ("inc", synth.synth_incompressible(300001, 4)),
                   ("zeros", np.zeros(270000, dtype=np.uint8))):
    for tr, en in (("BWT+RANK+ZRLT", "ANS0"), ("BWT+MTFT+ZRLT", "HUFFMAN"), ("NONE", "ANS0"), ("ZRLT", "HUFFMAN")):
        for bs in (65536, 1 << 18):
            c = ctx.compress(data, tr, en, bs)
            d = ctx.decompress(c, data.size)
            assert np.array_equal(d, data), (name, tr, en, bs)
print("sanitize target ok")
