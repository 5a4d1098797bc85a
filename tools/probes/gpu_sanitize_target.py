"""compute-sanitizer target (not a test): small round trips through every kernel family."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np
import synth
from kanzi_b200 import Context, KanziGpuError

what = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = Context(0, 1 << 18, 4)
PIPES = {
    "round1": (("BWT+RANK+ZRLT", "ANS0"), ("BWT+MTFT+ZRLT", "HUFFMAN"), ("NONE", "ANS0"), ("ZRLT", "HUFFMAN")),
    "round2": (("NONE", "ANS1"), ("BWT+SRT+ZRLT", "FPAQ"), ("LZ", "HUFFMAN"), ("LZX", "ANS0"), ("LZP", "NONE"),
               ("LZP+LZX", "ANS1")),
}
pipes = sum(PIPES.values(), ()) if what == "all" else PIPES[what]
for name, data in (("comp", synth.synth_compressible(700000, 3)), ("inc", synth.synth_incompressible(300001, 4)),
                   ("zeros", np.zeros(270000, dtype=np.uint8))):
    for tr, en in pipes:
        for bs in (65536, 1 << 18):
            for ck in ((0, 32, 64) if tr == "LZ" or en == "ANS0" else (0,)):
                ctx.set_checksum(ck)
                c = ctx.compress(data, tr, en, bs)
                ctx.set_checksum(0)
                d = ctx.decompress(c, data.size)
                assert np.array_equal(d, data), (name, tr, en, bs, ck)
                if name == "comp" and bs == 65536:  # a few damaged streams: the decoders must stay in bounds
                    for pos in (c.size // 3, c.size // 2, c.size - 50):
                        bad = c.copy()
                        bad[pos] ^= 0x21
                        try:
                            ctx.decompress(bad, data.size)
                        except KanziGpuError:
                            pass
print("sanitize target ok")
