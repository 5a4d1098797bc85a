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
    # round 2, second half: leading host stages (device stages start from the state they leave), skipBlocks
    "round2b": (("TEXT+UTF+PACK+MM+LZX", "HUFFMAN"), ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0"), ("DNA+LZ", "HUFFMAN"), ("MM", "ANS1")),
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
if what in ("all", "round2b"):
    mixed = np.concatenate([synth.synth_silesia(600000, 3), synth.synth_incompressible(200000, 5), synth.synth_text(100000, 6)])
    for tr, en in (("BWT+RANK+ZRLT", "ANS0"), ("LZ", "HUFFMAN"), ("NONE", "NONE")):
        for ck in (0, 32, 64):
            ctx.set_checksum(ck)
            ctx.set_skip_blocks(True)
            c = ctx.compress(mixed, tr, en, 65536)
            ctx.set_skip_blocks(False)
            ctx.set_checksum(0)
            assert np.array_equal(ctx.decompress(c, mixed.size), mixed), ("skipBlocks", tr, en, ck)
            assert np.array_equal(ctx.decompress_range(c, 3, 9, mixed.size), mixed[2 * 65536: 8 * 65536]), ("range", tr, en, ck)
    for tr, en in PIPES["round2b"]:
        c = ctx.compress(mixed, tr, en, 65536)
        assert np.array_equal(ctx.decompress(c, mixed.size), mixed), ("levels", tr, en)
print("sanitize target ok")
