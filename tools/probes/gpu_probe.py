"""Ad-hoc GPU probe (not a test): stage timings of the headline pipeline."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np
import synth
from kanzi_b200 import Context

size = int(sys.argv[1]) if len(sys.argv) > 1 else (256 << 20)
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
data = synth.synth_compressible(size, 2)
ctx = Context(0, 4 << 20, batch)
res = {}
for it in range(3):
    t = time.time(); comp = ctx.compress(data, "BWT+RANK+ZRLT", "ANS0", 4 << 20); te = time.time() - t
    tim_e = ctx.timings()
    t = time.time(); back = ctx.decompress(comp, size); td = time.time() - t
    tim_d = ctx.timings()
    ok = bool(np.array_equal(back, data))
    res = {"size": size, "comp": int(comp.size), "enc_s": te, "dec_s": td, "enc_MBps": size / te / 1e6,
           "dec_MBps": size / td / 1e6, "enc_stage_ms": tim_e, "dec_stage_ms": tim_d, "roundtrip_ok": ok,
           "launches": ctx.launches}
    print(json.dumps(res), flush=True)
t = time.time(); c2 = ctx.compress(data, "NONE", "ANS0", 4 << 20); te = time.time() - t
print(json.dumps({"none_ans0_enc_s": te, "comp": int(c2.size), "stage_ms": ctx.timings()}))
t = time.time(); b2 = ctx.decompress(c2, size); td = time.time() - t
print(json.dumps({"none_ans0_dec_s": td, "ok": bool(np.array_equal(b2, data)), "stage_ms": ctx.timings()}))
