"""Scratch probe (GPU box): end-to-end knz_compress / knz_decompress of the headline workload under different
encode sub-batch schedules (KNZ_ENC_BATCH=a,b,c: blocks of the first, second and every further sub-batch).
Each schedule runs in its own process (the schedule is read once per context)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CHILD = r'''
import os, sys, time
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np, torch
import synth
from kanzi_b200 import Context
size = int(sys.argv[2]) << 20
data = synth.synth_compressible(size, 2)
ctx = Context(0, 4 << 20, 256)
host = torch.from_numpy(data).pin_memory().numpy()
outc = torch.empty(size + size // 2, dtype=torch.uint8).pin_memory().numpy()
outp = torch.empty(size, dtype=torch.uint8).pin_memory().numpy()
c = ctx.compress(host, "BWT+RANK+ZRLT", "ANS0", 4 << 20, out=outc)
ctx.decompress(c, size, out=outp)
te = td = 0.0
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = ctx.compress(host, "BWT+RANK+ZRLT", "ANS0", 4 << 20, out=outc)
    t1 = time.perf_counter()
    ctx.decompress(c, size, out=outp)
    t2 = time.perf_counter()
    te += t1 - t0; td += t2 - t1
assert np.array_equal(outp, data)
print("sched %-12s enc %.1f ms dec %.1f ms  e2e %.0f MB/s" % (os.environ.get("KNZ_ENC_BATCH", "default"), te / 3 * 1e3, td / 3 * 1e3,
                                                              size / ((te + td) / 3) / 1e6), flush=True)
'''
for sched in (sys.argv[2:] or ["32,96,128", "32,64,80", "24,56,88", "16,48,64", "48,104,104", "64,96,96"]):
    env = dict(os.environ, KNZ_ENC_BATCH=sched)
    subprocess.run([sys.executable, "-c", CHILD, ROOT, sys.argv[1] if len(sys.argv) > 1 else "1024"], env=env, timeout=300)
