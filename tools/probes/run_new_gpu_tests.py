"""The three GPU tests added last (ZRLT mask walks, ZRLT token classes, deep RANK steps) without pytest / torch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200"), os.path.join(ROOT, "tests")]
t0 = time.time()
from kanzi_b200 import Context
from oracle.oracle import Oracle
import test_gpu_parity as T
ctx, o = Context(0, 4 << 20, 64), Oracle()
for f in (T.test_gpu_zrlt_mask_walks, T.test_gpu_zrlt_inverse_token_classes, T.test_gpu_rank_deep_steps):
    f(ctx, o)
    print("[%.1fs] %s passed" % (time.time() - t0, f.__name__), flush=True)
ctx.close()
