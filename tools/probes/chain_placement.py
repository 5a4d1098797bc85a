"""Where the inverse RANK chains run while the other block groups decode: decode groups x chains
per CTA x shared-memory request of the chain kernel.  Device time of knz_decode_blocks (CUDA events inside the
library), output compared with the input for every setting.  No torch import.  Needs the experimental hook
knz_tune_rank_inverse: apply tools/probes/chain_placement.patch and rebuild (result: profiles/r02d_chain_placement.log).
  python tools/probes/chain_placement.py [blocks=256] [lib]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np
import synth
from kanzi_b200 import Context, E_IDS, _ptr

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = sys.argv[2] if len(sys.argv) > 2 else None
bs = (4 << 20) if lib is None else (1 << 14)
T0 = time.time()
data = synth.synth_compressible(nb * bs, 2)
ctx = Context(0, bs, nb, lib_path=lib) if lib else Context(0, bs, nb)
ctx.lib.knz_tune_rank_inverse.argtypes = [ctypes.c_int, ctypes.c_int]
tt, et = ctx.transform_type("BWT+RANK+ZRLT"), E_IDS["ANS0"]
ostride = (bs + bs // 4 + 4096 + 131072 * (bs // (4 << 20) + 1) + 255) // 256 * 256
enc = np.empty(nb * ostride, dtype=np.uint8)
bits = np.zeros(nb, dtype=np.uint64)
flags = np.zeros(nb, dtype=np.uint8)
lens = np.full(nb, bs, dtype=np.int32)
dec = np.empty(nb * bs, dtype=np.uint8)
dlens = np.zeros(nb, dtype=np.int32)
ctx._check(ctx.lib.knz_encode_blocks(ctx.h, tt, et, bs, _ptr(data), bs, _ptr(lens), nb, bs, _ptr(enc), ostride,
                                     _ptr(bits), _ptr(flags)))
print("[%.1fs] encoded" % (time.time() - T0), flush=True)
K = 1024
settings = [(1, 1, 0), (1, 1, 0), (1, 4, 0), (1, 4, 200 * K),
            (4, 1, 0), (4, 4, 0), (4, 4, 100 * K), (4, 4, 200 * K), (4, 2, 100 * K), (4, 1, 56 * K),
            (8, 1, 0), (8, 4, 0), (8, 4, 100 * K), (8, 4, 200 * K), (2, 4, 200 * K), (2, 1, 0), (1, 1, 0)]
for groups, wpc, hog in settings:
    rc = ctx.lib.knz_tune_rank_inverse(wpc, hog)
    ctx._check(ctx.lib.knz_set_decode_groups(ctx.h, groups))
    ms = []
    for rep in range(2):
        dec[:: 4096] = 0xA5
        ctx._check(ctx.lib.knz_decode_blocks(ctx.h, tt, et, bs, _ptr(enc), ostride, _ptr(bits), nb, _ptr(dec), bs, _ptr(dlens)))
        ms.append(ctx.timings()["total"])
    ok = bool(np.array_equal(dec, data))
    print("groups %d chains/CTA %d hog %3d KiB (rc %d): %7.2f %7.2f ms  %s" % (groups, wpc, hog // K, rc, ms[0], ms[1], "ok" if ok else "WRONG OUTPUT"),
          flush=True)
ctx.lib.knz_tune_rank_inverse(1, 0)
print("[%.1fs] done" % (time.time() - T0))
ctx.close()
