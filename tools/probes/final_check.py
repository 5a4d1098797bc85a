"""Last GPU check of round 2 (no torch import: the box has ~70 s): the headline workload through the C ABI --
stream SHA-256 against the reference's, round trip, per-stage times of a full 256-block batch -- then ZRLT and
RANK stage parity against the oracle on inputs aimed at the mask walks / the deep-rank step.
  python tools/probes/final_check.py [blocks=256] [lib]"""
import ctypes, hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import synth
from kanzi_b200 import Context, E_IDS, _ptr

REF_SHA_1GIB = "79607d0602e7356633c69833a669b6ac46e7f0007dd19119fb4b8f41385e59d6"  # bench.py --impl reference
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = sys.argv[2] if len(sys.argv) > 2 else None
bs = (4 << 20) if lib is None else (1 << 16)
T0 = time.time()


def note(*a):
    print("[%.1fs]" % (time.time() - T0), *a, flush=True)


data = synth.synth_compressible(nb * bs, 2)
ctx = Context(0, bs, nb, lib_path=lib) if lib else Context(0, bs, nb)
note("context ready")
tt, et = ctx.transform_type("BWT+RANK+ZRLT"), E_IDS["ANS0"]
ostride = (bs + bs // 4 + 4096 + 131072 * (bs // (4 << 20) + 1) + 255) // 256 * 256
enc = np.empty(nb * ostride, dtype=np.uint8)
bits = np.zeros(nb, dtype=np.uint64)
flags = np.zeros(nb, dtype=np.uint8)
lens = np.full(nb, bs, dtype=np.int32)
dec = np.empty(nb * bs, dtype=np.uint8)
dlens = np.zeros(nb, dtype=np.int32)
SKIP_BIG = os.environ.get("FINAL_CHECK_SKIP_BIG") == "1"  # stage / small-stream parity only
for rep in range(0 if SKIP_BIG else 2):
    ctx._check(ctx.lib.knz_encode_blocks(ctx.h, tt, et, bs, _ptr(data), bs, _ptr(lens), nb, bs, _ptr(enc), ostride,
                                         _ptr(bits), _ptr(flags)))
    e = ctx.timings()
    ctx._check(ctx.lib.knz_decode_blocks(ctx.h, tt, et, bs, _ptr(enc), ostride, _ptr(bits), nb, _ptr(dec), bs, _ptr(dlens)))
    d = ctx.timings()
if SKIP_BIG:
    e = d = {}
    dec, dlens = data, lens
note("batch of", nb, "blocks: round trip", bool(np.array_equal(dec, data)) and bool((dlens == bs).all()))
print(" enc", {k: round(v, 3) for k, v in e.items()})
print(" dec", {k: round(v, 3) for k, v in d.items()})
if SKIP_BIG:
    data = data[: 8 * bs]
comp = ctx.compress(data, "BWT+RANK+ZRLT", "ANS0", bs)
sha = hashlib.sha256(comp.tobytes()).hexdigest()
note("stream", comp.size, "bytes sha256", sha, "matches reference" if sha == REF_SHA_1GIB else
     ("(no reference hash for this size)" if (nb, bs) != (256, 4 << 20) or SKIP_BIG else "DIFFERS FROM THE REFERENCE"))
back = ctx.decompress(comp, data.size)
note("stream round trip", bool(np.array_equal(back, data)))

# stage parity against the oracle
from oracle.oracle import Oracle
from test_sim_kernels import _zrlt_inputs
oracle = Oracle()
bad = 0
rng = np.random.default_rng(5)
cases = dict(_zrlt_inputs())
for name, x in cases.items():
    n = x.size
    a, applied = ctx.transform_forward("ZRLT", x, 2 * n + 64)
    b, fl = oracle.sequence_forward("ZRLT", x, n, 2 * n + 64)
    ok = applied == (fl != 0xFF) and (not applied or (a.size == b.size and np.array_equal(a, b)))
    if ok and applied:
        r, ok2 = ctx.transform_inverse("ZRLT", b, n + 64)
        ok = ok2 and np.array_equal(r, x)
    bad += not ok
    if not ok:
        print("ZRLT MISMATCH", name)
big = synth.synth_compressible(bs, 11)
big[rng.random(bs) < 0.6] = 0
rk = {"random": rng.integers(0, 256, min(bs, 1 << 18), dtype=np.uint8), "sparse": big[: min(bs, 1 << 20)],
      "walk": np.cumsum(rng.integers(-3, 4, min(bs, 1 << 18))).astype(np.uint8)}
for tname in ("ZRLT", "RANK", "MTFT"):
    for name, x in rk.items():
        n = x.size
        a, applied = ctx.transform_forward(tname, x, 2 * n + 64)
        b, fl = oracle.sequence_forward(tname, x, n, 2 * n + 64)
        ok = applied == (fl != 0xFF) and (not applied or (a.size == b.size and np.array_equal(a, b)))
        if ok and applied:
            r, ok2 = ctx.transform_inverse(tname, b, n + 64)
            ok = ok2 and np.array_equal(r, x)
        bad += not ok
        if not ok:
            print("STAGE MISMATCH", tname, name)
note("stage parity vs oracle:", "all equal" if bad == 0 else "%d MISMATCHES" % bad)
from cases import small_cases
bad = 0
for name, x in small_cases().items():
    for tname, ename in (("BWT+RANK+ZRLT", "ANS0"), ("ZRLT", "HUFFMAN"), ("BWT+MTFT+ZRLT", "ANS0")):
        sbs = 1 << 16
        got = ctx.compress(x, tname, ename, sbs)
        want = oracle.stream_compress(x, tname, ename, sbs)
        ok = got.size == want.size and np.array_equal(got, want)
        if ok:
            ok = np.array_equal(ctx.decompress(got, x.size), x)
        bad += not ok
        if not ok:
            print("STREAM MISMATCH", name, tname, ename)
note("small-case streams vs oracle:", "all equal" if bad == 0 else "%d MISMATCHES" % bad)
ctx.close()
