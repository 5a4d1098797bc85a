"""Differential fuzz of the kernels' logic on the emulator build against the unmodified reference (oracle/_ref):
random generators x sizes x block sizes x pipelines, stream bytes and round trips.  Not a test (open ended):
  python tools/fuzz_sim.py [seconds=300] [seed=1]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import synth
from kanzi_b200 import Context, KanziGpuError
from oracle.oracle import Ref

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ref = Ref.load()
assert ref is not None, "needs oracle/_ref (make -C oracle ref)"
sim = Context(0, 1 << 18, 4, lib_path=os.path.join(ROOT, "tests", "sim", "libknzsim.so"))
PIPES = [("BWT+RANK+ZRLT", "ANS0"), ("BWT+MTFT+ZRLT", "HUFFMAN"), ("ZRLT", "ANS0"), ("RANK", "NONE"), ("ZRLT", "NONE"),
         ("BWT+SRT+ZRLT", "ANS0"), ("NONE", "ANS1"), ("LZ+ZRLT", "HUFFMAN"), ("LZX", "ANS0"), ("LZP", "NONE"),
         ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0"), ("MTFT+ZRLT", "ANS0"), ("BWT+SRT+ZRLT", "FPAQ"), ("LZP+LZX", "HUFFMAN"),
         ("TEXT+UTF+PACK+MM+LZX", "HUFFMAN"), ("DNA+LZ", "HUFFMAN"), ("SRT", "ANS1"), ("BWT", "NONE"), ("NONE", "HUFFMAN"),
         ("SRT+ZRLT", "NONE"), ("BWT+SRT", "ANS0"), ("PACK+LZX", "ANS0"), ("MM+LZ", "NONE"), ("UTF+BWT+RANK+ZRLT", "HUFFMAN")]


def gen(n):
    k = rng.integers(0, 7)
    if k == 0:
        return rng.integers(0, 256, n, dtype=np.uint8)
    if k == 1:
        a = rng.integers(0, 256, n, dtype=np.uint8)
        a[rng.random(n) < rng.random()] = 0
        return a
    if k == 2:
        return synth.synth_text(n, int(rng.integers(1, 1 << 30))) if n else np.zeros(0, dtype=np.uint8)
    if k == 3:
        return synth.synth_compressible(n, int(rng.integers(1, 1 << 30))) if n else np.zeros(0, dtype=np.uint8)
    if k == 4:  # runs of random length of a few symbols, 0 / 0xFE / 0xFF among them
        syms = np.array([0, 0, 0xFF, 0xFE, 1, 2, 97], dtype=np.uint8)
        out = np.empty(n, dtype=np.uint8)
        i = 0
        while i < n:
            L = int(rng.geometric(1.0 / rng.choice([2, 9, 40, 3000])))
            out[i:i + L] = syms[rng.integers(0, syms.size)]
            i += L
        return out
    if k == 5:
        return np.cumsum(rng.integers(-3, 4, n)).astype(np.uint8)
    return (rng.integers(0, 4, n) * 85).astype(np.uint8)


t0, runs, skipped, toobig = time.time(), 0, 0, 0
while time.time() - t0 < budget:
    n = int(rng.choice([rng.integers(0, 300), rng.integers(300, 20000), rng.integers(20000, 150000)]))
    bs = int(rng.choice([1024, 4096, 16384, 65536, 1 << 17])) + 16 * int(rng.integers(0, 8))
    tname, ename = PIPES[rng.integers(0, len(PIPES))]
    ck = int(rng.choice([0, 0, 32, 64]))
    data = gen(n)
    sim.set_checksum(ck)
    try:
        got = sim.compress(data, tname, ename, bs)
    finally:
        sim.set_checksum(0)
    try:
        want = ref.stream_compress(data, tname, ename, bs, 1, ck)
    except AssertionError:  # the shim's output buffer (1.5 n + 64 KiB) is too small for a stream that doubles
        toobig += 1
        continue
    same = got.size == want.size and np.array_equal(got, want)
    back = sim.decompress(got, max(n, 1))
    rt = back.size == n and np.array_equal(back, data)
    if not rt or not same:
        # reference corners documented in DESIGN.md (quirks 1, 2): an expanding ZRLT at odd parity / reused buffers
        r2, rc = ref.stream_decompress(got, max(n, 1))
        known = rt and rc == 0 and np.array_equal(r2[:n], data) and "ZRLT" in tname
        if known:
            skipped += 1
        else:
            np.save("/tmp/fuzz_fail.npy", data)
            print("MISMATCH", dict(n=n, bs=bs, t=tname, e=ename, ck=ck, same=same, rt=rt), flush=True)
            sys.exit(1)
    runs += 1
print("fuzz: %d runs, %d known reference corners, %d beyond the shim's buffer, 0 mismatches in %.0f s" % (runs, skipped, toobig, time.time() - t0))
