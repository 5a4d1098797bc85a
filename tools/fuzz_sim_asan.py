"""Corrupted streams through the decoders of the AddressSanitizer build of the emulator library (not a test):
  make -C tests/sim -j8 asan
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
      python tools/fuzz_sim_asan.py [seconds=300] [seed=1] [streams|stages]
Every valid stream (random pipeline, block size, checksum) is damaged six ways (bit flips, byte overwrites, 0xFF
runs, truncation) and decoded; the decoder may reject or accept, ASan aborts on any out-of-bounds access.
`stages`: arbitrary bytes straight into every inverse transform and entropy decoder of the stage-level ABI."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+'/kanzi-cpp_b200', ROOT+'/tests']
import numpy as np, synth
from kanzi_b200 import Context, KanziGpuError
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng=np.random.default_rng(seed)
sim=Context(0,1<<16,4,lib_path=os.path.join(ROOT, 'tests', 'sim', 'libknzsim_asan.so'))
PIPES=[("BWT+RANK+ZRLT","ANS0"),("BWT+MTFT+ZRLT","HUFFMAN"),("ZRLT","ANS0"),("ZRLT","NONE"),("BWT+SRT+ZRLT","FPAQ"),("NONE","ANS1"),("LZ+ZRLT","HUFFMAN"),("LZX","ANS0"),("LZP","NONE"),("TEXT+UTF+PACK+MM+LZX","HUFFMAN"),("RANK","NONE")]
mode = sys.argv[3] if len(sys.argv) > 3 else "streams"
if mode == "stages":
    t0 = time.time(); runs = errs = 0
    alph = [np.arange(256, dtype=np.uint8), np.array([0, 1, 0xFF, 2, 0xFE], dtype=np.uint8), np.array([0, 0, 0, 1, 7, 0x80], dtype=np.uint8)]
    while time.time() - t0 < budget:
        n = int(rng.choice([rng.integers(1, 64), rng.integers(64, 5000), rng.integers(5000, 70000)]))
        a = alph[rng.integers(0, len(alph))]
        x = a[rng.integers(0, a.size, n)]
        if rng.random() < 0.3:  # plausible headers: small little-endian / varint fields up front
            x[: min(n, 16)] = rng.integers(0, 4, min(n, 16))
        for t in ("ZRLT", "RANK", "MTFT", "BWT", "SRT", "LZ", "LZX", "LZP", "PACK", "DNA", "MM", "UTF", "TEXT"):
            try:
                sim.transform_inverse(t, x, int(rng.integers(1, 70000)))
            except KanziGpuError:
                errs += 1
            runs += 1
        for e in ("ANS0", "ANS1", "HUFFMAN", "FPAQ"):
            try:
                sim.entropy_decode(e, x, int(rng.integers(1, 8 * n + 1)), int(rng.integers(1, 66000)))
            except KanziGpuError:
                errs += 1
            runs += 1
    print("asan fuzz (stages): %d calls on arbitrary bytes (%d rejected), no sanitizer report, %.0f s" % (runs, errs, time.time() - t0))
    sys.exit(0)
t0=time.time(); runs=errs=oks=0
while time.time()-t0<budget:
    n=int(rng.integers(100,60000)); bs=int(rng.choice([1024,4096,16384,65536]))
    t,e=PIPES[rng.integers(0,len(PIPES))]
    data=synth.synth_compressible(n,int(rng.integers(1,1<<30))) if rng.random()<0.7 else rng.integers(0,256,n,dtype=np.uint8)
    ck=int(rng.choice([0,32,64]))
    sim.set_checksum(ck); comp=sim.compress(data,t,e,bs); sim.set_checksum(0)
    for m in range(6):
        bad=comp.copy()
        k=int(rng.integers(1,4))
        for _ in range(k):
            pos=int(rng.integers(20 if rng.random()<0.8 else 0,bad.size))
            mode=rng.integers(0,3)
            if mode==0: bad[pos]^=1<<int(rng.integers(0,8))
            elif mode==1: bad[pos]=rng.integers(0,256)
            else: bad[pos:pos+int(rng.integers(1,16))]=0xFF
        if rng.random()<0.2: bad=bad[:int(rng.integers(1,bad.size))]
        try:
            out=sim.decompress(bad,n+1024)
            oks+=1
        except KanziGpuError:
            errs+=1
        runs+=1
print("asan fuzz: %d corrupted streams decoded (%d rejected, %d accepted), no sanitizer report, %.0f s"%(runs,errs,oks,time.time()-t0))
