"""Profiling target for ncu (not a test): one pass over every kernel family.
  python tools/probes/profile_all.py headline [nblocks]   BWT+RANK+ZRLT / ANS0, 4 MiB blocks, encode + decode
  python tools/probes/profile_all.py entropy  [nblocks]   -t NONE -e HUFFMAN / ANS0 / ANS1 on 4 MiB blocks, ANS1 on 64 KiB blocks
  python tools/probes/profile_all.py config5  [nblocks]   BWT+SRT+ZRLT / FPAQ, 4 MiB blocks
  python tools/probes/profile_all.py lz       [nblocks]   LZ / LZX / LZP with HUFFMAN, 4 MiB blocks (+ the reference's time)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200")]
import numpy as np, torch
import synth
from kanzi_b200 import Context, E_IDS, sharded

what = sys.argv[1] if len(sys.argv) > 1 else "headline"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)


def run(tname, ename, bs, nblocks, seed, reps=1):
    data = synth.synth_compressible(nblocks * bs, seed)
    ctx = Context(0, bs, nblocks)
    d_in = torch.from_numpy(data).view(nblocks, bs).to(dev)
    ostride = (bs + bs // 4 + 4096 + 131072 * (bs // (4 << 20) + 1) + 255) // 256 * 256
    d_blk = torch.zeros((nblocks, ostride), dtype=torch.uint8, device=dev)
    d_bits = torch.zeros(nblocks, dtype=torch.int64, device=dev)
    d_dec = torch.zeros((nblocks, bs), dtype=torch.uint8, device=dev)
    tt, et = ctx.transform_type(tname), E_IDS[ename]
    for _ in range(reps):
        sharded.encode_shard(ctx, tt, et, bs, d_in, [bs] * nblocks, bs, d_blk, d_bits)
        e = ctx.timings()
        try:
            sharded.decode_shard(ctx, tt, et, bs, d_blk, d_bits.cpu().numpy().astype(np.uint64), d_dec)
        except RuntimeError as ex:
            print("decode:", str(ex)[:100])
        d = ctx.timings()
    torch.cuda.synchronize()
    print(tname, ename, bs, nblocks, "round trip", bool(torch.equal(d_dec, d_in)))
    print(" enc", e)
    print(" dec", d)
    ctx.close()


if what == "headline":
    run("BWT+RANK+ZRLT", "ANS0", 4 << 20, nb, 2, reps=2)
elif what == "entropy":
    for en in ("HUFFMAN", "ANS0", "ANS1"):
        run("NONE", en, 4 << 20, nb, 4)
    run("NONE", "ANS1", 64 << 10, nb * 16, 4)
elif what == "config5":
    run("BWT+SRT+ZRLT", "FPAQ", 4 << 20, nb, 5)
elif what == "lz":
    import time
    from oracle.oracle import Ref
    ref = Ref.load()
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else ("LZ", "LZX", "LZP")
    for tn in only:
        run(tn, "HUFFMAN", 4 << 20, nb, 6, reps=(1 if len(sys.argv) > 3 else 2))
        if ref is not None and len(sys.argv) <= 3:
            data = synth.synth_compressible(nb * (4 << 20), 6)
            jobs = min(16, os.cpu_count() or 1)
            t0 = time.time()
            comp = ref.stream_compress(data, tn, "HUFFMAN", 4 << 20, jobs=jobs)
            t1 = time.time()
            ref.stream_decompress(comp, data.size, jobs=jobs)
            t2 = time.time()
            print(" reference jobs=%d: encode %.0f ms decode %.0f ms ratio %.3f" % (jobs, 1e3 * (t1 - t0), 1e3 * (t2 - t1),
                                                                                  comp.size / data.size))
