"""Scratch probe (GPU box): BASELINE config 4 shape -- -t NONE -e {HUFFMAN,ANS0,ANS1} over block sizes,
device-resident blocks, kernel-only and stage times from the library's CUDA events."""
import os, sys, time, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "kanzi-cpp_b200")):
    sys.path.insert(0, p)
import torch
import synth
from kanzi_b200 import Context, E_IDS, sharded

total = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 256 << 20
data = synth.synth_compressible(total, 4)
dev = torch.device("cuda", 0)
d_all = torch.from_numpy(data).to(dev)
for bs in (64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 32 << 20):
    nblocks = total // bs
    batch = min(nblocks, max(1, (64 << 20) // bs * 4))
    for ename in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("HUFFMAN", "ANS0", "ANS1")):
        ctx = Context(0, bs, batch)
        ostride = (bs + bs // 4 + 4096 + 131072 * (bs // (4 << 20) + 1) + 255) // 256 * 256
        d_in = d_all[: nblocks * bs].view(nblocks, bs)
        d_blk = torch.zeros((nblocks, ostride), dtype=torch.uint8, device=dev)
        d_bits = torch.zeros(nblocks, dtype=torch.int64, device=dev)
        d_dec = torch.empty((nblocks, bs), dtype=torch.uint8, device=dev)
        lens = np.full(nblocks, bs, dtype=np.int32)
        tt, et = ctx.transform_type("NONE"), E_IDS[ename]
        for rep in range(2):
            sharded.encode_shard(ctx, tt, et, bs, d_in, lens, bs, d_blk, d_bits)
            te = ctx.timings()
            try:
                sharded.decode_shard(ctx, tt, et, bs, d_blk, d_bits.cpu().numpy().astype(np.uint64), d_dec)
            except RuntimeError as ex:  # ANS1: blocks the reference cannot decode either
                print("   decode:", str(ex)[:80])
            td = ctx.timings()
        ok = bool(torch.equal(d_dec, d_in))
        e = int((d_bits.sum().item() + 7) // 8)
        gb = (nblocks * bs + e) / 1e9
        print(f"bs={bs>>10:6d}K {ename:8s} ok={ok} ratio={e/(nblocks*bs):.3f} "
              f"enc stage {te['entropy']:8.2f} ms kernel {te['ans_enc_kernel']:8.2f} ms ({gb/ (te['ans_enc_kernel'] or 1e9)*1e3:7.1f} GB/s)  "
              f"dec stage {td['entropy']:8.2f} ms kernel {td['ans_dec_kernel']:8.2f} ms ({gb/(td['ans_dec_kernel'] or 1e9)*1e3:7.1f} GB/s)", flush=True)
        ctx.close()
        del d_blk, d_bits, d_dec
        torch.cuda.empty_cache()
