"""Scratch probe (2+ GPU box, torchrun): step-by-step progress of the multi-GPU C-ABI path."""
import os, sys, time, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "kanzi-cpp_b200")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
import synth
from kanzi_b200 import Context, E_IDS, _ptr

def log(*a):
    print(f"[rank {os.environ.get('RANK')}] {time.time():.2f}", *a, file=sys.stderr, flush=True)

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
log("pg up")
BS = 1 << 20
ctx = Context(lr, BS, 8)
log("ctx created")
ctx.dist_init(rank, world)
log("dist_init done")
data = synth.synth_compressible(13 * BS + 777, 3)
comp = ctx.compress_dist(data, "BWT+RANK+ZRLT", "ANS0", BS)
log("compress_dist done", comp.size)
n = torch.tensor([comp.size], dtype=torch.int64, device=dev)
dist.broadcast(n, 0)
buf = torch.zeros(int(n.item()), dtype=torch.uint8, device=dev)
if rank == 0:
    buf.copy_(torch.from_numpy(comp))
dist.broadcast(buf, 0)
comp = buf.cpu().numpy()
log("stream shared")
out = np.zeros(data.size, dtype=np.uint8)
back = ctx.decompress_dist(comp, data.size, out=out)
ok = all(np.array_equal(out[i * BS:(i + 1) * BS], data[i * BS:(i + 1) * BS]) for i in range(rank, 14, world))
log("decompress_dist done", back.size, ok)
if rank == 0:
    ref = Context(lr, BS, 8)
    want = ref.compress(data, "BWT+RANK+ZRLT", "ANS0", BS)
    log("matches single-GPU stream:", bool(want.size == comp.size and np.array_equal(want, comp)))
    ref.close()
dist.barrier()
ctx.close()
log("ctx closed")
dist.barrier()
dist.destroy_process_group()
log("done")
