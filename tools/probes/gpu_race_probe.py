import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200"), os.path.join(ROOT, "tests")]
import synth
from kanzi_b200 import Context
from oracle.oracle import Oracle
oracle = Oracle()
bs = 1 << 18
ctx = Context(0, bs, 8, lib_path=os.environ.get('KNZ_LIB'))
data = synth.synth_compressible(5 * bs + 1000, 31)
res = []
for i in (2, 3):
    blk = data[i*bs:(i+1)*bs]
    b, flags = oracle.sequence_forward("BWT", blk, blk.size, blk.size + 64)
    for rep in range(4):
        a, applied = ctx.transform_forward("BWT", blk, blk.size + 64)
        d = np.nonzero(a != b)[0]
        res.append((i, rep, int(d.size), int(d[0]) if d.size else -1))
print(os.environ.get("KNZ_TX_ALLGATHER"), os.environ.get("KNZ_TX_REGEN_SLOW"), res)
