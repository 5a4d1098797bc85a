"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref,
built from /root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

Each record names a deterministic input (generator, seed, size -- see
kanzi-cpp_b200/synth.py and tests/cases.py), the pipeline, the block size, and
the reference's compressed stream: full bytes (hex) when small, else length +
SHA-256.  The reference ships no golden vectors of its own, so
these fixtures are what pins the oracle and the GPU path on boxes without
/root/reference.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200"), os.path.join(ROOT, "tests")]

import synth  # noqa: E402
from cases import small_cases  # noqa: E402
from oracle.oracle import Ref  # noqa: E402

STREAMS = [
    # (generator, seed, size, transform, entropy, block size)
    ("text", 1, 70000, "NONE", "ANS0", 65536),
    ("text", 1, 70000, "BWT+RANK+ZRLT", "ANS0", 65536),
    ("compressible", 2, 300000, "BWT+RANK+ZRLT", "ANS0", 65536),
    ("compressible", 2, 300000, "BWT+MTFT+ZRLT", "ANS0", 1 << 18),
    ("incompressible", 9, 100000, "BWT+RANK+ZRLT", "ANS0", 65536),
    ("compressible", 3, 65536 * 2 + 9, "ZRLT", "ANS0", 65536),
    ("compressible", 4, 200000, "BWT", "NONE", 65536),
    ("compressible", 2, 9 << 20, "BWT+RANK+ZRLT", "ANS0", 4 << 20),
    ("text", 1, 16 << 20, "NONE", "ANS0", 4 << 20),
    ("text", 1, 16 << 20, "NONE", "HUFFMAN", 4 << 20),  # BASELINE.json configs[0]
    ("compressible", 5, 200000, "BWT+RANK+ZRLT", "HUFFMAN", 65536),
    ("compressible", 2, 64 << 20, "BWT+RANK+ZRLT", "ANS0", 4 << 20),
    # order-1 rANS (BASELINE.json configs[3])
    ("text", 1, 70000, "NONE", "ANS1", 65536),
    ("compressible", 2, 300000, "BWT+RANK+ZRLT", "ANS1", 65536),
    ("incompressible", 9, 100000, "NONE", "ANS1", 65536),
    ("compressible", 4, 9 << 20, "NONE", "ANS1", 4 << 20),
    ("compressible", 4, (9 << 20) + 3, "NONE", "ANS1", 8 << 20),  # two order-1 chunks per block
    # BASELINE.json configs[4] shape: BWT+SRT+ZRLT / FPAQ
    ("compressible", 5, 300000, "BWT+SRT+ZRLT", "FPAQ", 65536),
    ("text", 1, 70000, "SRT", "ANS0", 65536),
    ("incompressible", 9, 100000, "NONE", "FPAQ", 65536),
    ("compressible", 5, 9 << 20, "BWT+SRT+ZRLT", "FPAQ", 4 << 20),
    ("compressible", 5, (40 << 20) + 5, "BWT+SRT+ZRLT", "FPAQ", 32 << 20),
    # LZ family
    ("compressible", 6, 300000, "LZ", "HUFFMAN", 65536),
    ("text", 1, 70000, "LZX", "ANS0", 65536),
    ("compressible", 6, 300000, "LZP", "NONE", 65536),
    ("compressible", 6, (9 << 20) + 7, "LZ", "HUFFMAN", 4 << 20),
    ("compressible", 6, (9 << 20) + 7, "LZX", "HUFFMAN", 4 << 20),
    ("compressible", 6, (9 << 20) + 7, "LZP+LZX", "ANS0", 4 << 20),
    # block checksums (7th field: 32 = XXHash32, 64 = XXHash64)
    ("text", 1, 70000, "BWT+RANK+ZRLT", "ANS0", 65536, 32),
    ("text", 1, 70000, "BWT+RANK+ZRLT", "ANS0", 65536, 64),
    ("compressible", 2, (9 << 20) + 11, "BWT+RANK+ZRLT", "ANS0", 4 << 20, 32),
    ("compressible", 2, (9 << 20) + 11, "NONE", "HUFFMAN", 4 << 20, 64),
    ("incompressible", 9, 100000, "NONE", "NONE", 65536, 64),
]


def main():
    ref = Ref.load()
    assert ref is not None, "oracle/_ref missing: run `make -C oracle ref` where /root/reference exists"
    out = {"streams": [], "stages": []}
    for gen, seed, size, tname, ename, bs, *rest in STREAMS:
        ck = rest[0] if rest else 0
        data = synth.GENERATORS[gen](size, seed)
        # jobs=1: the buffer model the GPU path reproduces
        comp = ref.stream_compress(data, tname, ename, bs, jobs=1, checksum=ck)
        rec = {"gen": gen, "seed": seed, "size": size, "transform": tname, "entropy": ename, "block": bs, "checksum": ck,
               "input_sha256": synth.sha256(data), "len": int(comp.size),
               "sha256": hashlib.sha256(comp.tobytes()).hexdigest()}
        if comp.size <= 50000:
            rec["hex"] = comp.tobytes().hex()
        out["streams"].append(rec)
        print(gen, size, tname, ename, bs, "->", comp.size)
    # stage-level vectors on the named small cases (reference BWT bytes + primary indexes, ANS0 bit strings)
    cases = small_cases()
    for name in ("mississippi", "pi", "sixmixed", "rnd256_257", "rnd4_1000", "zeros_5000", "fe_ff_heavy",
                 "text_16387", "two_sym_9000"):
        data = cases[name]
        rec = {"case": name, "input_hex": data.tobytes().hex()}
        if data.size >= 2:
            bwt, pidx = ref.bwt_forward(data)
            rec["bwt_hex"] = bwt.tobytes().hex()
            rec["primary"] = pidx[: (8 if data.size >= 256 else 1)]
        enc, bits = ref.entropy_encode("ANS0", data)
        rec["ans0_hex"] = enc.tobytes().hex()
        rec["ans0_bits"] = int(bits)
        enc1, bits1 = ref.entropy_encode("ANS1", data)
        rec["ans1_hex"] = enc1.tobytes().hex()
        rec["ans1_bits"] = int(bits1)
        encf, bitsf = ref.entropy_encode("FPAQ", data)
        rec["fpaq_hex"] = encf.tobytes().hex()
        rec["fpaq_bits"] = int(bitsf)
        for t in ("ZRLT", "RANK", "MTFT"):
            o, fl, ok = ref.sequence_forward(t, data, data.size + 64, data.size + 64)
            rec[t.lower() + "_hex"] = o.tobytes().hex() if fl != 0xFF else None
        o, fl, ok = ref.sequence_forward("SRT", data, data.size + 1152, data.size + 1152)
        rec["srt_hex"] = o.tobytes().hex() if fl != 0xFF else None
        out["stages"].append(rec)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote golden.json", os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
