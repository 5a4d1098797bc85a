"""CPU suite: the oracle against the committed golden fixtures (generated from the
unmodified reference by tests/golden/make_golden.py), and the C-ABI library's
loadability / exported symbols (no compute without a GPU)."""
import ctypes
import hashlib
import json
import os
import re

import numpy as np
import pytest

import synth
from cases import small_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def _input(rec):
    data = synth.GENERATORS[rec["gen"]](rec["size"], rec["seed"])
    assert synth.sha256(data) == rec["input_sha256"], "synthetic generator drifted from the fixtures"
    return data


@pytest.mark.parametrize("idx", range(len(GOLD["streams"])))
def test_oracle_stream_matches_golden(oracle, idx):
    rec = GOLD["streams"][idx]
    if rec["size"] > (10 << 20) and "BWT" in rec["transform"]:
        pytest.skip("oracle suffix sorter is too slow for this size on CPU; covered by the GPU suite")
    data = _input(rec)
    comp = oracle.stream_compress(data, rec["transform"], rec["entropy"], rec["block"], checksum=rec.get("checksum", 0))
    assert comp.size == rec["len"]
    assert hashlib.sha256(comp.tobytes()).hexdigest() == rec["sha256"]
    if "hex" in rec:
        assert comp.tobytes().hex() == rec["hex"]
    if rec["size"] <= (1 << 20):
        dec, n = oracle.stream_decompress(comp, data.size)
        assert n == data.size and np.array_equal(dec, data)


def test_oracle_stage_vectors(oracle):
    cases = small_cases()
    for rec in GOLD["stages"]:
        data = cases[rec["case"]]
        assert data.tobytes().hex() == rec["input_hex"]
        if "bwt_hex" in rec:
            bwt, pidx = oracle.bwt_forward(data)
            assert bwt.tobytes().hex() == rec["bwt_hex"], rec["case"]
            assert pidx[: len(rec["primary"])] == rec["primary"], rec["case"]
        enc, bits = oracle.entropy_encode("ANS0", data)
        assert bits == rec["ans0_bits"] and enc.tobytes().hex() == rec["ans0_hex"], rec["case"]
        enc, bits = oracle.entropy_encode("ANS1", data)
        assert bits == rec["ans1_bits"] and enc.tobytes().hex() == rec["ans1_hex"], rec["case"]
        enc, bits = oracle.entropy_encode("FPAQ", data)
        assert bits == rec["fpaq_bits"] and enc.tobytes().hex() == rec["fpaq_hex"], rec["case"]
        o, fl = oracle.sequence_forward("SRT", data, data.size + 1152, data.size + 1152)
        assert (o.tobytes().hex() if fl != 0xFF else None) == rec["srt_hex"], rec["case"]
        for t in ("ZRLT", "RANK", "MTFT"):
            o, fl = oracle.sequence_forward(t, data, data.size + 64, data.size + 64)
            want = rec[t.lower() + "_hex"]
            if want is None:
                assert fl == 0xFF
            else:
                assert o.tobytes().hex() == want, (rec["case"], t)


def test_mississippi_known_answer(oracle):
    # the only known answer the reference itself documents (transform/BWT.hpp:41-55)
    data = np.frombuffer(b"mississippi", dtype=np.uint8)
    bwt, pidx = oracle.bwt_forward(data)
    assert bwt.tobytes() == b"ipssmpissii" and pidx[0] == 5


def test_library_exports_declared_symbols():
    lib_path = os.path.join(ROOT, "kanzi-cpp_b200", "libknzgpu.so")
    assert os.path.exists(lib_path), "libknzgpu.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "knz_gpu.h")).read()
    names = set(re.findall(r"\b(knz_[a-z_0-9]+)\s*\(", header))
    assert len(names) >= 18
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/knz_gpu.h but not exported"


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kanzi_b200 import Context, KanziGpuError
    with pytest.raises(KanziGpuError):
        Context(0, 1 << 20, 2)  # there is no CPU fallback


def test_type_words():
    lib = ctypes.CDLL(os.path.join(ROOT, "kanzi-cpp_b200", "libknzgpu.so"))
    lib.knz_transform_type.restype = ctypes.c_uint64
    from oracle.oracle import transform_word
    for name in ("NONE", "BWT", "BWT+RANK+ZRLT", "BWT+MTFT+ZRLT", "ZRLT", "RANK+ZRLT", "BWT+SRT+ZRLT", "SRT", "LZ", "LZX", "LZP",
                 "LZP+LZX", "LZ+ZRLT"):
        assert lib.knz_transform_type(name.encode()) == transform_word(name)
    assert lib.knz_transform_type(b"ROLZ") == 0xFFFFFFFFFFFFFFFF
    assert lib.knz_entropy_type(b"ANS0") == 5 and lib.knz_entropy_type(b"ANS1") == 8
    assert lib.knz_entropy_type(b"FPAQ") == 2 and lib.knz_entropy_type(b"TPAQ") == -1
    hdr = (ctypes.c_uint8 * 32)()
    n = lib.knz_stream_header(ctypes.c_uint64(transform_word("BWT+RANK+ZRLT")), 5, 4 << 20, ctypes.c_int64(1 << 30), hdr)
    assert n == 24 and bytes(hdr[:4]) == b"KANZ"
