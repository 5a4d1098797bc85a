"""Shared seeded test inputs (edge cases modelled on the reference's own tests:
TestTransforms.cpp:887-975, TestEntropyCodec.cpp:450-558, TestBWT.cpp:28-169)."""
import numpy as np

import synth


def rng_bytes(n, seed, alphabet=256):
    r = np.random.RandomState(seed)
    return r.randint(0, alphabet, size=n).astype(np.uint8)


def small_cases():
    """name -> uint8 array; sizes the CPU oracle handles in milliseconds."""
    c = {}
    c["mississippi"] = np.frombuffer(b"mississippi", dtype=np.uint8).copy()
    c["pi"] = np.frombuffer(b"3.14159265358979323846264338327950288419716939937510", dtype=np.uint8).copy()
    c["sixmixed"] = np.frombuffer(b"SIX.MIXED.PIXIES.SIFT.SIXTY.PIXIE.DUST.BOXES", dtype=np.uint8).copy()
    for n in (1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 34, 40, 63, 64, 65, 255, 256, 257, 1000, 4099):
        c[f"rnd256_{n}"] = rng_bytes(n, 100 + n)
        c[f"rnd4_{n}"] = rng_bytes(n, 200 + n, 4)
    for n in (16, 256, 5000, 20000):
        c[f"const_{n}"] = np.full(n, 0x41, dtype=np.uint8)
        c[f"zeros_{n}"] = np.zeros(n, dtype=np.uint8)
    c["two_sym_9000"] = rng_bytes(9000, 7, 2) * 255
    c["fe_ff_heavy"] = (rng_bytes(20003, 8, 4) + 252).astype(np.uint8)
    z = rng_bytes(70001, 9, 256)
    z[rng_bytes(70001, 10, 100) < 97] = 0
    c["zero_heavy_70001"] = z
    c["runs_40000"] = np.repeat(rng_bytes(400, 11, 7), 100)
    c["text_50000"] = synth.synth_text(50000, 3)
    c["text_16384"] = synth.synth_text(16384, 4)
    c["text_16387"] = synth.synth_text(16387, 5)
    c["comp_200k"] = synth.synth_compressible(200000, 6)
    c["incomp_70000"] = synth.synth_incompressible(70000, 7)
    c["ramp_66000"] = (np.arange(66000) & 255).astype(np.uint8)
    return c
