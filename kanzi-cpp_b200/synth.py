"""Deterministic synthetic inputs for the kanzi block-pipeline benchmarks.

Thin ctypes wrapper over csrc/synth.c (libknzsynth.so, host-only, built by
__graft_entry__.build()).  Generators are seeded and return numpy.uint8 arrays;
SHA-256 of each generated buffer is recorded next to the results.

  synth_text(n, seed)           Zipf-like words, skewed 26-letter alphabet (config 1)
  synth_compressible(n, seed)   50 % text / 25 % 32-byte records / 25 % byte random
                                walk in [64 KiB, 1 MiB) segments + a 4 KiB copy from
                                ~1 MiB earlier every 64 KiB (configs 2, 4, 5)
  synth_incompressible(n, seed) uniform random bytes (skipped-stage edge cases)
  synth_silesia(n, seed)        45 % text / mark-up, 25 % executable-like, 12 % records, 18 % smooth
                                16-bit samples in [256 KiB, 2 MiB) segments (config 3, the `-l N` levels)
"""
import ctypes
import hashlib
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libknzsynth.so")
        if not os.path.exists(path):
            raise RuntimeError("libknzsynth.so missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _LIB = ctypes.CDLL(path)
    return _LIB


def _gen(fn, n, seed):
    n = int(n)
    out = np.empty(n, dtype=np.uint8)
    if n > 0:
        getattr(_lib(), fn)(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n), ctypes.c_uint64(seed))
    return out


def synth_text(n, seed=1):
    return _gen("knz_synth_text", n, seed)


def synth_compressible(n, seed=2):
    return _gen("knz_synth_compressible", n, seed)


def synth_incompressible(n, seed=9):
    return _gen("knz_synth_incompressible", n, seed)


def synth_silesia(n, seed=3):
    return _gen("knz_synth_silesia", n, seed)


def sha256(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


GENERATORS = {
    "text": synth_text,
    "compressible": synth_compressible,
    "incompressible": synth_incompressible,
    "silesia": synth_silesia,
}
