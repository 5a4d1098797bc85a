"""ctypes loaders for the CPU oracle (TEST INFRASTRUCTURE -- see oracle/kanzi_oracle.c header).

  Oracle()  -> oracle/liboracle.so     plain-C restatement ("port")
  Ref()     -> oracle/_ref/libkanzi_ref.so   the unmodified reference behind ref_shim.cpp
               (prebuilt in the build container; None if absent)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

T_IDS = {"NONE": 0, "BWT": 1, "LZ": 3, "ZRLT": 6, "MTFT": 7, "RANK": 8, "SRT": 13, "LZP": 14, "LZX": 16, "TEXT": 10, "MM": 15, "UTF": 17, "PACK": 18, "DNA": 19}
E_IDS = {"NONE": 0, "HUFFMAN": 1, "FPAQ": 2, "ANS0": 5, "ANS1": 8}


def transform_word(name):
    """'BWT+RANK+ZRLT' -> 48-bit type word (TransformFactory.hpp:100-137)."""
    word, shift = 0, 42
    for tok in name.split("+"):
        t = T_IDS[tok]
        if t != 0:
            word |= t << shift
            shift -= 6
    return word


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle oracle`")
        L = self.lib = ctypes.CDLL(path)
        L.ko_entropy_encode.restype = ctypes.c_int64
        L.ko_encode_block.restype = ctypes.c_int64
        L.ko_stream_compress.restype = ctypes.c_int64
        L.ko_stream_compress_ck.restype = ctypes.c_int64
        L.ko_stream_decompress.restype = ctypes.c_int64

    def entropy_encode(self, name, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(data.size * 2 + 65536, dtype=np.uint8)
        bits = self.lib.ko_entropy_encode(E_IDS[name], _ptr(data), data.size, _ptr(out), ctypes.c_int64(out.size))
        assert bits >= 0
        return out[: (bits + 7) // 8].copy(), bits

    def entropy_decode(self, name, enc, nbits, n):
        enc = np.ascontiguousarray(enc, dtype=np.uint8)
        out = np.zeros(max(n, 1), dtype=np.uint8)
        rc = self.lib.ko_entropy_decode(E_IDS[name], _ptr(enc), ctypes.c_int64(nbits), _ptr(out), n)
        return out[:n], rc

    def sequence_forward(self, name, data, in_cap=None, out_cap=None):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n = data.size
        in_cap = n if in_cap is None else in_cap
        out_cap = n + 64 if out_cap is None else out_cap
        out = np.zeros(out_cap + 64, dtype=np.uint8)
        flags = ctypes.c_int(0)
        m = self.lib.ko_sequence_forward(ctypes.c_uint64(transform_word(name)), _ptr(data), n, in_cap, _ptr(out),
                                         out_cap, ctypes.byref(flags))
        return out[:m].copy(), flags.value

    def sequence_inverse(self, name, flags, data, out_cap):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(out_cap + 64, dtype=np.uint8)
        ol = ctypes.c_int(0)
        ok = self.lib.ko_sequence_inverse(ctypes.c_uint64(transform_word(name)), flags, _ptr(data), data.size,
                                          _ptr(out), out_cap, ctypes.byref(ol))
        return out[: ol.value].copy(), ok

    def bwt_forward(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(data.size, dtype=np.uint8)
        pidx = (ctypes.c_int * 8)()
        self.lib.ko_bwt_forward(_ptr(data), data.size, _ptr(out), pidx)
        return out, list(pidx)

    def suffix_ranks(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        rk = np.zeros(data.size, dtype=np.int32)
        self.lib.ko_suffix_ranks(_ptr(data), data.size, _ptr(rk))
        return rk

    def encode_block(self, data, tname, ename, data_cap, buf_cap):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(data.size * 2 + 65536, dtype=np.uint8)
        bits = self.lib.ko_encode_block(_ptr(data), data.size, ctypes.c_uint64(transform_word(tname)), E_IDS[ename],
                                        data_cap, buf_cap, _ptr(out), ctypes.c_int64(out.size))
        assert bits >= 0
        return out[: (bits + 7) // 8].copy(), bits

    def block_hash(self, data, bits):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        self.lib.ko_block_hash.restype = ctypes.c_uint64
        return int(self.lib.ko_block_hash(_ptr(data), data.size, bits))

    def stream_compress(self, data, tname, ename, block_size, checksum=0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(data.size + data.size // 2 + 65536, dtype=np.uint8)
        n = self.lib.ko_stream_compress_ck(_ptr(data), ctypes.c_int64(data.size), ctypes.c_uint64(transform_word(tname)),
                                           E_IDS[ename], block_size, checksum, _ptr(out), ctypes.c_int64(out.size))
        assert n >= 0, n
        return out[:n].copy()

    def stream_decompress(self, comp, cap):
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        n = self.lib.ko_stream_decompress(_ptr(comp), ctypes.c_int64(comp.size), _ptr(out), ctypes.c_int64(cap))
        return out[: max(n, 0)], n


class Ref:
    """The unmodified reference (oracle/_ref).  Ref.load() returns None when absent."""

    @staticmethod
    def load():
        path = os.path.join(_HERE, "_ref", "libkanzi_ref.so")
        return Ref(path) if os.path.exists(path) else None

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)

    def stream_compress(self, data, tname, ename, block_size, jobs=1, checksum=0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(data.size + data.size // 2 + 65536, dtype=np.uint8)
        ol = ctypes.c_int64(0)
        rc = self.lib.kref_stream_compress(_ptr(data), ctypes.c_int64(data.size), tname.encode(), ename.encode(),
                                           block_size, jobs, checksum, _ptr(out), ctypes.c_int64(out.size), ctypes.byref(ol))
        assert rc == 0, rc
        return out[: ol.value].copy()

    def stream_decompress(self, comp, cap, jobs=1):
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        ol = ctypes.c_int64(0)
        rc = self.lib.kref_stream_decompress(_ptr(comp), ctypes.c_int64(comp.size), jobs, _ptr(out),
                                             ctypes.c_int64(cap), ctypes.byref(ol))
        return out[: ol.value], rc

    EVT_FIELDS = ("type", "blockId", "size", "hash", "hashType", "offset", "skipFlags")

    def stream_compress_ctx(self, data, tname, ename, block_size, jobs=1, checksum=0, skip_blocks=0, events=False):
        """Through the reference's Context constructor (skipBlocks, listeners).  Returns (stream, events)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(data.size + data.size // 2 + 65536, dtype=np.uint8)
        ol = ctypes.c_int64(0)
        ev_cap = 8 * (data.size // block_size + 2)
        ev = np.zeros((ev_cap, 7), dtype=np.int64)
        ec = ctypes.c_int(0)
        rc = self.lib.kref_stream_compress_ctx(_ptr(data), ctypes.c_int64(data.size), tname.encode(), ename.encode(),
                                               block_size, jobs, checksum, int(skip_blocks), _ptr(out),
                                               ctypes.c_int64(out.size), ctypes.byref(ol),
                                               _ptr(ev) if events else None, ev_cap, ctypes.byref(ec))
        assert rc == 0, rc
        return out[: ol.value].copy(), [dict(zip(self.EVT_FIELDS, (int(x) for x in row))) for row in ev[: ec.value]]

    def stream_decompress_ctx(self, comp, cap, jobs=1, from_block=0, to_block=0, events=False):
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        ol = ctypes.c_int64(0)
        ev_cap = 65536
        ev = np.zeros((ev_cap, 7), dtype=np.int64)
        ec = ctypes.c_int(0)
        rc = self.lib.kref_stream_decompress_ctx(_ptr(comp), ctypes.c_int64(comp.size), jobs, from_block, to_block,
                                                 _ptr(out), ctypes.c_int64(cap), ctypes.byref(ol),
                                                 _ptr(ev) if events else None, ev_cap, ctypes.byref(ec))
        return out[: ol.value], rc, [dict(zip(self.EVT_FIELDS, (int(x) for x in row))) for row in ev[: ec.value]]

    def sequence_forward(self, name, data, in_cap=None, out_cap=None):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n = data.size
        in_cap = n if in_cap is None else in_cap
        out_cap = n + 64 if out_cap is None else out_cap
        out = np.zeros(out_cap + 64, dtype=np.uint8)
        ol, fl = ctypes.c_int(0), ctypes.c_int(0)
        ok = self.lib.kref_transform_forward(name.encode(), _ptr(data), n, in_cap, _ptr(out), out_cap,
                                             ctypes.byref(ol), ctypes.byref(fl))
        return out[: ol.value].copy(), fl.value, ok

    def sequence_inverse(self, name, flags, data, out_cap):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(out_cap + 64, dtype=np.uint8)
        ol = ctypes.c_int(0)
        ok = self.lib.kref_transform_inverse(name.encode(), flags, _ptr(data), data.size, _ptr(out), out_cap,
                                             ctypes.byref(ol))
        return out[: ol.value].copy(), ok

    def bwt_forward(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(data.size, dtype=np.uint8)
        pidx = (ctypes.c_int * 8)()
        ok = self.lib.kref_bwt_forward(_ptr(data), data.size, _ptr(out), pidx)
        assert ok == 1
        return out, list(pidx)

    def entropy_encode(self, name, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(data.size * 2 + 65536, dtype=np.uint8)
        bits = ctypes.c_int64(0)
        ok = self.lib.kref_entropy_encode(name.encode(), _ptr(data), data.size, _ptr(out), ctypes.c_int64(out.size),
                                          ctypes.byref(bits))
        assert ok == 1, ok
        return out[: (bits.value + 7) // 8].copy(), bits.value

    def entropy_decode(self, name, enc, n):
        enc = np.ascontiguousarray(enc, dtype=np.uint8)
        out = np.zeros(max(n, 1), dtype=np.uint8)
        br = ctypes.c_int64(0)
        ok = self.lib.kref_entropy_decode(name.encode(), _ptr(enc), ctypes.c_int64(enc.size), _ptr(out), n,
                                          ctypes.byref(br))
        return out[:n], ok, br.value
