"""kanzi_b200 -- Python binding of libknzgpu.so (include/knz_gpu.h), the B200-native
replacement for kanzi's per-block transform -> entropy path.

This module is plumbing only: every byte of work happens in the hand-written
sm_100a kernels behind the C ABI.  There is no CPU fallback: constructing a
`Context` without the compiled library or without a CUDA device raises.

Class and method names mirror the reference's Python binding and stream API
(src/api/kanzi.py:18,86 Compressor/Decompressor; CompressedOutputStream::write/close).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libknzgpu.so")

T_IDS = {"NONE": 0, "BWT": 1, "LZ": 3, "ZRLT": 6, "MTFT": 7, "RANK": 8, "SRT": 13, "LZP": 14, "LZX": 16, "TEXT": 10, "MM": 15, "UTF": 17, "PACK": 18, "DNA": 19}
E_IDS = {"NONE": 0, "HUFFMAN": 1, "FPAQ": 2, "ANS0": 5, "ANS1": 8}


class _Event(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("blockId", ctypes.c_int), ("size", ctypes.c_int64), ("hash", ctypes.c_uint64),
                ("hashBits", ctypes.c_int), ("offset", ctypes.c_int64), ("skipFlags", ctypes.c_uint8)]


_EVENT_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(_Event))
EVENT_NAMES = {2: "BEFORE_TRANSFORM", 3: "AFTER_TRANSFORM", 4: "BEFORE_ENTROPY", 5: "AFTER_ENTROPY", 9: "BLOCK_INFO"}


class KanziGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"knz error {code}: {msg}")
        self.code = code


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _load(path):
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(no CPU fallback exists for this path)")
    L = ctypes.CDLL(path)
    L.knz_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.knz_destroy.argtypes = [ctypes.c_void_p]
    L.knz_last_error.restype = ctypes.c_char_p
    L.knz_last_error.argtypes = [ctypes.c_void_p]
    L.knz_transform_type.restype = ctypes.c_uint64
    L.knz_transform_type.argtypes = [ctypes.c_char_p]
    L.knz_entropy_type.argtypes = [ctypes.c_char_p]
    L.knz_set_checksum.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.knz_set_decode_groups.restype = ctypes.c_int
    L.knz_set_decode_groups.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.knz_launch_count.restype = ctypes.c_uint64
    L.knz_launch_count.argtypes = [ctypes.c_void_p]
    L.knz_stream.restype = ctypes.c_void_p
    L.knz_stream.argtypes = [ctypes.c_void_p]
    L.knz_last_timings.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
    L.knz_compress.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, i32, vp, i64, vp, i64, ctypes.POINTER(i64)]
    L.knz_decompress.argtypes = [vp, vp, i64, vp, i64, ctypes.POINTER(i64)]
    L.knz_decompress_range.argtypes = [vp, vp, i64, i32, i32, vp, i64, ctypes.POINTER(i64)]
    L.knz_decompress_seek.argtypes = [vp, vp, i64, i64, i32, vp, i64, ctypes.POINTER(i64)]
    L.knz_set_skip_blocks.argtypes = [vp, i32]
    L.knz_set_listener.argtypes = [vp, vp, vp]
    L.knz_encode_blocks.argtypes = [vp, u64, i32, i32, vp, i64, vp, i32, i32, vp, i64, vp, vp]
    L.knz_decode_blocks.argtypes = [vp, u64, i32, i32, vp, i64, vp, i32, vp, i64, vp]
    L.knz_encode_blocks_dev.argtypes = [vp, u64, i32, i32, vp, i64, vp, i32, i32, vp, i64, vp, vp]
    L.knz_decode_blocks_dev.argtypes = [vp, u64, i32, i32, vp, i64, vp, i32, vp, i64, vp]
    L.knz_assemble_stream_dev.argtypes = [vp, vp, i64, vp, i32, vp, i64, u64, ctypes.POINTER(u64)]
    L.knz_stream_header.argtypes = [u64, i32, i32, i64, vp]
    L.knz_transform_forward.argtypes = [vp, i32, vp, i32, vp, i32, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.knz_transform_inverse.argtypes = [vp, i32, vp, i32, vp, i32, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.knz_entropy_encode.argtypes = [vp, i32, vp, i32, vp, i64, ctypes.POINTER(i64)]
    L.knz_entropy_decode.argtypes = [vp, i32, vp, i64, vp, i32]
    L.knz_dist_unique_id.argtypes = [vp]
    L.knz_dist_init.argtypes = [vp, i32, i32, vp]
    L.knz_dist_init_transport.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    L.knz_compress_dist.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, i32, vp, i64, vp, i64, ctypes.POINTER(i64)]
    L.knz_decompress_dist.argtypes = [vp, vp, i64, vp, i64, ctypes.POINTER(i64)]
    L.knz_dist_encode_dev.argtypes = [vp, u64, i32, i32, vp, i64, vp, i32, i32, i32, vp, i64, u64, vp,
                                      ctypes.POINTER(u64)]
    L.knz_dist_decode_dev.argtypes = [vp, u64, i32, i32, vp, i64, u64, vp, i32, vp, i64, vp]
    return L


class Context:
    """One GPU context (device buffers, stream, kernels).  `lib_path` is only
    overridden by the CPU tests that load the emulator build of the same sources."""

    def __init__(self, device=0, max_block_size=4 << 20, max_batch_blocks=64, lib_path=None):
        self.lib = _load(lib_path or LIB_PATH)
        h = ctypes.c_void_p()
        rc = self.lib.knz_create(device, max_block_size, max_batch_blocks, ctypes.byref(h))
        if rc != 0:
            raise KanziGpuError(rc, "knz_create failed (no usable CUDA device? there is no CPU fallback)")
        self.h = h
        self.max_block_size = max_block_size
        self.max_batch_blocks = max_batch_blocks

    def close(self):
        if getattr(self, "h", None):
            self.lib.knz_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise KanziGpuError(rc, self.lib.knz_last_error(self.h).decode())

    @property
    def launches(self):
        return int(self.lib.knz_launch_count(self.h))

    @property
    def cuda_stream(self):
        return int(self.lib.knz_stream(self.h) or 0)

    def set_checksum(self, bits):
        """Block checksums written by the encoders of this context: 0, 32 (XXHash32) or 64 (XXHash64)."""
        self._check(self.lib.knz_set_checksum(self.h, int(bits)))

    def set_skip_blocks(self, on):
        """`skipBlocks` of the reference's context: incompressible blocks are stored as copy blocks."""
        self._check(self.lib.knz_set_skip_blocks(self.h, 1 if on else 0))

    def set_listener(self, fn):
        """fn(dict) is called with every per-block event (type, blockId, size, hash, hashBits, offset,
        skipFlags) after each device batch; None switches the events off."""
        if fn is None:
            self._listener = None
            self._check(self.lib.knz_set_listener(self.h, None, None))
            return

        def tramp(_user, evt):
            e = evt.contents
            fn({"type": EVENT_NAMES.get(e.type, e.type), "blockId": e.blockId, "size": e.size, "hash": e.hash,
                "hashBits": e.hashBits, "offset": e.offset, "skipFlags": e.skipFlags})

        self._listener = _EVENT_FN(tramp)  # keep the trampoline alive
        self._check(self.lib.knz_set_listener(self.h, ctypes.cast(self._listener, ctypes.c_void_p), None))

    def set_decode_groups(self, groups):
        """Block groups decoded concurrently (1 = serial stages with per-stage timings)."""
        self._check(self.lib.knz_set_decode_groups(self.h, int(groups)))

    def timings(self):
        ms = (ctypes.c_float * 8)()
        self.lib.knz_last_timings(self.h, ms)
        return dict(zip(("bwt", "rank", "zrlt", "entropy", "assembly", "total", "ans_enc_kernel", "ans_dec_kernel"),
                        [float(x) for x in ms]))

    def transform_type(self, name):
        return int(self.lib.knz_transform_type(name.encode()))

    # ---- stream level (CompressedOutputStream write+close / CompressedInputStream read)
    def compress(self, data, transform="BWT+RANK+ZRLT", entropy="ANS0", block_size=4 << 20, out=None):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        # worst case: every block expanded 2x by a transform stage, 25 % by the entropy stage, plus the
        # ANS1 context headers (see knz_compress); untouched pages of the buffer cost nothing
        nblk = data.size // block_size + 1
        cap = 2 * data.size + data.size // 2 + nblk * (1024 + (131072 * (block_size // (4 << 20) + 1) if entropy == "ANS1" else 0)) + 65536
        if out is None:
            out = np.empty(cap, dtype=np.uint8)
        n = ctypes.c_int64(0)
        self._check(self.lib.knz_compress(self.h, transform.encode(), entropy.encode(), block_size, _ptr(data),
                                          data.size, _ptr(out), out.size, ctypes.byref(n)))
        return out[: n.value]

    def decompress(self, comp, cap, out=None):
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        if out is None:
            out = np.empty(max(cap, 1), dtype=np.uint8)
        n = ctypes.c_int64(0)
        self._check(self.lib.knz_decompress(self.h, _ptr(comp), comp.size, _ptr(out), cap, ctypes.byref(n)))
        return out[: n.value]

    def decompress_range(self, comp, from_block, to_block, cap, out=None):
        """Blocks with from_block <= id < to_block (1-based ids), like the reference's `from` / `to`."""
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        if out is None:
            out = np.empty(max(cap, 1), dtype=np.uint8)
        n = ctypes.c_int64(0)
        self._check(self.lib.knz_decompress_range(self.h, _ptr(comp), comp.size, int(from_block), int(to_block),
                                                  _ptr(out), cap, ctypes.byref(n)))
        return out[: n.value]

    def decompress_seek(self, comp, bit_pos, n_blocks, cap, out=None):
        """Up to n_blocks blocks starting at the block boundary at bit `bit_pos` (the offset of a BLOCK_INFO event)."""
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        if out is None:
            out = np.empty(max(cap, 1), dtype=np.uint8)
        n = ctypes.c_int64(0)
        self._check(self.lib.knz_decompress_seek(self.h, _ptr(comp), comp.size, int(bit_pos), int(n_blocks), _ptr(out), cap,
                                                 ctypes.byref(n)))
        return out[: n.value]

    # ---- multi-GPU: one process per GPU, blocks sharded round-robin (see include/knz_gpu.h)
    def dist_init(self, rank, world):
        """Collective.  With the NCCL backend the library opens its own communicator (the 128-byte
        unique id travels over torch.distributed); with gloo (CPU tests on the emulator build) the
        three collectives the library needs are served by torch.distributed through callbacks."""
        self.rank, self.world = rank, world
        buf = np.zeros(128, dtype=np.uint8)
        if world == 1:
            self._check(self.lib.knz_dist_init(self.h, 0, 1, _ptr(buf)))
            return
        import torch
        import torch.distributed as dist
        if dist.get_backend() == "nccl":
            if rank == 0:
                self._check(self.lib.knz_dist_unique_id(_ptr(buf)))
            t = torch.from_numpy(buf).cuda()
            dist.broadcast(t, 0)
            buf = t.cpu().numpy()
            self._check(self.lib.knz_dist_init(self.h, rank, world, _ptr(buf)))
            return

        def view(ptr, n):
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(max(int(n), 1),))[: int(n)]

        def allgather(user, send, nbytes, recv):
            ts = torch.from_numpy(view(send, nbytes).copy())
            outs = [torch.empty_like(ts) for _ in range(world)]
            dist.all_gather(outs, ts)
            view(recv, world * nbytes)[:] = torch.cat(outs).numpy()
            return 0

        def gather(user, send, nbytes, recv):
            ts = torch.from_numpy(view(send, nbytes).copy())
            if rank == 0:
                outs = [torch.empty_like(ts) for _ in range(world)]
                dist.gather(ts, outs, dst=0)
                view(recv, world * nbytes)[:] = torch.cat(outs).numpy()
            else:
                dist.gather(ts, None, dst=0)
            return 0

        def bcast(user, buf_, nbytes):
            ts = torch.from_numpy(view(buf_, nbytes).copy())
            dist.broadcast(ts, 0)
            view(buf_, nbytes)[:] = ts.numpy()
            return 0

        ft = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)
        fb = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64)
        self._cbs = (ft(allgather), ft(gather), fb(bcast))  # keep the trampolines alive
        self._check(self.lib.knz_dist_init_transport(self.h, rank, world, ctypes.cast(self._cbs[0], ctypes.c_void_p),
                                                     ctypes.cast(self._cbs[1], ctypes.c_void_p),
                                                     ctypes.cast(self._cbs[2], ctypes.c_void_p), None))

    def compress_dist(self, data, transform="BWT+RANK+ZRLT", entropy="ANS0", block_size=4 << 20, out=None):
        """Collective: every rank passes the whole input; rank 0 gets the stream (others an empty array)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        nblk = data.size // block_size + 1
        cap = 2 * data.size + data.size // 2 + nblk * (1024 + (131072 * (block_size // (4 << 20) + 1) if entropy == "ANS1" else 0)) + 65536
        if out is None:
            out = np.empty(cap if getattr(self, "rank", 0) == 0 else 16, dtype=np.uint8)
        n = ctypes.c_int64(0)
        self._check(self.lib.knz_compress_dist(self.h, transform.encode(), entropy.encode(), block_size, _ptr(data),
                                               data.size, _ptr(out), out.size, ctypes.byref(n)))
        return out[: n.value]

    def decompress_dist(self, comp, cap, out=None):
        """Collective: every rank passes the whole stream; each rank fills the blocks it owns in `out`."""
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        if out is None:
            out = np.zeros(max(cap, 1), dtype=np.uint8)
        n = ctypes.c_int64(0)
        self._check(self.lib.knz_decompress_dist(self.h, _ptr(comp), comp.size, _ptr(out), cap, ctypes.byref(n)))
        return out[: n.value]

    # ---- block level (EncodingTask::run / DecodingTask::run bodies)
    def encode_blocks(self, blocks, transform, entropy, block_size, first_block_len=None):
        """blocks: list of uint8 arrays.  Returns [(bytes, nbits, skipFlags)]."""
        nb = len(blocks)
        stride = (max(b.size for b in blocks) + 255) // 256 * 256
        inp = np.zeros(nb * stride, dtype=np.uint8)
        lens = np.zeros(nb, dtype=np.int32)
        for i, b in enumerate(blocks):
            inp[i * stride: i * stride + b.size] = b
            lens[i] = b.size
        ostride = (block_size + block_size // 4 + 4096 + 131072 * (block_size // (4 << 20) + 1) + 255) // 256 * 256
        out = np.zeros(nb * ostride, dtype=np.uint8)
        bits = np.zeros(nb, dtype=np.uint64)
        flags = np.zeros(nb, dtype=np.uint8)
        first = int(lens[0]) if first_block_len is None else first_block_len
        self._check(self.lib.knz_encode_blocks(self.h, self.transform_type(transform), E_IDS[entropy], block_size,
                                               _ptr(inp), stride, _ptr(lens), nb, first, _ptr(out), ostride,
                                               _ptr(bits), _ptr(flags)))
        res = []
        for i in range(nb):
            nbytes = (int(bits[i]) + 7) // 8
            res.append((out[i * ostride: i * ostride + nbytes].copy(), int(bits[i]), int(flags[i])))
        return res

    def decode_blocks(self, enc, transform, entropy, block_size):
        """enc: list of (bytes, nbits).  Returns list of uint8 arrays."""
        nb = len(enc)
        stride = (max(e[0].size for e in enc) + 64 + 255) // 256 * 256
        inp = np.zeros(nb * stride, dtype=np.uint8)
        bits = np.zeros(nb, dtype=np.uint64)
        for i, (e, nbit) in enumerate(enc):
            inp[i * stride: i * stride + e.size] = e
            bits[i] = nbit
        ostride = (block_size + 255) // 256 * 256
        out = np.zeros(nb * ostride, dtype=np.uint8)
        lens = np.zeros(nb, dtype=np.int32)
        self._check(self.lib.knz_decode_blocks(self.h, self.transform_type(transform), E_IDS[entropy], block_size,
                                               _ptr(inp), stride, _ptr(bits), nb, _ptr(out), ostride, _ptr(lens)))
        return [out[i * ostride: i * ostride + int(lens[i])].copy() for i in range(nb)]

    # ---- stage level (Transform<byte>::forward/inverse, EntropyEncoder::encode, EntropyDecoder::decode)
    def transform_forward(self, name, data, cap=None):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        cap = data.size + (1088 if name == "SRT" else 64) if cap is None else cap
        out = np.zeros(cap + 64, dtype=np.uint8)
        ol, ap = ctypes.c_int(0), ctypes.c_int(0)
        self._check(self.lib.knz_transform_forward(self.h, T_IDS[name], _ptr(data), data.size, _ptr(out), cap,
                                                   ctypes.byref(ol), ctypes.byref(ap)))
        return out[: ol.value].copy(), bool(ap.value)

    def transform_inverse(self, name, data, cap):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(cap + 64, dtype=np.uint8)
        ol, ap = ctypes.c_int(0), ctypes.c_int(0)
        self._check(self.lib.knz_transform_inverse(self.h, T_IDS[name], _ptr(data), data.size, _ptr(out), cap,
                                                   ctypes.byref(ol), ctypes.byref(ap)))
        return out[: ol.value].copy(), bool(ap.value)

    def entropy_encode(self, name, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(data.size + data.size // 4 + 8192 + (140000 if name == "ANS1" else 0), dtype=np.uint8)
        bits = ctypes.c_int64(0)
        self._check(self.lib.knz_entropy_encode(self.h, E_IDS[name], _ptr(data), data.size, _ptr(out), out.size,
                                                ctypes.byref(bits)))
        return out[: (bits.value + 7) // 8].copy(), bits.value

    def entropy_decode(self, name, enc, nbits, n):
        enc = np.ascontiguousarray(enc, dtype=np.uint8)
        out = np.zeros(max(n, 1), dtype=np.uint8)
        self._check(self.lib.knz_entropy_decode(self.h, E_IDS[name], _ptr(enc), nbits, _ptr(out), n))
        return out[:n]


class Compressor:
    """Mirror of the reference binding's Compressor (src/api/kanzi.py:18): init with
    the stream parameters, then compress() whole buffers."""

    def __init__(self, transform="BWT+RANK+ZRLT", entropy="ANS0", block_size=4 << 20, device=0, max_batch_blocks=64,
                 ctx=None):
        self.transform, self.entropy, self.block_size = transform, entropy, block_size
        self.ctx = ctx or Context(device, block_size, max_batch_blocks)

    def compress(self, data):
        return self.ctx.compress(data, self.transform, self.entropy, self.block_size)


class Decompressor:
    """Mirror of src/api/kanzi.py:86."""

    def __init__(self, max_block_size=4 << 20, device=0, max_batch_blocks=64, ctx=None):
        self.ctx = ctx or Context(device, max_block_size, max_batch_blocks)

    def decompress(self, comp, original_size):
        return self.ctx.decompress(comp, original_size)
