"""Thin helpers around the device-resident block entry points (used by the profiling probes under
tools/probes).  The multi-GPU data path itself -- round-robin ownership, the all-gather of bit
counts, the NCCL gather of payloads, the stream assembly on rank 0 and the prefix walk on decode --
lives behind the C ABI in csrc/dist.cu (knz_dist_*, knz_compress_dist, knz_decompress_dist); Python
only calls it (Context.dist_init / compress_dist / decompress_dist)."""
import ctypes

import numpy as np


def shard_blocks(nblocks, rank, world):
    """Indices of the blocks owned by `rank` (block i -> rank i % world)."""
    return list(range(rank, nblocks, world))


def encode_shard(ctx, ttype, etype, block_size, d_in, lens, first_block_len, d_blk, d_bits):
    """knz_encode_blocks_dev on torch device tensors."""
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    rc = ctx.lib.knz_encode_blocks_dev(ctx.h, ttype, etype, block_size, d_in.data_ptr(), d_in.shape[1],
                                       lens.ctypes.data_as(ctypes.c_void_p), len(lens), first_block_len,
                                       d_blk.data_ptr(), d_blk.shape[1], d_bits.data_ptr(), None)
    if rc != 0:
        raise RuntimeError(f"knz_encode_blocks_dev: {rc} {ctx.lib.knz_last_error(ctx.h).decode()}")


def decode_shard(ctx, ttype, etype, block_size, d_blk, bits_host, d_out):
    """knz_decode_blocks_dev on torch device tensors; returns the decoded lengths."""
    bits_host = np.ascontiguousarray(bits_host, dtype=np.uint64)
    out_lens = np.zeros(len(bits_host), dtype=np.int32)
    rc = ctx.lib.knz_decode_blocks_dev(ctx.h, ttype, etype, block_size, d_blk.data_ptr(), d_blk.shape[1],
                                       bits_host.ctypes.data_as(ctypes.c_void_p), len(bits_host), d_out.data_ptr(),
                                       d_out.shape[1], out_lens.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(f"knz_decode_blocks_dev: {rc} {ctx.lib.knz_last_error(ctx.h).decode()}")
    return out_lens
