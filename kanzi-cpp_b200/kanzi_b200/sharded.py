"""Multi-GPU block sharding for the kanzi block pipeline (one process per GPU).

Blocks are independent (src/Transform.hpp:27-29), so block i goes to rank i % world
(host-side round-robin, no data-path collective).  The only exchange step is
gathering the compressed blocks back into stream order on rank 0, where the
bit-concatenation kernel lays them down at their bit offsets
(io/CompressedOutputStream.cpp:852-864 did this under a mutex, block after block).

The code is backend-agnostic: with NCCL the tensors live in HBM; the CPU tests run
the very same functions under gloo against the emulator build of the library.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist


def shard_blocks(nblocks, rank, world):
    """Indices of the blocks owned by `rank`."""
    return list(range(rank, nblocks, world))


def gather_blocks(blk, bits, rank, world):
    """blk: [nb, stride] uint8 block bit strings of this rank, bits: [nb] int64.
    Returns on rank 0 (ordered_blocks [nblocks, max_bytes], ordered_bits [nblocks]) in
    stream order; None on the other ranks.  Requires the same nb on every rank."""
    if world == 1:
        return blk, bits
    bits_all = [torch.empty_like(bits) for _ in range(world)]
    dist.all_gather(bits_all, bits)
    all_bits = torch.stack(bits_all, 1).reshape(-1).contiguous()  # block i = rank i%world, slot i//world
    max_bytes = (int(all_bits.max().item()) + 7) // 8
    max_bytes = min((max_bytes + 255) // 256 * 256, blk.shape[1])
    mine = blk[:, :max_bytes].contiguous()
    if rank == 0:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.gather(mine, parts, dst=0)
        ordered = torch.stack(parts, 1).reshape(-1, max_bytes).contiguous()
        return ordered, all_bits
    dist.gather(mine, None, dst=0)
    return None


def assemble_stream(ctx, blk, bits, d_stream, start_bit):
    """Bit-concatenate ordered blocks into d_stream (zeroed by the caller) starting at
    start_bit; returns the end bit position.  Batches chain through the library."""
    lib = ctx.lib
    total, stride = blk.shape[0], blk.shape[1]
    pos = ctypes.c_uint64(start_bit)
    for off in range(0, total, ctx.max_batch_blocks):
        n = min(ctx.max_batch_blocks, total - off)
        rc = lib.knz_assemble_stream_dev(ctx.h, blk.data_ptr() + off * stride, stride, bits.data_ptr() + 8 * off, n,
                                         d_stream.data_ptr(), d_stream.numel(), pos, ctypes.byref(pos))
        if rc != 0:
            raise RuntimeError(f"knz_assemble_stream_dev: {rc} {lib.knz_last_error(ctx.h).decode()}")
    return pos.value


def encode_shard(ctx, ttype, etype, block_size, d_in, lens, first_block_len, d_blk, d_bits):
    """Forward transforms + entropy coding of this rank's blocks (device resident)."""
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    rc = ctx.lib.knz_encode_blocks_dev(ctx.h, ttype, etype, block_size, d_in.data_ptr(), d_in.shape[1],
                                       lens.ctypes.data_as(ctypes.c_void_p), len(lens), first_block_len,
                                       d_blk.data_ptr(), d_blk.shape[1], d_bits.data_ptr(), None)
    if rc != 0:
        raise RuntimeError(f"knz_encode_blocks_dev: {rc} {ctx.lib.knz_last_error(ctx.h).decode()}")


def decode_shard(ctx, ttype, etype, block_size, d_blk, bits_host, d_out):
    """Entropy decoding + inverse transforms of this rank's blocks; returns decoded lengths."""
    bits_host = np.ascontiguousarray(bits_host, dtype=np.uint64)
    out_lens = np.zeros(len(bits_host), dtype=np.int32)
    rc = ctx.lib.knz_decode_blocks_dev(ctx.h, ttype, etype, block_size, d_blk.data_ptr(), d_blk.shape[1],
                                       bits_host.ctypes.data_as(ctypes.c_void_p), len(bits_host), d_out.data_ptr(),
                                       d_out.shape[1], out_lens.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(f"knz_decode_blocks_dev: {rc} {ctx.lib.knz_last_error(ctx.h).decode()}")
    return out_lens
