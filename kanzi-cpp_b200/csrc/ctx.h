// ctx.h -- private definition of the context behind the C ABI (include/knz_gpu.h), shared by
// api.cu (single-GPU entry points) and dist.cu (one-process-per-GPU sharding).
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "../../include/knz_gpu.h"
#include "kernels.h"

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d CUDA error %d: %s", __FILE__, __LINE__,    \
                     (int)e_, cudaGetErrorString(e_));                                             \
            return KNZ_ERR_PROCESS_BLOCK;                                                          \
        }                                                                                          \
    } while (0)

#define KNZ_MAX_GROUPS 8

struct KnzDist; // dist.cu

struct knz_ctx {
    int device, maxBlockSize, maxBatch;
    cudaStream_t stream;
    cudaStream_t copyStream; // host->device staging of the stream-level API (overlaps with compute)
    cudaStream_t d2hStream;  // device->host copies of finished output (second DMA direction)
    cudaEvent_t evCopy[2], evDone[2];
    cudaEvent_t ev[10];
    cudaEvent_t evStage[16]; // per-stage brackets: recorded during a batch, read once after it (no mid-batch sync)
    std::recursive_mutex mtx; // every C-ABI entry point locks its context (reference worker threads share one)
    int stageCap;             // bytes a stage buffer slot can hold (bstride - 64)
    int encSched[3];          // knz_compress sub-batch schedule (env KNZ_ENC_BATCH)
    int decBwtGroups;         // knz_decompress: groups of the last inverse-BWT stage (env KNZ_DEC_GROUPS)
    cudaStream_t gStream[KNZ_MAX_GROUPS]; // decode: one stream per block group (stages of different groups overlap)
    cudaEvent_t gEv[KNZ_MAX_GROUPS + 1];
    int decGroups;                        // 1 = one group, per-stage timings valid
    i64 bstride;     // stride of the ping-pong stage buffers
    u8 *bufA, *bufB; // [maxBatch * bstride]
    u8* dStageIn;    // host API: staged input blocks [maxBatch * bstride]
    u8* dOut;        // per-block output bit strings [maxBatch * outStride]
    i64 outStride;
    BlkState* st;    // [10][maxBatch]
    int *capEven, *capOdd;
    int maxChunks;
    u8* slots;
    u32 *hdrBits, *payBytes, *payOff;
    u64 *chunkOff, *blockBits, *blockOff, *streamPos;
    u64 *chunkPos, *dInBits, *dPayStart;
    int* dPreLen;
    int* errFlag;
    Workspace ws;
    Ans1Work a1; // order-1 rANS scratch, allocated on the first use of ANS1
    bool a1Ready;
    SrtWork srt; // SRT scratch, allocated on the first SRT stage
    bool srtReady;
    LzWork lz; // LZ / LZX / LZP scratch, allocated on the first stage of that family
    bool lzReady;
    // pinned host mirrors
    BlkState* h_st;
    int *h_capEven, *h_capOdd, *h_err, *h_preLen;
    u64 *h_bits, *h_payStart, *h_pos;
    // stream-level buffers (grown on demand)
    u8* dStream;
    i64 dStreamCap;
    u8* dPlain;
    i64 dPlainCap;
    u8* dPlain2;
    i64 dPlain2Cap;
    int checksumBits;  // 0, 32 or 64: block checksums written by the encoders of this context (knz_set_checksum)
    u64 *blockHash, *expectHash; // [maxBatch] device: hashes of a batch / values read from the block headers
    u64* h_hash;       // pinned mirror
    int skipBlocks;    // knz_set_skip_blocks: store incompressible blocks as copy blocks
    int *dSkip, *h_skip, *dLog2Tab; // [maxBatch] decisions of a batch (device, pinned mirror); log2 table
    knz_event_fn listener; // knz_set_listener: per-block events re-emitted after each batch
    void* listenerUser;
    int evBlockBase;       // stream-level calls: id of the first block of the current batch minus one
    // host stages in front of the device stages (pre.cu): set by the stream-level entry points around a batch
    int nHost;             // leading stages of the sequence already applied on the host (0: none)
    BlkState* h_init;      // [maxBatch] state of every block behind the host stages (length, swaps, skip flags)
    int *h_dtype, *dDtype; // [maxBatch] Global::DataType of every block as the host stages left it
    u8* h_pre;             // pinned staging of the blocks behind the host stages [maxBatch * bstride], on demand
    KnzDist* dist; // multi-process sharding state (dist.cu), NULL until knz_dist_init*
    u64 launches;
    float ms[8];
    char err[256];
};


// MSB-first reader over host bytes (stream header, block prefixes, block headers)
struct HostBitReader {
    const u8* p;
    u64 nbits, pos;
    bool bad;
    u64 get(int n)
    {
        u64 v = 0;
        for (int k = 0; k < n; k++) {
            u64 b = 0;
            if (pos < nbits)
                b = (p[pos >> 3] >> (7 - (pos & 7))) & 1;
            else
                bad = true;
            v = (v << 1) | b;
            pos++;
        }
        return v;
    }
};

// internals of api.cu used by dist.cu
i64 knz_round_up(i64 v, i64 a);
int knz_grow(knz_ctx* ctx, u8** buf, i64* cap, i64 need);
int knz_encode_batch(knz_ctx* ctx, u64 tType, int eType, int blockSize, const u8* d_in, i64 inStride,
                     const int32_t* lens, int nB, int firstBlockLen, u8* d_out, i64 outStride, u64* d_bits,
                     u8* h_flags);
// h_expectHash (may be NULL): the checksums read from the block headers, verified after the inverse transforms
int knz_decode_batch(knz_ctx* ctx, u64 tType, int eType, int blockSize, const u8* d_in, i64 inStride,
                     const u64* h_payStart, const u64* h_endBit, const int* h_preLen, const u8* h_flags, int nB,
                     u8* d_out, i64 outStride, int32_t* h_outLens, u8* h_sink = NULL, int* h_sinkBlocks = NULL,
                     const u64* h_expectHash = NULL, int ckBits = 0);
// Parse one block's private header (mode byte, [skip flags], length, [checksum of ckBits bits]) at r.pos:
// 0 ok, 1 copy block, <0 error.
int knz_parse_block_header(HostBitReader& r, int blockSize, u8* flags, int* preLen, int ckBits = 0, u64* checksum = NULL);
u64 knz_frame_small_block(const u8* in, int len, u8* out, int ckBits = 0);
// Stream header fields (io/CompressedInputStream.cpp:511-663); returns KNZ_OK and leaves r behind the header.
struct KnzStreamInfo {
    int eType, blockSize, ckBits;
    u64 tType;
    i64 origSize; // -1 when the header does not carry it
};
int knz_parse_stream_header(knz_ctx* ctx, HostBitReader& r, KnzStreamInfo* info);
void knz_dist_destroy(knz_ctx* ctx);
// host stages around a device batch (api.cu; pre.cu holds the stages)
int knz_split_types(u64 tType, int* types);
int knz_host_prefix_len(const int* types, int nt); // leading host stages, -1 if one follows a device stage
int knz_ensure_h_pre(knz_ctx* ctx);
int knz_host_prefix_forward(knz_ctx* ctx, const int* types, int hs, int eType, int blockSize, const u8* in, i64 inStride,
                            const int32_t* lens, int nb);
void knz_host_prefix_inverse(knz_ctx* ctx, const int* types, int hs, int eType, int blockSize, const u8* flags,
                             const int* lens, int nb, u8* const* out, const int* outCap, int* outLen, u8* ok);
