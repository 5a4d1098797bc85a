// dist.cu -- block sharding across GPUs, one process per GPU (include/knz_gpu.h, "multi-GPU").
//
// Blocks are independent (src/Transform.hpp:27-29; a fresh sequence and codec per block,
// io/CompressedOutputStream.cpp:719,818), so block i belongs to rank i % world: host-side
// round-robin, no data-path collective while the transforms and the entropy coders run.  The one
// exchange step is the ordered append of the block bit strings to the shared bitstream, which the
// reference serialises with _processedBlockId + a condition variable
// (io/CompressedOutputStream.cpp:836-868).  Here:
//   encode   all-gather of the per-block bit counts (8 B per block), gather of the block payloads to
//            rank 0 over NCCL (NVLink / NVSwitch), one bit-concatenation kernel on rank 0 that lays
//            every block down at its bit offset (blocks arrive grouped by rank; the kernel reads them
//            through an index), one device -> host copy of the finished stream
//   decode   the host walks the 5 + lw bit length prefixes (io/CompressedInputStream.cpp:823-856) and
//            ships only the bit ranges of its own blocks to its GPU; decoded blocks are byte aligned
//            and fixed size, so they go straight to their place in the output (no collective)
// The transport is NCCL (knz_dist_init) or three caller-supplied callbacks (knz_dist_init_transport;
// the CPU tests run this file on the emulator with gloo underneath).
#include "ctx.h"

#ifndef KNZ_SIM
#include <dlfcn.h>
#include <nccl.h>

// NCCL is bound at knz_dist_init time, not at link time: a process that also runs torch.distributed must
// use the ONE libnccl.so.2 already mapped (the torch-bundled build), whichever library was loaded first;
// a plain C++ host gets the system library.  Only the eight entry points below are used.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    bool ok;
};

static const NcclApi* nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); // already mapped (torch): share it
        if (h == NULL)
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (h == NULL)
            h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h == NULL)
            return;
        bool ok = true;
#define NCCL_SYM(field, name)                                  \
    *(void**)(&api.field) = dlsym(h, name);                    \
    ok = ok && (*(void**)(&api.field) != NULL)
        NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        NCCL_SYM(CommInitRank, "ncclCommInitRank");
        NCCL_SYM(CommDestroy, "ncclCommDestroy");
        NCCL_SYM(AllGather, "ncclAllGather");
        NCCL_SYM(Broadcast, "ncclBroadcast");
        NCCL_SYM(Send, "ncclSend");
        NCCL_SYM(Recv, "ncclRecv");
        NCCL_SYM(GroupStart, "ncclGroupStart");
        NCCL_SYM(GroupEnd, "ncclGroupEnd");
#undef NCCL_SYM
        api.ok = ok;
    });
    return api.ok ? &api : NULL;
}
#endif

struct KnzDist {
    int rank, world;
    bool useNccl;
#ifndef KNZ_SIM
    ncclComm_t comm;
#endif
    knz_allgather_fn ag;
    knz_gather_fn ga;
    knz_bcast_fn bc;
    void* user;
    u8* dIn;      // own input blocks (host API) / own compressed ranges (decode)
    i64 dInCap;
    u8* dBlk;     // own blocks' bit strings [nbMax][outStride]
    i64 dBlkCap;
    u8* dPack;    // own blocks repacked at the gather stride
    i64 dPackCap;
    u8* dGather;  // rank 0: [world][nbMax][gather stride]
    i64 dGatherCap;
    u8* dMeta;    // own bits | all bits | ordered bits | block offsets | source index
    i64 dMetaCap;
    u8* dPlain;   // decoded own blocks (host API)
    i64 dPlainCap;
    cudaEvent_t evBatch[64];
    int nEvBatch;
};

#define DCK(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d CUDA error %d: %s", __FILE__, __LINE__,   \
                     (int)e_, cudaGetErrorString(e_));                                            \
            return KNZ_ERR_PROCESS_BLOCK;                                                         \
        }                                                                                         \
    } while (0)

void knz_dist_destroy(knz_ctx* ctx)
{
    KnzDist* D = ctx->dist;
    if (D == NULL)
        return;
    void* dev[] = { D->dIn, D->dBlk, D->dPack, D->dGather, D->dMeta, D->dPlain };
    for (size_t i = 0; i < sizeof(dev) / sizeof(dev[0]); i++)
        if (dev[i])
            cudaFree(dev[i]);
    for (int i = 0; i < D->nEvBatch; i++)
        cudaEventDestroy(D->evBatch[i]);
#ifndef KNZ_SIM
    if (D->useNccl)
        nccl_api()->CommDestroy(D->comm);
#endif
    delete D;
    ctx->dist = NULL;
}

static int dist_new(knz_ctx* ctx, int rank, int world)
{
    if (ctx == NULL || world < 1 || rank < 0 || rank >= world)
        return KNZ_ERR_INVALID_PARAM;
    knz_dist_destroy(ctx);
    KnzDist* D = new (std::nothrow) KnzDist();
    if (D == NULL)
        return KNZ_ERR_CREATE_COMPRESSOR;
    D->rank = rank;
    D->world = world;
    ctx->dist = D;
    return KNZ_OK;
}

extern "C" int knz_dist_unique_id(uint8_t id[128])
{
    if (id == NULL)
        return KNZ_ERR_INVALID_PARAM;
    memset(id, 0, 128);
#ifndef KNZ_SIM
    ncclUniqueId u;
    const NcclApi* N = nccl_api();
    if (N == NULL || N->GetUniqueId(&u) != ncclSuccess)
        return KNZ_ERR_CREATE_COMPRESSOR;
    static_assert(sizeof(ncclUniqueId) <= 128, "unique id does not fit");
    memcpy(id, &u, sizeof(u));
    return KNZ_OK;
#else
    return KNZ_ERR_CREATE_COMPRESSOR; // the emulator build has no NCCL: use knz_dist_init_transport
#endif
}

extern "C" int knz_dist_init(knz_ctx* ctx, int rank, int world, const uint8_t id[128])
{
    if (ctx == NULL || id == NULL)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    const int rc = dist_new(ctx, rank, world);
    if (rc != KNZ_OK)
        return rc;
#ifndef KNZ_SIM
    if (world > 1) {
        cudaSetDevice(ctx->device);
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        const NcclApi* N = nccl_api();
        if (N == NULL || N->CommInitRank(&ctx->dist->comm, world, u, rank) != ncclSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), N ? "ncclCommInitRank failed" : "libnccl.so.2 not found");
            delete ctx->dist;
            ctx->dist = NULL;
            return KNZ_ERR_CREATE_COMPRESSOR;
        }
        ctx->dist->useNccl = true;
    }
    return KNZ_OK;
#else
    return (world == 1) ? KNZ_OK : KNZ_ERR_CREATE_COMPRESSOR;
#endif
}

extern "C" int knz_dist_init_transport(knz_ctx* ctx, int rank, int world, knz_allgather_fn allgather,
                                       knz_gather_fn gather, knz_bcast_fn bcast, void* user)
{
    if (ctx == NULL || (world > 1 && (!allgather || !gather || !bcast)))
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    const int rc = dist_new(ctx, rank, world);
    if (rc != KNZ_OK)
        return rc;
    ctx->dist->ag = allgather;
    ctx->dist->ga = gather;
    ctx->dist->bc = bcast;
    ctx->dist->user = user;
    return KNZ_OK;
}

// ---- transport: every rank calls these in the same order; buffers are device memory
static int t_allgather(knz_ctx* ctx, const void* send, i64 bytes, void* recv)
{
    KnzDist* D = ctx->dist;
    cudaStream_t s = ctx->stream;
    if (D->world == 1) {
        DCK(cudaMemcpyAsync(recv, send, (size_t)bytes, cudaMemcpyDeviceToDevice, s));
        return KNZ_OK;
    }
#ifndef KNZ_SIM
    if (D->useNccl) {
        if (nccl_api()->AllGather(send, recv, (size_t)bytes, ncclUint8, D->comm, s) != ncclSuccess)
            return KNZ_ERR_PROCESS_BLOCK;
        return KNZ_OK;
    }
#endif
    DCK(cudaStreamSynchronize(s));
    return D->ag(D->user, send, bytes, recv) == 0 ? KNZ_OK : KNZ_ERR_PROCESS_BLOCK;
}

static int t_gather(knz_ctx* ctx, const void* send, i64 bytes, void* recv)
{
    KnzDist* D = ctx->dist;
    cudaStream_t s = ctx->stream;
    if (D->world == 1) {
        DCK(cudaMemcpyAsync(recv, send, (size_t)bytes, cudaMemcpyDeviceToDevice, s));
        return KNZ_OK;
    }
#ifndef KNZ_SIM
    if (D->useNccl) {
        const NcclApi* N = nccl_api();
        bool ok = N->GroupStart() == ncclSuccess;
        if (D->rank == 0)
            for (int r = 0; r < D->world; r++)
                ok = ok && N->Recv((u8*)recv + (i64)r * bytes, (size_t)bytes, ncclUint8, r, D->comm, s) == ncclSuccess;
        ok = ok && N->Send(send, (size_t)bytes, ncclUint8, 0, D->comm, s) == ncclSuccess;
        ok = (N->GroupEnd() == ncclSuccess) && ok;
        return ok ? KNZ_OK : KNZ_ERR_PROCESS_BLOCK;
    }
#endif
    DCK(cudaStreamSynchronize(s));
    return D->ga(D->user, send, bytes, recv) == 0 ? KNZ_OK : KNZ_ERR_PROCESS_BLOCK;
}

static int t_bcast(knz_ctx* ctx, void* buf, i64 bytes)
{
    KnzDist* D = ctx->dist;
    cudaStream_t s = ctx->stream;
    if (D->world == 1)
        return KNZ_OK;
#ifndef KNZ_SIM
    if (D->useNccl)
        return nccl_api()->Broadcast(buf, buf, (size_t)bytes, ncclUint8, 0, D->comm, s) == ncclSuccess ? KNZ_OK
                                                                                                : KNZ_ERR_PROCESS_BLOCK;
#endif
    DCK(cudaStreamSynchronize(s));
    return D->bc(D->user, buf, bytes) == 0 ? KNZ_OK : KNZ_ERR_PROCESS_BLOCK;
}

static int own_count(int nBlocks, int rank, int world) { return (nBlocks > rank) ? (nBlocks - rank + world - 1) / world : 0; }

static u32 prefix_bits(u64 w) // lw of a block of w bits (io/CompressedOutputStream.cpp:833)
{
    u32 lw = 3;
    while (lw < 35 && (w >> lw) != 0)
        lw++;
    return lw;
}

// Copy nbytes bytes that start at an arbitrary bit position of a device bit string.
__global__ void dist_copy_bits_kernel(const u8* __restrict__ src, u64 bitPos, int nbytes, u8* __restrict__ dst)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += gridDim.x * blockDim.x) {
        const u64 p = bitPos + 8ull * (u64)i;
        const u32 w = ((u32)src[p >> 3] << 8) | (u32)src[(p >> 3) + 1];
        dst[i] = (u8)(w >> (8 - (int)(p & 7)));
    }
}

// 16 bytes of every listed block, starting at the byte that holds its first bit.
__global__ void dist_heads_kernel(const u8* __restrict__ stream, const u64* __restrict__ startBit, int n,
                                  u8* __restrict__ heads)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 16)
        return;
    heads[i] = stream[(startBit[i >> 4] >> 3) + (u64)(i & 15)];
}

// Encode this rank's blocks (device resident), gather every rank's blocks on rank 0 and assemble
// the stream body there.  h_allBits (optional, every rank): bit count of every block in stream order.
// Leading host stages of the sequence (pre.cu) for the stream-level entry point: the rank's blocks are still in
// host memory, block k of the rank at base + k * stride; each sub-batch goes through the host stages and is
// uploaded just before it is encoded.
struct HostPrefix {
    const u8* base;
    i64 stride;
    int hs;
    int types[8];
};

static int dist_encode_dev(knz_ctx* ctx, u64 tType, int eType, int blockSize, const u8* d_in, i64 inStride,
                           const int32_t* lens, int nbOwn, int nBlocks, int firstBlockLen, u8* d_stream,
                           i64 streamCap, u64 startBit, u64* h_allBits, u64* endBit, const cudaEvent_t* batchReady,
                           int batchBlocks, const HostPrefix* hp = NULL)
{
    KnzDist* D = ctx->dist;
    cudaStream_t s = ctx->stream;
    const int W = D->world, R = D->rank;
    if (nbOwn != own_count(nBlocks, R, W) || nBlocks < 1 || nBlocks > 65535)
        return KNZ_ERR_INVALID_PARAM;
    const int nbMax = (nBlocks + W - 1) / W;
    int rc = knz_grow(ctx, &D->dBlk, &D->dBlkCap, (i64)nbMax * ctx->outStride + 256);
    if (rc != KNZ_OK)
        return rc;
    // meta layout (u64 units): own bits [nbMax] | all bits [W*nbMax] | ordered bits [nBlocks] | offsets [nBlocks + 8] | then int srcIndex[nBlocks]
    const i64 metaWords = (i64)nbMax + (i64)W * nbMax + 2 * (i64)nBlocks + 16;
    rc = knz_grow(ctx, &D->dMeta, &D->dMetaCap, metaWords * 8 + (i64)nBlocks * 4 + 256);
    if (rc != KNZ_OK)
        return rc;
    u64* dOwnBits = reinterpret_cast<u64*>(D->dMeta);
    u64* dAllBits = dOwnBits + nbMax;
    u64* dOrdered = dAllBits + (i64)W * nbMax;
    u64* dOff = dOrdered + nBlocks;
    int* dSrcIdx = reinterpret_cast<int*>(dOff + nBlocks + 16);
    DCK(cudaMemsetAsync(dOwnBits, 0, sizeof(u64) * (size_t)nbMax, s));
    float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    const int step = (batchBlocks > 0 && batchBlocks < ctx->maxBatch) ? batchBlocks : ctx->maxBatch;
    for (int off = 0, k = 0; off < nbOwn; off += step, k++) {
        int nb = (nbOwn - off < step) ? nbOwn - off : step;
        if (batchReady)
            DCK(cudaStreamWaitEvent(s, batchReady[k], 0));
        int ng = nb;
        if (lens[off + nb - 1] <= 15)
            ng = nb - 1; // only the last block of a stream can be that small: framed on the host
        if (ng > 0 && hp) {
            rc = knz_host_prefix_forward(ctx, hp->types, hp->hs, eType, blockSize, hp->base + (i64)off * hp->stride, hp->stride,
                                         lens + off, ng);
            if (rc != KNZ_OK)
                return rc;
            for (int b = 0; b < ng; b++)
                DCK(cudaMemcpyAsync(const_cast<u8*>(d_in) + (i64)(off + b) * inStride, ctx->h_pre + (i64)b * ctx->bstride,
                                    (size_t)ctx->h_init[b].len, cudaMemcpyHostToDevice, s));
            ctx->nHost = hp->hs;
        }
        if (ng > 0) {
            rc = knz_encode_batch(ctx, tType, eType, blockSize, d_in + (i64)off * inStride, inStride, lens + off, ng,
                                  firstBlockLen, D->dBlk + (i64)off * ctx->outStride, ctx->outStride, dOwnBits + off,
                                  NULL);
            ctx->nHost = 0;
            if (rc != KNZ_OK)
                return rc;
            for (int i = 0; i < 8; i++)
                acc[i] += ctx->ms[i];
        }
        if (ng < nb) {
            u8 raw[16], tmp[32];
            memset(tmp, 0, sizeof(tmp));
            const int len = lens[off + ng];
            if (hp)
                memcpy(raw, hp->base + (i64)(off + ng) * hp->stride, (size_t)len);
            else
                DCK(cudaMemcpyAsync(raw, d_in + (i64)(off + ng) * inStride, (size_t)len, cudaMemcpyDeviceToHost, s));
            DCK(cudaStreamSynchronize(s));
            const u64 bits = knz_frame_small_block(raw, len, tmp, ctx->checksumBits);
            DCK(cudaMemcpyAsync(D->dBlk + (i64)(off + ng) * ctx->outStride, tmp, 32, cudaMemcpyHostToDevice, s));
            DCK(cudaMemcpyAsync(dOwnBits + off + ng, &bits, sizeof(u64), cudaMemcpyHostToDevice, s));
            DCK(cudaStreamSynchronize(s));
        }
    }
    cudaEvent_t e0 = ctx->ev[5], e1 = ctx->ev[6];
    DCK(cudaEventRecord(e0, s));
    // ---- exchange: bit counts to everybody, payloads to rank 0
    rc = t_allgather(ctx, dOwnBits, (i64)nbMax * 8, dAllBits);
    if (rc != KNZ_OK)
        return rc;
    const size_t allN = (size_t)W * (size_t)nbMax;
    u64* hAll = (u64*)malloc(sizeof(u64) * (allN + 1));
    u64* hOrd = (u64*)malloc(sizeof(u64) * ((size_t)nBlocks + 1));
    int* hIdx = (int*)malloc(sizeof(int) * ((size_t)nBlocks + 1));
    cudaError_t ce = cudaMemcpyAsync(hAll, dAllBits, sizeof(u64) * allN, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess)
        ce = cudaStreamSynchronize(s);
    u64 maxBits = 0, total = startBit;
    for (int i = 0; i < nBlocks && ce == cudaSuccess; i++) {
        const int src = (i % W) * nbMax + i / W;
        hIdx[i] = src;
        hOrd[i] = hAll[src];
        if (hOrd[i] > maxBits)
            maxBits = hOrd[i];
        total += 5ull + prefix_bits(hOrd[i]) + hOrd[i];
        if (h_allBits)
            h_allBits[i] = hOrd[i];
    }
    i64 gstride = knz_round_up((i64)((maxBits + 7) >> 3) + 16, 256);
    if (gstride > ctx->outStride)
        gstride = ctx->outStride;
    rc = (ce == cudaSuccess) ? KNZ_OK : KNZ_ERR_PROCESS_BLOCK;
    if (rc == KNZ_OK)
        rc = knz_grow(ctx, &D->dPack, &D->dPackCap, (i64)nbMax * gstride + 256);
    if (rc == KNZ_OK && R == 0)
        rc = knz_grow(ctx, &D->dGather, &D->dGatherCap, (i64)W * nbMax * gstride + 256);
    if (rc == KNZ_OK && R == 0 && (streamCap < 0 || ((total + 8 + 31) >> 5) * 4 > (u64)streamCap))
        rc = KNZ_ERR_OUTPUT_TOO_SMALL;
    // (an error on one rank must not leave the others waiting in a collective: report after the exchange)
    const int rcLocal = rc;
    const u8* gatherSrc = D->dBlk;
    if (rcLocal == KNZ_OK && W > 1) {
        ce = cudaMemcpy2DAsync(D->dPack, (size_t)gstride, D->dBlk, (size_t)ctx->outStride, (size_t)gstride, (size_t)nbMax,
                               cudaMemcpyDeviceToDevice, s);
        gatherSrc = D->dPack;
    }
    const u8* blocks = D->dBlk;
    i64 bstrideG = ctx->outStride;
    if (W > 1) {
        rc = t_gather(ctx, gatherSrc, (i64)nbMax * gstride, (R == 0) ? D->dGather : D->dPack);
        blocks = D->dGather;
        bstrideG = gstride;
    }
    if (rc == KNZ_OK)
        rc = rcLocal;
    u64 endB = total;
    if (rc == KNZ_OK && R == 0) {
        u64* dStart = dOff + nBlocks;
        ce = cudaMemcpyAsync(dOrdered, hOrd, sizeof(u64) * (size_t)nBlocks, cudaMemcpyHostToDevice, s);
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(dSrcIdx, hIdx, sizeof(int) * (size_t)nBlocks, cudaMemcpyHostToDevice, s);
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(dStart, &startBit, sizeof(u64), cudaMemcpyHostToDevice, s);
        if (ce == cudaSuccess) {
            launch_stream_assemble(blocks, bstrideG, dOrdered, nBlocks, dStart, dOff, dStart + 1, d_stream, s,
                                   &ctx->launches, (W > 1) ? dSrcIdx : NULL);
            ce = cudaMemcpyAsync(&endB, dStart + 1, sizeof(u64), cudaMemcpyDeviceToHost, s);
        }
        if (ce == cudaSuccess)
            ce = cudaStreamSynchronize(s);
        if (ce != cudaSuccess)
            rc = KNZ_ERR_PROCESS_BLOCK;
    }
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    acc[4] = ms; // exchange + assembly
    for (int i = 0; i < 8; i++)
        ctx->ms[i] = acc[i];
    free(hAll);
    free(hOrd);
    free(hIdx);
    if (endBit)
        *endBit = endB;
    return rc;
}

extern "C" int knz_dist_encode_dev(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* d_in,
                                   int64_t inStride, const int32_t* lens, int nbOwn, int nBlocks, int firstBlockLen,
                                   uint8_t* d_stream, int64_t streamCap, uint64_t startBit, uint64_t* h_allBits,
                                   uint64_t* endBit)
{
    if (!ctx || !ctx->dist || !d_in || !lens || (ctx->dist->rank == 0 && !d_stream))
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    cudaSetDevice(ctx->device);
    return dist_encode_dev(ctx, tType, eType, blockSize, d_in, inStride, lens, nbOwn, nBlocks, firstBlockLen, d_stream,
                           streamCap, startBit, h_allBits, endBit, NULL, 0);
}

// Decode this rank's blocks out of the assembled stream (device memory).  d_stream holds the stream on
// rank 0 and is a receive buffer of the same size on the other ranks (broadcast inside).
extern "C" int knz_dist_decode_dev(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, uint8_t* d_stream,
                                   int64_t streamBytes, uint64_t startBit, const uint64_t* h_allBits, int nBlocks,
                                   uint8_t* d_out, int64_t outStride, int32_t* h_outLens)
{
    if (!ctx || !ctx->dist || !d_stream || !h_allBits || !d_out || !h_outLens || nBlocks < 1 || streamBytes < 0)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    cudaSetDevice(ctx->device);
    KnzDist* D = ctx->dist;
    cudaStream_t s = ctx->stream;
    const int W = D->world, R = D->rank;
    const int nbOwn = own_count(nBlocks, R, W);
    cudaEvent_t e0 = ctx->ev[5], e1 = ctx->ev[6];
    DCK(cudaEventRecord(e0, s));
    int rc = t_bcast(ctx, d_stream, streamBytes);
    DCK(cudaEventRecord(e1, s));
    if (rc != KNZ_OK)
        return rc;
    if (nbOwn == 0)
        return KNZ_OK;
    // block starts from the bit counts (the prefixes themselves are skipped)
    u64* start = (u64*)malloc(sizeof(u64) * (size_t)nbOwn);
    u64* endb = (u64*)malloc(sizeof(u64) * (size_t)nbOwn);
    u64* pay = (u64*)malloc(sizeof(u64) * (size_t)nbOwn);
    int* pre = (int*)malloc(sizeof(int) * (size_t)nbOwn);
    u8* fl = (u8*)malloc((size_t)nbOwn);
    u8* heads = (u8*)malloc((size_t)nbOwn * 16);
    u64* cks = (u64*)malloc(sizeof(u64) * (size_t)nbOwn);
    const int ckBits = ctx->checksumBits;
    u64 pos = startBit;
    for (int i = 0, k = 0; i < nBlocks; i++) {
        const u64 st = pos + 5 + prefix_bits(h_allBits[i]);
        if (i % W == R) {
            start[k] = st;
            endb[k] = st + h_allBits[i];
            k++;
        }
        pos = st + h_allBits[i];
    }
    rc = knz_grow(ctx, &D->dMeta, &D->dMetaCap, (i64)nbOwn * 32 + 256);
    cudaError_t ce = cudaSuccess;
    if (rc == KNZ_OK) {
        u64* dStart = reinterpret_cast<u64*>(D->dMeta);
        u8* dHeads = D->dMeta + (i64)nbOwn * 8;
        ce = cudaMemcpyAsync(dStart, start, sizeof(u64) * (size_t)nbOwn, cudaMemcpyHostToDevice, s);
        if (ce == cudaSuccess) {
            KLAUNCH(dist_heads_kernel, (nbOwn * 16 + 255) / 256, 256, s, d_stream, dStart, nbOwn, dHeads);
            ce = cudaMemcpyAsync(heads, dHeads, (size_t)nbOwn * 16, cudaMemcpyDeviceToHost, s);
        }
        if (ce == cudaSuccess)
            ce = cudaStreamSynchronize(s);
        if (ce != cudaSuccess)
            rc = KNZ_ERR_PROCESS_BLOCK;
    }
    float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int off = 0; off < nbOwn && rc == KNZ_OK; off += ctx->maxBatch) {
        int nb = (nbOwn - off < ctx->maxBatch) ? nbOwn - off : ctx->maxBatch;
        int ng = 0;
        for (int b = 0; b < nb && rc == KNZ_OK; b++) {
            const int g = off + b;
            const u64 rel = start[g] & 7;
            HostBitReader hb = { heads + 16 * g, 128, rel, false };
            const int k = knz_parse_block_header(hb, blockSize, &fl[g], &pre[g], ckBits, &cks[g]);
            if (k < 0) {
                rc = KNZ_ERR_INVALID_FILE;
            } else if (k == 1) { // copy block (only the last, short block of a stream): raw bytes follow
                if (b != nb - 1 || pre[g] > outStride) {
                    rc = KNZ_ERR_INVALID_FILE;
                } else {
                    KLAUNCH(dist_copy_bits_kernel, 1, 64, s, d_stream, start[g] + (hb.pos - rel), pre[g],
                            d_out + (i64)g * outStride);
                    h_outLens[g] = pre[g];
                }
            } else {
                pay[g] = start[g] + (hb.pos - rel);
                ng++;
            }
        }
        if (rc == KNZ_OK && ng > 0) {
            rc = knz_decode_batch(ctx, tType, eType, blockSize, d_stream, 0, pay + off, endb + off, pre + off, fl + off, ng,
                                  d_out + (i64)off * outStride, outStride, h_outLens + off, NULL, NULL, cks + off, ckBits);
            for (int i = 0; i < 8; i++)
                acc[i] += ctx->ms[i];
        }
    }
    cudaStreamSynchronize(s);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    acc[4] = ms; // broadcast of the stream
    for (int i = 0; i < 8; i++)
        ctx->ms[i] = acc[i];
    free(start);
    free(endb);
    free(pay);
    free(pre);
    free(fl);
    free(heads);
    free(cks);
    return rc;
}

static int ensure_events(knz_ctx* ctx, int n)
{
    KnzDist* D = ctx->dist;
    if (n > 64)
        return KNZ_ERR_INVALID_PARAM;
    while (D->nEvBatch < n) {
        DCK(cudaEventCreateWithFlags(&D->evBatch[D->nEvBatch], cudaEventDisableTiming));
        D->nEvBatch++;
    }
    return KNZ_OK;
}

// Stream level, host buffers: every rank passes the whole input; rank 0 receives the stream.
extern "C" int knz_compress_dist(knz_ctx* ctx, const char* transform, const char* entropy, int blockSize,
                                 const uint8_t* in, int64_t n, uint8_t* out, int64_t cap, int64_t* outLen)
{
    if (!ctx || !ctx->dist || !in || !outLen || n <= 0 || (ctx->dist->rank == 0 && !out))
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    const u64 tType = knz_transform_type(transform);
    const int eType = knz_entropy_type(entropy);
    if (tType == (u64)-1 || eType < 0)
        return KNZ_ERR_INVALID_CODEC;
    if (blockSize < 1024 || blockSize > ctx->maxBlockSize || (blockSize & 15))
        return KNZ_ERR_BLOCK_SIZE;
    cudaSetDevice(ctx->device);
    KnzDist* D = ctx->dist;
    cudaStream_t s = ctx->stream;
    const int W = D->world, R = D->rank;
    const i64 nBlocks64 = (n + blockSize - 1) / blockSize;
    if (nBlocks64 > 65535)
        return KNZ_ERR_INVALID_PARAM;
    const int nBlocks = (int)nBlocks64;
    const int nbOwn = own_count(nBlocks, R, W);
    const int nbMax = (nBlocks + W - 1) / W;
    u8 hdr[32];
    const int hdrBytes = knz_stream_header_ex(tType, eType, blockSize, n, ctx->checksumBits, hdr);
    HostPrefix hp;
    hp.hs = knz_host_prefix_len(hp.types, knz_split_types(tType, hp.types));
    if (hp.hs < 0 || (hp.hs > 0 && ctx->skipBlocks))
        return KNZ_ERR_INVALID_CODEC;
    hp.base = in + (i64)R * blockSize;
    hp.stride = (i64)W * blockSize;
    const i64 inSlot = (hp.hs > 0) ? ctx->bstride : (i64)blockSize; // host stages may expand a block
    // own blocks -> device, all copies queued up front on the copy stream, one event per sub-batch
    int rc = knz_grow(ctx, &D->dIn, &D->dInCap, (i64)(nbMax > 0 ? nbMax : 1) * inSlot + 256);
    if (rc != KNZ_OK)
        return rc;
    const int step = ctx->maxBatch;
    const int nBatches = (nbOwn + step - 1) / step;
    rc = ensure_events(ctx, nBatches > 0 ? nBatches : 1);
    if (rc != KNZ_OK)
        return rc;
    int32_t* lens = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nbOwn + 1));
    for (int k = 0; k < nbOwn; k++) {
        const i64 i = (i64)R + (i64)k * W;
        const i64 rem = n - i * blockSize;
        lens[k] = (int)((rem < blockSize) ? rem : blockSize);
        if (hp.hs > 0)
            continue; // uploaded per sub-batch, behind the host stages
        cudaMemcpyAsync(D->dIn + (i64)k * blockSize, in + i * blockSize, (size_t)lens[k], cudaMemcpyHostToDevice,
                        ctx->copyStream);
        if ((k + 1) % step == 0 || k == nbOwn - 1)
            cudaEventRecord(D->evBatch[k / step], ctx->copyStream);
    }
    // rank 0 assembles into the context's stream buffer
    const i64 refCapBlk = ((i64)blockSize + (blockSize >> 3) > 262144) ? (i64)blockSize + (blockSize >> 3) : 262144;
    const i64 worstBlk = ((2 * (i64)blockSize < refCapBlk) ? 2 * (i64)blockSize : refCapBlk);
    const i64 perBlk = worstBlk + (worstBlk >> 2) + 1024 + ((eType == E_ANS1) ? 131072 * (i64)((blockSize >> 22) + 1) : 0);
    const i64 streamCap = knz_round_up((i64)nBlocks * perBlk + 65536, 256);
    if (R == 0) {
        rc = knz_grow(ctx, &ctx->dStream, &ctx->dStreamCap, streamCap);
        if (rc == KNZ_OK && cudaMemsetAsync(ctx->dStream, 0, (size_t)streamCap, s) != cudaSuccess)
            rc = KNZ_ERR_PROCESS_BLOCK;
    }
    u64 endBit = 0;
    const int firstLen = (int)((n < blockSize) ? n : blockSize);
    // (a rank that failed locally still has to take part in the exchange: dist_encode_dev returns its own rc)
    if (rc == KNZ_OK)
        rc = dist_encode_dev(ctx, tType, eType, blockSize, D->dIn, inSlot, lens, nbOwn, nBlocks, firstLen, ctx->dStream,
                             streamCap, 8ull * (u64)hdrBytes, NULL, &endBit, (hp.hs > 0) ? NULL : D->evBatch, step,
                             (hp.hs > 0) ? &hp : NULL);
    free(lens);
    if (rc != KNZ_OK)
        return rc;
    *outLen = 0;
    if (R == 0) {
        const u64 endB = endBit + 8; // end marker: 5 + 3 zero bits (io/CompressedOutputStream.cpp:416-417)
        const i64 total = (i64)((endB + 7) >> 3);
        if (total > cap)
            return KNZ_ERR_OUTPUT_TOO_SMALL;
        DCK(cudaMemcpyAsync(out, ctx->dStream, (size_t)total, cudaMemcpyDeviceToHost, s));
        DCK(cudaStreamSynchronize(s));
        memcpy(out, hdr, (size_t)hdrBytes);
        *outLen = total;
    }
    return KNZ_OK;
}

// Every rank passes the whole stream and an output buffer for the whole original; each rank fills the
// blocks it owns (block i at out + i * blockSize).  *outLen = original length (stream header) or, when
// the header does not carry it, the end of the last block this rank decoded.
extern "C" int knz_decompress_dist(knz_ctx* ctx, const uint8_t* in, int64_t n, uint8_t* out, int64_t cap,
                                   int64_t* outLen)
{
    if (!ctx || !ctx->dist || !in || !out || !outLen || n < 20)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    cudaSetDevice(ctx->device);
    KnzDist* D = ctx->dist;
    cudaStream_t s = ctx->stream;
    const int W = D->world, R = D->rank;
    HostBitReader r = { in, 8ull * (u64)n, 0, false };
    KnzStreamInfo info;
    int rc = knz_parse_stream_header(ctx, r, &info);
    if (rc != KNZ_OK)
        return rc;
    const int blockSize = info.blockSize;
    int types[8];
    const int hs = knz_host_prefix_len(types, knz_split_types(info.tType, types));
    if (hs < 0)
        return KNZ_ERR_INVALID_CODEC;
    if (hs > 0 && knz_ensure_h_pre(ctx) != KNZ_OK)
        return KNZ_ERR_PROCESS_BLOCK;
    // ---- walk the length prefixes of the whole stream, keep the ranges of the blocks this rank owns
    struct Own {
        u64 start, bits, pay, ck;
        int pre, index;
        u8 flags;
    };
    Own* own = NULL;
    int nOwn = 0, capOwn = 0, nBlocks = 0;
    i64 lastEnd = 0;
    while (rc == KNZ_OK) {
        const int lr = 3 + (int)r.get(5);
        const u64 bits = r.get(lr);
        if (r.bad) {
            rc = KNZ_ERR_INVALID_FILE;
            break;
        }
        if (bits == 0)
            break;
        const u64 start = r.pos;
        if (start + bits > r.nbits) {
            rc = KNZ_ERR_INVALID_FILE;
            break;
        }
        if (nBlocks % W == R) {
            HostBitReader hb = { in, start + bits, start, false };
            u8 f = 0;
            int pl = 0;
            u64 ck = 0;
            const int k = knz_parse_block_header(hb, blockSize, &f, &pl, info.ckBits, &ck);
            const i64 dstOff = (i64)nBlocks * blockSize;
            if (k < 0) {
                rc = KNZ_ERR_INVALID_FILE;
                break;
            }
            if (k == 1) { // copy block: resolved on the host
                if (hb.pos + 8ull * (u64)pl > start + bits || dstOff + pl > cap) {
                    rc = KNZ_ERR_INVALID_FILE;
                    break;
                }
                for (int i = 0; i < pl; i++)
                    out[dstOff + i] = (u8)hb.get(8);
                if (info.ckBits && knz_xxhash_host(out + dstOff, pl, info.ckBits) != ck) {
                    rc = KNZ_ERR_CRC_CHECK;
                    break;
                }
                if (dstOff + pl > lastEnd)
                    lastEnd = dstOff + pl;
            } else {
                if (nOwn == capOwn) {
                    capOwn = capOwn ? 2 * capOwn : 64;
                    own = (Own*)realloc(own, sizeof(Own) * (size_t)capOwn);
                }
                Own o;
                o.start = start, o.bits = bits, o.pay = hb.pos, o.pre = pl, o.index = nBlocks, o.flags = f, o.ck = ck;
                own[nOwn++] = o;
            }
        }
        r.pos = start + bits;
        nBlocks++;
    }
    // ---- ship the bit ranges of the own blocks (whole bytes around them), decode, copy out
    u64 maxBits = 0;
    for (int k = 0; k < nOwn; k++)
        if (own[k].bits > maxBits)
            maxBits = own[k].bits;
    const i64 istride = knz_round_up((i64)((maxBits + 7) >> 3) + 32, 256);
    const int step = ctx->maxBatch;
    const int nBatches = (nOwn + step - 1) / step;
    if (rc == KNZ_OK)
        rc = ensure_events(ctx, nBatches > 0 ? nBatches : 1);
    if (rc == KNZ_OK)
        rc = knz_grow(ctx, &D->dIn, &D->dInCap, (i64)(nOwn > 0 ? nOwn : 1) * istride + 256);
    if (rc == KNZ_OK)
        rc = knz_grow(ctx, &D->dPlain, &D->dPlainCap, (i64)(nOwn > 0 ? nOwn : 1) * blockSize + 256);
    u64* pay = (u64*)malloc(sizeof(u64) * (size_t)(nOwn + 1));
    u64* endb = (u64*)malloc(sizeof(u64) * (size_t)(nOwn + 1));
    int* pre = (int*)malloc(sizeof(int) * (size_t)(nOwn + 1));
    u8* fl = (u8*)malloc((size_t)nOwn + 1);
    int32_t* ol = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nOwn + 1));
    u64* cks = (u64*)malloc(sizeof(u64) * (size_t)(nOwn + 1));
    for (int k = 0; k < nOwn; k++)
        cks[k] = own[k].ck;
    if (rc == KNZ_OK && nOwn > 0) {
        if (cudaMemsetAsync(D->dIn, 0, (size_t)((i64)nOwn * istride), ctx->copyStream) != cudaSuccess)
            rc = KNZ_ERR_PROCESS_BLOCK;
        for (int k = 0; k < nOwn && rc == KNZ_OK; k++) {
            const u64 b0 = own[k].start >> 3, b1 = (own[k].start + own[k].bits + 7) >> 3;
            if (cudaMemcpyAsync(D->dIn + (i64)k * istride, in + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice,
                                ctx->copyStream) != cudaSuccess)
                rc = KNZ_ERR_PROCESS_BLOCK;
            const u64 rel = own[k].start & 7;
            pay[k] = rel + (own[k].pay - own[k].start);
            endb[k] = rel + own[k].bits;
            pre[k] = own[k].pre;
            fl[k] = own[k].flags;
            if ((k + 1) % step == 0 || k == nOwn - 1)
                cudaEventRecord(D->evBatch[k / step], ctx->copyStream);
        }
    }
    float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int off = 0, kb = 0; off < nOwn && rc == KNZ_OK; off += step, kb++) {
        const int nb = (nOwn - off < step) ? nOwn - off : step;
        if (cudaStreamWaitEvent(s, D->evBatch[kb], 0) != cudaSuccess) {
            rc = KNZ_ERR_PROCESS_BLOCK;
            break;
        }
        if (hs > 0) {
            // device stages into stage-sized slots, then the leading host stages are undone on the host threads and
            // the blocks land at their place in the output
            ctx->nHost = hs;
            rc = knz_decode_batch(ctx, info.tType, info.eType, blockSize, D->dIn + (i64)off * istride, istride, pay + off,
                                  endb + off, pre + off, fl + off, nb, ctx->dStageIn, ctx->bstride, ol + off, NULL, NULL, NULL, 0);
            ctx->nHost = 0;
            if (rc != KNZ_OK)
                break;
            for (int i = 0; i < 8; i++)
                acc[i] += ctx->ms[i];
            for (int b = 0; b < nb; b++)
                cudaMemcpyAsync(ctx->h_pre + (i64)b * ctx->bstride, ctx->dStageIn + (i64)b * ctx->bstride, (size_t)ol[off + b],
                                cudaMemcpyDeviceToHost, s);
            if (cudaStreamSynchronize(s) != cudaSuccess) {
                rc = KNZ_ERR_PROCESS_BLOCK;
                break;
            }
            u8** dstp = (u8**)malloc(sizeof(u8*) * (size_t)nb);
            int* dcap = (int*)malloc(sizeof(int) * (size_t)nb * 3);
            int *dlen = dcap + nb, *ilen = dcap + 2 * nb;
            u8* good = (u8*)malloc((size_t)nb);
            for (int b = 0; b < nb; b++) {
                const i64 dstOff = (i64)own[off + b].index * blockSize;
                dstp[b] = out + ((dstOff < cap) ? dstOff : cap);
                const i64 room = cap - dstOff;
                dcap[b] = (int)((room < 0) ? 0 : (room < blockSize ? room : blockSize));
                ilen[b] = ol[off + b];
            }
            knz_host_prefix_inverse(ctx, types, hs, info.eType, blockSize, fl + off, ilen, nb, dstp, dcap, dlen, good);
            for (int b = 0; b < nb && rc == KNZ_OK; b++) {
                const i64 dstOff = (i64)own[off + b].index * blockSize;
                if (!good[b])
                    rc = (dlen[b] > dcap[b]) ? KNZ_ERR_OUTPUT_TOO_SMALL : KNZ_ERR_PROCESS_BLOCK;
                else if (info.ckBits && knz_xxhash_host(out + dstOff, dlen[b], info.ckBits) != cks[off + b])
                    rc = KNZ_ERR_CRC_CHECK;
                else if (dstOff + dlen[b] > lastEnd)
                    lastEnd = dstOff + dlen[b];
            }
            free(dstp);
            free(dcap);
            free(good);
            continue;
        }
        rc = knz_decode_batch(ctx, info.tType, info.eType, blockSize, D->dIn + (i64)off * istride, istride, pay + off,
                              endb + off, pre + off, fl + off, nb, D->dPlain + (i64)off * blockSize, blockSize, ol + off,
                              NULL, NULL, cks + off, info.ckBits);
        if (rc != KNZ_OK)
            break;
        for (int i = 0; i < 8; i++)
            acc[i] += ctx->ms[i];
        // decoded blocks go to their place while the next sub-batch decodes
        for (int b = 0; b < nb; b++) {
            const int k = off + b;
            const i64 dstOff = (i64)own[k].index * blockSize;
            if (ol[k] > blockSize || dstOff + ol[k] > cap) {
                rc = KNZ_ERR_OUTPUT_TOO_SMALL;
                break;
            }
            if (cudaMemcpyAsync(out + dstOff, D->dPlain + (i64)k * blockSize, (size_t)ol[k], cudaMemcpyDeviceToHost,
                                ctx->d2hStream) != cudaSuccess) {
                rc = KNZ_ERR_PROCESS_BLOCK;
                break;
            }
            if (dstOff + ol[k] > lastEnd)
                lastEnd = dstOff + ol[k];
        }
    }
    cudaStreamSynchronize(ctx->copyStream);
    cudaStreamSynchronize(ctx->d2hStream);
    for (int i = 0; i < 8; i++)
        ctx->ms[i] = acc[i];
    free(pay);
    free(endb);
    free(pre);
    free(fl);
    free(ol);
    free(cks);
    free(own);
    if (rc != KNZ_OK)
        return rc;
    *outLen = (info.origSize >= 0) ? info.origSize : lastEnd;
    return KNZ_OK;
}
