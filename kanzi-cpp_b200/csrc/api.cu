// api.cu -- context, C ABI (include/knz_gpu.h) and host-side orchestration:
// block scheduling, kanzi bitstream framing (stream header, block prefixes,
// end marker) and the batched launch sequence
//   transforms (BWT -> RANK/MTFT -> ZRLT)  ->  rANS  ->  bit assembly.
// The reference's equivalents are CompressedOutputStream / EncodingTask and
// CompressedInputStream / DecodingTask (io/Compressed{Output,Input}Stream.cpp).
#include <math.h>
#include <atomic>
#include <thread>
#include <vector>

#include "ctx.h"
#include "pre.h"

i64 knz_round_up(i64 v, i64 a) { return (v + a - 1) / a * a; }
static i64 round_up(i64 v, i64 a) { return knz_round_up(v, a); }

int knz_split_types(u64 tType, int* types)
{
    int n = 0;
    for (int i = 0; i < 8; i++) {
        const int t = (int)((tType >> (42 - 6 * i)) & 63);
        if (t != T_NONE || i == 0)
            types[n++] = t;
    }
    return n;
}

static bool entropy_supported(int e)
{
    return e == E_RAW || e == E_ANS0 || e == E_HUF || e == E_ANS1 || e == E_FPAQ;
}

static bool type_supported(int t)
{
    return t == T_NONE || t == T_BWT || t == T_ZRLT || t == T_MTFT || t == T_RANK || t == T_SRT || t == T_LZ ||
           t == T_LZX || t == T_LZP;
}
static bool is_lz(int t) { return t == T_LZ || t == T_LZX || t == T_LZP; }

// Transform<T>::getMaxEncodedLength: BWTBlockCodec n + 33, SRT n + 1024 (transform/SRT.hpp:38),
// LZ / LZX n + max(16, n/64) + 2 and LZP without the + 2 (transform/LZCodec.hpp:91-95, :158-161)
static int stage_max_len(int t, int n)
{
    if (is_lz(t))
        return ((n <= 1024) ? n + 16 : n + n / 64) + ((t == T_LZP) ? 0 : 2);
    if (knz_is_host_stage(t))
        return knz_pre_max_len(t, n);
    return (t == T_BWT) ? n + 33 : (t == T_SRT) ? n + 1024 : n;
}

static int required_size(const int* types, int nt, int n)
{
    int r = n;
    for (int i = 0; i < nt; i++) {
        const int m = stage_max_len(types[i], r);
        if (m > r)
            r = m;
    }
    return r;
}

extern "C" uint64_t knz_transform_type(const char* name)
{
    if (name == NULL)
        return (uint64_t)-1;
    u64 word = 0;
    int shift = 42, n = 0;
    const char* p = name;
    while (*p) {
        const char* q = strchr(p, '+');
        const size_t len = q ? (size_t)(q - p) : strlen(p);
        int t = -1;
        if (len == 4 && !strncmp(p, "NONE", 4))
            t = T_NONE;
        else if (len == 3 && !strncmp(p, "BWT", 3))
            t = T_BWT;
        else if (len == 4 && !strncmp(p, "ZRLT", 4))
            t = T_ZRLT;
        else if (len == 4 && !strncmp(p, "MTFT", 4))
            t = T_MTFT;
        else if (len == 4 && !strncmp(p, "RANK", 4))
            t = T_RANK;
        else if (len == 3 && !strncmp(p, "SRT", 3))
            t = T_SRT;
        else if (len == 2 && !strncmp(p, "LZ", 2))
            t = T_LZ;
        else if (len == 3 && !strncmp(p, "LZX", 3))
            t = T_LZX;
        else if (len == 3 && !strncmp(p, "LZP", 3))
            t = T_LZP;
        else if (len == 4 && !strncmp(p, "PACK", 4)) // host stages (pre.cu): a prefix of the sequence
            t = KNZ_T_PACK;
        else if (len == 3 && !strncmp(p, "DNA", 3))
            t = KNZ_T_DNA;
        else if (len == 2 && !strncmp(p, "MM", 2))
            t = KNZ_T_MM;
        else if (len == 3 && !strncmp(p, "UTF", 3))
            t = KNZ_T_UTF;
        else if (len == 4 && !strncmp(p, "TEXT", 4))
            t = KNZ_T_TEXT;
        if (t < 0 || ++n > 8)
            return (uint64_t)-1;
        if (t != T_NONE) {
            word |= (u64)t << shift;
            shift -= 6;
        }
        p += len;
        if (*p == '+')
            p++;
    }
    return word;
}

extern "C" int knz_entropy_type(const char* name)
{
    if (name == NULL)
        return -1;
    if (!strcmp(name, "NONE"))
        return E_RAW;
    if (!strcmp(name, "ANS0"))
        return E_ANS0;
    if (!strcmp(name, "HUFFMAN"))
        return E_HUF;
    if (!strcmp(name, "ANS1"))
        return E_ANS1;
    if (!strcmp(name, "FPAQ"))
        return E_FPAQ;
    return -1;
}

template <class T>
static cudaError_t dalloc(T** p, i64 count)
{
    return cudaMalloc((void**)p, (size_t)(count > 0 ? count : 1) * sizeof(T));
}

extern "C" int knz_create(int device, int maxBlockSize, int maxBatchBlocks, knz_ctx** out)
{
    if (out == NULL || maxBlockSize < 1024 || maxBlockSize > (64 << 20) || (maxBlockSize & 15) || maxBatchBlocks < 1)
        return KNZ_ERR_INVALID_PARAM;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return KNZ_ERR_CREATE_COMPRESSOR; // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess)
        return KNZ_ERR_CREATE_COMPRESSOR;
    knz_ctx* ctx = new (std::nothrow) knz_ctx();
    if (ctx == NULL)
        return KNZ_ERR_CREATE_COMPRESSOR;
    ctx->device = device;
    ctx->maxBlockSize = maxBlockSize;
    ctx->maxBatch = maxBatchBlocks;
    const int nb = maxBatchBlocks;
    // A stage buffer slot holds what the reference's task buffers can hold: max(bs + bs/8, 256 KiB)
    // (io/CompressedOutputStream.cpp:138-146) -- ZRLT at odd swap parity may legally expand a block up to
    // that capacity -- plus the BWT headers and 64 bytes of slack for vector accesses.
    const i64 refCap = ((i64)maxBlockSize + (maxBlockSize >> 3) > 262144) ? (i64)maxBlockSize + (maxBlockSize >> 3) : 262144;
    ctx->bstride = round_up(refCap + 33 * 8 + 64, 256);
    ctx->stageCap = (int)(ctx->bstride - 64);
    ctx->encSched[0] = 32, ctx->encSched[1] = 96, ctx->encSched[2] = 128;
    {
        const char* e = getenv("KNZ_ENC_BATCH");
        if (e)
            sscanf(e, "%d,%d,%d", &ctx->encSched[0], &ctx->encSched[1], &ctx->encSched[2]);
        for (int i = 0; i < 3; i++)
            if (ctx->encSched[i] < 1)
                ctx->encSched[i] = 1;
        const char* g = getenv("KNZ_DEC_GROUPS");
        ctx->decBwtGroups = (g && atoi(g) > 0) ? atoi(g) : 2; // 2: the per-launch latency of the node ranking outweighs more overlap
    }
    // a block's output: the largest post-transform block, 25 % expansion room, and the 256 context headers of every order-1
    // rANS chunk (<= 401 bytes each) -- incompressible data under ANS1 costs up to ~100 KiB per chunk
    ctx->outStride = round_up(refCap + (refCap >> 2) + 4096 + 131072 * (i64)((maxBlockSize >> 22) + 1), 256);
    ctx->maxChunks = (int)((ctx->bstride + ANS_CHUNK - 1) / ANS_CHUNK);
    bool ok = true;
#define A(call) ok = ok && ((call) == cudaSuccess)
    A(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    A(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
    A(cudaStreamCreateWithFlags(&ctx->d2hStream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        A(cudaEventCreate(&ctx->evCopy[i]));
        A(cudaEventCreate(&ctx->evDone[i]));
    }
    for (int i = 0; i < 10; i++)
        A(cudaEventCreate(&ctx->ev[i]));
    for (int i = 0; i < 16; i++)
        A(cudaEventCreate(&ctx->evStage[i]));
    for (int i = 0; i < KNZ_MAX_GROUPS; i++)
        A(cudaStreamCreateWithFlags(&ctx->gStream[i], cudaStreamNonBlocking));
    for (int i = 0; i <= KNZ_MAX_GROUPS; i++)
        A(cudaEventCreate(&ctx->gEv[i]));
    {
        // Block groups of the decode pipeline.  Default 1: measured on B200 (256 x 4 MiB blocks) 2 groups
        // = 193 ms (same as 1), 4 groups = 213 ms -- the inverse RANK chain warps lose issue slots to
        // the throughput kernels they share an SM with, which costs more than the overlap returns.
        const char* e = getenv("KNZ_DEC_OVERLAP");
        const int g = e ? atoi(e) : 1;
        ctx->decGroups = g < 1 ? 1 : (g > KNZ_MAX_GROUPS ? KNZ_MAX_GROUPS : g);
    }
    A(dalloc(&ctx->bufA, nb * ctx->bstride + 256));
    A(dalloc(&ctx->bufB, nb * ctx->bstride + 256));
    A(dalloc(&ctx->dStageIn, nb * ctx->bstride + 256));
    A(dalloc(&ctx->dOut, nb * ctx->outStride + 256));
    A(dalloc(&ctx->st, 10 * (i64)nb));
    A(dalloc(&ctx->capEven, nb));
    A(dalloc(&ctx->capOdd, nb));
    const i64 nch = (i64)nb * ctx->maxChunks;
    A(dalloc(&ctx->slots, nch * ANS_SLOT));
    A(dalloc(&ctx->hdrBits, nch));
    A(dalloc(&ctx->payBytes, nch));
    A(dalloc(&ctx->payOff, nch));
    A(dalloc(&ctx->chunkOff, nch));
    A(dalloc(&ctx->chunkPos, nch));
    A(dalloc(&ctx->blockBits, nb));
    A(dalloc(&ctx->blockOff, nb));
    A(dalloc(&ctx->streamPos, 4));
    A(dalloc(&ctx->dInBits, nb));
    A(dalloc(&ctx->dPayStart, nb));
    A(dalloc(&ctx->dPreLen, nb));
    A(dalloc(&ctx->errFlag, 4));
    A(dalloc(&ctx->blockHash, nb));
    A(dalloc(&ctx->expectHash, nb));
    A(cudaMallocHost((void**)&ctx->h_hash, sizeof(u64) * nb));
    A(dalloc(&ctx->dSkip, nb));
    A(dalloc(&ctx->dLog2Tab, 257));
    A(cudaMallocHost((void**)&ctx->h_skip, sizeof(int) * nb));
    A(dalloc(&ctx->dDtype, nb));
    A(cudaMallocHost((void**)&ctx->h_dtype, sizeof(int) * nb));
    A(cudaMallocHost((void**)&ctx->h_init, sizeof(BlkState) * nb));
    if (ok) { // Global::LOG2_4096 (Global.cpp:47-74) is round(4096 * log2(i)); tests compare all 257 entries
        int tab[257];
        tab[0] = 0;
        for (int i = 1; i <= 256; i++)
            tab[i] = (int)floor(4096.0 * log2((double)i) + 0.5);
        A(cudaMemcpy(ctx->dLog2Tab, tab, sizeof(tab), cudaMemcpyHostToDevice));
    }
    A(cudaMallocHost((void**)&ctx->h_st, sizeof(BlkState) * nb));
    A(cudaMallocHost((void**)&ctx->h_capEven, sizeof(int) * nb));
    A(cudaMallocHost((void**)&ctx->h_capOdd, sizeof(int) * nb));
    A(cudaMallocHost((void**)&ctx->h_err, sizeof(int) * 4));
    A(cudaMallocHost((void**)&ctx->h_preLen, sizeof(int) * nb));
    A(cudaMallocHost((void**)&ctx->h_bits, sizeof(u64) * nb));
    A(cudaMallocHost((void**)&ctx->h_payStart, sizeof(u64) * nb));
    A(cudaMallocHost((void**)&ctx->h_pos, sizeof(u64) * 4));
    ok = ok && workspace_alloc(ctx->ws, nb, (int)ctx->bstride);
#undef A
    if (!ok) {
        knz_destroy(ctx);
        return KNZ_ERR_CREATE_COMPRESSOR;
    }
    *out = ctx;
    return KNZ_OK;
}

extern "C" void knz_destroy(knz_ctx* ctx)
{
    if (ctx == NULL)
        return;
    cudaSetDevice(ctx->device);
    if (ctx->stream)
        cudaStreamSynchronize(ctx->stream);
    knz_dist_destroy(ctx);
    workspace_free(ctx->ws);
    if (ctx->a1Ready)
        ans1_work_free(ctx->a1);
    if (ctx->srtReady)
        srt_work_free(ctx->srt);
    if (ctx->lzReady)
        lz_work_free(ctx->lz);
    void* dev[] = { ctx->bufA, ctx->bufB, ctx->dStageIn, ctx->dOut, ctx->st, ctx->capEven, ctx->capOdd, ctx->slots,
                    ctx->hdrBits, ctx->payBytes, ctx->payOff, ctx->chunkOff, ctx->chunkPos, ctx->blockBits,
                    ctx->blockOff, ctx->streamPos, ctx->dInBits, ctx->dPayStart, ctx->dPreLen, ctx->errFlag,
                    ctx->dStream, ctx->dPlain, ctx->dPlain2, ctx->blockHash, ctx->expectHash, ctx->dSkip,
                    ctx->dLog2Tab, ctx->dDtype };
    for (size_t i = 0; i < sizeof(dev) / sizeof(dev[0]); i++)
        if (dev[i])
            cudaFree(dev[i]);
    void* hst[] = { ctx->h_st, ctx->h_capEven, ctx->h_capOdd, ctx->h_err, ctx->h_preLen, ctx->h_bits,
                    ctx->h_payStart, ctx->h_pos, ctx->h_hash, ctx->h_skip, ctx->h_dtype, ctx->h_init, ctx->h_pre };
    for (size_t i = 0; i < sizeof(hst) / sizeof(hst[0]); i++)
        if (hst[i])
            cudaFreeHost(hst[i]);
    for (int i = 0; i < 10; i++)
        if (ctx->ev[i])
            cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 16; i++)
        if (ctx->evStage[i])
            cudaEventDestroy(ctx->evStage[i]);
    for (int i = 0; i < 2; i++) {
        if (ctx->evCopy[i])
            cudaEventDestroy(ctx->evCopy[i]);
        if (ctx->evDone[i])
            cudaEventDestroy(ctx->evDone[i]);
    }
    for (int i = 0; i < KNZ_MAX_GROUPS; i++)
        if (ctx->gStream[i])
            cudaStreamDestroy(ctx->gStream[i]);
    for (int i = 0; i <= KNZ_MAX_GROUPS; i++)
        if (ctx->gEv[i])
            cudaEventDestroy(ctx->gEv[i]);
    if (ctx->copyStream)
        cudaStreamDestroy(ctx->copyStream);
    if (ctx->d2hStream)
        cudaStreamDestroy(ctx->d2hStream);
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int knz_set_decode_groups(knz_ctx* ctx, int groups)
{
    if (ctx == NULL || groups < 1 || groups > KNZ_MAX_GROUPS)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    ctx->decGroups = groups;
    return KNZ_OK;
}

extern "C" int knz_set_checksum(knz_ctx* ctx, int bits)
{
    if (ctx == NULL || (bits != 0 && bits != 32 && bits != 64))
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    ctx->checksumBits = bits;
    return KNZ_OK;
}

extern "C" int knz_set_skip_blocks(knz_ctx* ctx, int on)
{
    if (ctx == NULL)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    ctx->skipBlocks = on ? 1 : 0;
    return KNZ_OK;
}

extern "C" int knz_set_listener(knz_ctx* ctx, knz_event_fn fn, void* user)
{
    if (ctx == NULL)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    ctx->listener = fn;
    ctx->listenerUser = user;
    return KNZ_OK;
}

static void emit_event(knz_ctx* ctx, int type, int blockId, i64 size, u64 hash, int hashBits, i64 offset, int skipFlags)
{
    knz_event e;
    e.type = type;
    e.blockId = blockId;
    e.size = size;
    e.hash = hashBits ? hash : 0;
    e.hashBits = hashBits;
    e.offset = offset;
    e.skipFlags = (uint8_t)skipFlags;
    ctx->listener(ctx->listenerUser, &e);
}

extern "C" const char* knz_last_error(const knz_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" uint64_t knz_launch_count(const knz_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* knz_stream(const knz_ctx* ctx) { return ctx ? (void*)ctx->stream : NULL; }
extern "C" void knz_last_timings(const knz_ctx* ctx, float ms[8])
{
    for (int i = 0; i < 8; i++)
        ms[i] = ctx ? ctx->ms[i] : 0.f;
}

static int map_kerr(knz_ctx* ctx, int kerr)
{
    if (kerr == 0)
        return KNZ_OK;
    snprintf(ctx->err, sizeof(ctx->err), "device error flag %d (%s)", kerr,
             kerr == KERR_OUT_OVERFLOW  ? "output buffer overflow"
             : kerr == KERR_BAD_STREAM  ? "invalid bitstream"
             : kerr == KERR_UNSUPPORTED ? "unsupported stream feature"
             : kerr == KERR_CRC         ? "block checksum mismatch"
                                        : "internal");
    return (kerr == KERR_BAD_STREAM) ? KNZ_ERR_INVALID_FILE : (kerr == KERR_CRC) ? KNZ_ERR_CRC_CHECK : KNZ_ERR_PROCESS_BLOCK;
}

static void add_stage_time(knz_ctx* ctx, int t, float ms)
{
    const int slot = (t == T_BWT) ? 0 : (t == T_RANK || t == T_MTFT || t == T_SRT || is_lz(t)) ? 1 : (t == T_ZRLT) ? 2 : 5;
    if (slot < 5)
        ctx->ms[slot] += ms;
}

// ANS1 keeps a 256 x 256 table set and a pre-mapped record stream per chunk: only contexts that
// use it pay for the memory.
static int ensure_ans1(knz_ctx* ctx, int eType)
{
    if (eType != E_ANS1 || ctx->a1Ready)
        return KNZ_OK;
    if (!ans1_work_alloc(ctx->a1, ctx->maxBatch, ctx->stageCap)) {
        snprintf(ctx->err, sizeof(ctx->err), "out of device memory for the ANS1 tables");
        return KNZ_ERR_CREATE_COMPRESSOR;
    }
    ctx->a1Ready = true;
    return KNZ_OK;
}

static int ensure_bwt(knz_ctx* ctx, const int* types, int nt)
{
    for (int i = 0; i < nt; i++) {
        if (types[i] == T_BWT && !ctx->ws.bwtReady && !workspace_alloc_bwt(ctx->ws)) {
            snprintf(ctx->err, sizeof(ctx->err), "out of device memory for the suffix-sort workspace");
            return KNZ_ERR_CREATE_COMPRESSOR;
        }
        if (types[i] == T_SRT && !ctx->srtReady) {
            if (!srt_work_alloc(ctx->srt, ctx->maxBatch, (int)ctx->bstride)) {
                snprintf(ctx->err, sizeof(ctx->err), "out of device memory for the SRT workspace");
                return KNZ_ERR_CREATE_COMPRESSOR;
            }
            ctx->srtReady = true;
        }
        if (is_lz(types[i]) && !ctx->lzReady) {
            if (!lz_work_alloc(ctx->lz, ctx->maxBatch, ctx->bstride)) {
                snprintf(ctx->err, sizeof(ctx->err), "out of device memory for the LZ workspace");
                return KNZ_ERR_CREATE_COMPRESSOR;
            }
            ctx->lzReady = true;
        }
    }
    return KNZ_OK;
}

static void launch_forward_stage(knz_ctx* ctx, int type, const StageLaunch& L, cudaStream_t s)
{
    switch (type) {
    case T_NONE:
        launch_none_forward(L, s, &ctx->launches);
        break;
    case T_BWT:
        launch_bwt_forward(L, ctx->ws, s, &ctx->launches);
        break;
    case T_ZRLT:
        launch_zrlt_forward(L, ctx->ws, s, &ctx->launches);
        break;
    case T_MTFT:
        launch_sbrt_forward(L, 1, ctx->ws, s, &ctx->launches);
        break;
    case T_RANK:
        launch_sbrt_forward(L, 2, ctx->ws, s, &ctx->launches);
        break;
    case T_SRT:
        launch_srt_forward(L, ctx->ws, ctx->srt, s, &ctx->launches);
        break;
    case T_LZ:
    case T_LZX:
    case T_LZP:
        launch_lz_forward(L, type, ctx->lz, s, &ctx->launches);
        break;
    }
}

static void launch_inverse_stage(knz_ctx* ctx, int type, const StageLaunch& L, cudaStream_t s)
{
    switch (type) {
    case T_NONE:
        launch_none_forward(L, s, &ctx->launches); // inverse of a copy is a copy
        break;
    case T_BWT:
        launch_bwt_inverse(L, ctx->ws, s, &ctx->launches);
        break;
    case T_ZRLT:
        launch_zrlt_inverse(L, ctx->ws, s, &ctx->launches);
        break;
    case T_MTFT:
        launch_sbrt_inverse(L, 1, ctx->ws, s, &ctx->launches);
        break;
    case T_RANK:
        launch_sbrt_inverse(L, 2, ctx->ws, s, &ctx->launches);
        break;
    case T_SRT:
        launch_srt_inverse(L, s, &ctx->launches);
        break;
    case T_LZ:
    case T_LZX:
    case T_LZP:
        launch_lz_inverse(L, type, ctx->lz, s, &ctx->launches);
        break;
    }
}

// Forward transforms + entropy for one batch of blocks resident on the device.
// All lens[i] must be > 15 (small blocks are framed on the host).
int knz_encode_batch(knz_ctx* ctx, u64 tType, int eType, int blockSize, const u8* d_in, i64 inStride,
                        const int32_t* lens, int nB, int firstBlockLen, u8* d_out, i64 outStride, u64* d_bits,
                        u8* h_flags)
{
    int types[8];
    const int nt = knz_split_types(tType, types);
    const int hs = ctx->nHost; // leading stages the caller has applied on the host already (pre.cu)
    for (int i = 0; i < nt; i++)
        if (!(type_supported(types[i]) || (i < hs && knz_is_host_stage(types[i])))) {
            snprintf(ctx->err, sizeof(ctx->err), "transform id %d not implemented%s", types[i],
                     knz_is_host_stage(types[i]) ? " here (host stages: leading stages of knz_compress / knz_decompress only)" : "");
            return KNZ_ERR_INVALID_CODEC;
        }
    if (!entropy_supported(eType)) {
        snprintf(ctx->err, sizeof(ctx->err), "entropy id %d not implemented", eType);
        return KNZ_ERR_INVALID_CODEC;
    }
    if (hs > 0 && ctx->skipBlocks)
        return KNZ_ERR_INVALID_PARAM; // the entropy test belongs to the original block, which the device does not see
    {
        int rc1 = ensure_ans1(ctx, eType);
        if (rc1 == KNZ_OK)
            rc1 = ensure_bwt(ctx, types + hs, nt - hs);
        if (rc1 != KNZ_OK)
            return rc1;
    }
    if (nB > ctx->maxBatch || ((uintptr_t)d_in & 15) || (inStride & 15) || ((uintptr_t)d_out & 15) || (outStride & 15))
        return KNZ_ERR_INVALID_PARAM;
    // the context is provisioned for maxBlockSize: larger stream parameters would overrun its scratch
    if (blockSize < 1 || blockSize > ctx->maxBlockSize || firstBlockLen < 0 || firstBlockLen > ctx->maxBlockSize)
        return KNZ_ERR_BLOCK_SIZE;
    cudaStream_t s = ctx->stream;
    const int dataCap = (blockSize + (blockSize >> 3) > 262144) ? blockSize + (blockSize >> 3) : 262144;
    const int reqFirst = required_size(types, nt, firstBlockLen);
    int maxLen = 0;
    for (int b = 0; b < nB; b++) {
        if (lens[b] <= 15 || lens[b] > ctx->maxBlockSize)
            return KNZ_ERR_BLOCK_SIZE;
        const int req = required_size(types, nt, lens[b]);
        ctx->h_st[b].len = lens[b];
        ctx->h_st[b].cur = 2;
        ctx->h_st[b].swaps = 0;
        ctx->h_st[b].flags = 0xFF;
        if (hs > 0) { // length, ping-pong parity and skip flags as the host stages left them
            ctx->h_st[b] = ctx->h_init[b];
            ctx->h_st[b].cur = 2;
            if (ctx->h_st[b].len < 1 || ctx->h_st[b].len > ctx->stageCap)
                return KNZ_ERR_BLOCK_SIZE;
        }
        ctx->h_capEven[b] = (reqFirst > req) ? reqFirst : req;
        ctx->h_capOdd[b] = (dataCap >= req) ? dataCap : req;
        // never beyond what a stage buffer slot holds (cannot bind for blockSize <= maxBlockSize)
        if (ctx->h_capEven[b] > ctx->stageCap)
            ctx->h_capEven[b] = ctx->stageCap;
        if (ctx->h_capOdd[b] > ctx->stageCap)
            ctx->h_capOdd[b] = ctx->stageCap;
        if (ctx->h_st[b].len > maxLen)
            maxLen = ctx->h_st[b].len;
        if (lens[b] > maxLen)
            maxLen = lens[b];
    }
    CK(cudaMemcpyAsync(ctx->st, ctx->h_st, sizeof(BlkState) * nB, cudaMemcpyHostToDevice, s));
    if (hs > 0)
        CK(cudaMemcpyAsync(ctx->dDtype, ctx->h_dtype, sizeof(int) * nB, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->capEven, ctx->h_capEven, sizeof(int) * nB, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->capOdd, ctx->h_capOdd, sizeof(int) * nB, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->errFlag, 0, sizeof(int) * 4, s));

    BufTable bt;
    bt.base[0] = ctx->bufA;
    bt.base[1] = ctx->bufB;
    bt.base[2] = const_cast<u8*>(d_in);
    bt.stride[0] = bt.stride[1] = ctx->bstride;
    bt.stride[2] = inStride;

    for (int i = 0; i < 8; i++)
        ctx->ms[i] = 0.f;
    CK(cudaEventRecord(ctx->ev[0], s));
    if (ctx->checksumBits && hs > 0) // hashed by the caller before the host stages ran (ctx->h_hash)
        CK(cudaMemcpyAsync(ctx->blockHash, ctx->h_hash, sizeof(u64) * nB, cudaMemcpyHostToDevice, s));
    else if (ctx->checksumBits) // XXHash of every block before the transforms (io/CompressedOutputStream.cpp:674-682)
        launch_xxhash(bt, ctx->st, nB, ctx->checksumBits, ctx->blockHash, NULL, ctx->errFlag, s, &ctx->launches);
    if (ctx->skipBlocks) // entropy / signature test (io/CompressedOutputStream.cpp:697-715)
        launch_skip_decide(bt, ctx->st, nB, ctx->dLog2Tab, ctx->dSkip, s, &ctx->launches);
    // Upper bound of the block lengths a stage may meet (its launches size their grids with it): 33 bytes per
    // stage so far (the BWT header; bwt.cu subtracts exactly that to recover its input bound), the SRT header
    // (up to 1024 bytes, transform/SRT.hpp:38) once an SRT stage is involved, and whatever a slot can hold behind a
    // ZRLT stage, which may legally expand a block at odd swap parity (reference quirk 2).
    int grown = 0;
    bool afterZrlt = false;
    for (int i = hs; i < nt; i++) {
        StageLaunch L;
        L.bt = bt;
        L.stIn = ctx->st + (i64)(i - hs) * ctx->maxBatch;
        L.stOut = ctx->st + (i64)(i - hs + 1) * ctx->maxBatch;
        L.dtype = (hs > 0) ? ctx->dDtype : NULL;
        L.stageIdx = i;
        L.nBlocks = nB;
        if (types[i] == T_SRT)
            grown += 1024 - 33;
        L.maxLen = maxLen + 33 * (i + 1) + grown;
        if (afterZrlt || L.maxLen > ctx->stageCap)
            L.maxLen = ctx->stageCap;
        if (types[i] == T_ZRLT)
            afterZrlt = true;
        L.capEven = ctx->capEven;
        L.capOdd = ctx->capOdd;
        L.errFlag = ctx->errFlag;
        L.wsBlock0 = 0;
        CK(cudaEventRecord(ctx->evStage[2 * i], s));
        launch_forward_stage(ctx, types[i], L, s);
        CK(cudaEventRecord(ctx->evStage[2 * i + 1], s));
    }
    const BlkState* stFinal = ctx->st + (i64)(nt - hs) * ctx->maxBatch;
    CK(cudaEventRecord(ctx->ev[3], s));
    EncodeLaunch E;
    E.bt = bt;
    E.st = stFinal;
    E.nBlocks = nB;
    E.maxChunks = ctx->maxChunks;
    E.eType = eType;
    E.nTransforms = nt | ((ctx->checksumBits >> 3) << 8);
    E.blockHash = ctx->blockHash;
    E.slots = ctx->slots;
    E.hdrBits = ctx->hdrBits;
    E.payBytes = ctx->payBytes;
    E.payOff = ctx->payOff;
    E.chunkOff = ctx->chunkOff;
    E.blockBits = d_bits;
    E.out = d_out;
    E.outStride = outStride;
    E.errFlag = ctx->errFlag;
    E.evK0 = ctx->ev[8];
    E.evK1 = ctx->ev[9];
    E.a1 = &ctx->a1;
    launch_entropy_encode(E, s, &ctx->launches);
    if (ctx->skipBlocks) { // flagged blocks: their private buffer becomes the copy block
        launch_copy_frame(bt, ctx->st, nB, ctx->dSkip, ctx->checksumBits >> 3, ctx->blockHash, d_out, outStride, d_bits, s,
                          &ctx->launches);
        CK(cudaMemcpyAsync(ctx->h_skip, ctx->dSkip, sizeof(int) * nB, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaEventRecord(ctx->ev[4], s));
    CK(cudaMemcpyAsync(ctx->h_err, ctx->errFlag, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_st, stFinal, sizeof(BlkState) * nB, cudaMemcpyDeviceToHost, s));
    if (ctx->listener) {
        CK(cudaMemcpyAsync(ctx->h_bits, d_bits, sizeof(u64) * nB, cudaMemcpyDeviceToHost, s));
        if (ctx->checksumBits)
            CK(cudaMemcpyAsync(ctx->h_hash, ctx->blockHash, sizeof(u64) * nB, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    for (int b = 0; b < nB; b++) {
        if (!ctx->skipBlocks)
            ctx->h_skip[b] = 0;
        if (ctx->h_skip[b]) { // what the reference reports for a copy block: one NullTransform, applied
            ctx->h_st[b].len = lens[b];
            ctx->h_st[b].flags = 0x7F;
        }
    }
    float ms = 0.f;
    for (int i = hs; i < nt; i++) { // stage brackets were recorded without stalling the stream
        cudaEventElapsedTime(&ms, ctx->evStage[2 * i], ctx->evStage[2 * i + 1]);
        add_stage_time(ctx, types[i], ms);
    }
    cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
    ctx->ms[3] = ms;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[4]);
    ctx->ms[5] = ms;
    if (eType != E_RAW) {
        cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]);
        ctx->ms[6] = ms;
    }
    if (h_flags)
        for (int b = 0; b < nB; b++)
            h_flags[b] = (u8)ctx->h_st[b].flags;
    return map_kerr(ctx, ctx->h_err[0] ? ctx->h_err[0] : ctx->h_err[1] ? ctx->h_err[1] : ctx->h_err[2]);
}

extern "C" int knz_encode_blocks_dev(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* d_in,
                                     int64_t inStride, const int32_t* lens, int nBlocks, int firstBlockLen,
                                     uint8_t* d_blockOut, int64_t outStride, uint64_t* d_outBits,
                                     uint8_t* h_skipFlags)
{
    if (!ctx || !d_in || !lens || !d_blockOut || !d_outBits || nBlocks < 0)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    cudaSetDevice(ctx->device);
    float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int off = 0; off < nBlocks; off += ctx->maxBatch) {
        const int nb = (nBlocks - off < ctx->maxBatch) ? nBlocks - off : ctx->maxBatch;
        const int rc = knz_encode_batch(ctx, tType, eType, blockSize, d_in + (i64)off * inStride, inStride, lens + off, nb,
                                    firstBlockLen, d_blockOut + (i64)off * outStride, outStride, d_outBits + off,
                                    h_skipFlags ? h_skipFlags + off : NULL);
        if (rc != KNZ_OK)
            return rc;
        for (int i = 0; i < 8; i++)
            acc[i] += ctx->ms[i];
    }
    for (int i = 0; i < 8; i++)
        ctx->ms[i] = acc[i];
    return KNZ_OK;
}

// Small blocks (<= 15 bytes) are always copy blocks: mode 0x80 | skipFlags>>4
// with the single NullTransform applied (flags 0x7F), one length byte, raw bytes
// (io/CompressedOutputStream.cpp:38,691-695).
u64 knz_frame_small_block(const u8* in, int len, u8* out, int ckBits)
{
    out[0] = 0x87;
    out[1] = (u8)len;
    int p = 2;
    if (ckBits) {
        const u64 h = knz_xxhash_host(in, len, ckBits);
        for (int sh = ckBits - 8; sh >= 0; sh -= 8)
            out[p++] = (u8)(h >> sh);
    }
    memcpy(out + p, in, (size_t)len);
    return 8ull * (u64)(p + len);
}

extern "C" int knz_encode_blocks(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* in,
                                 int64_t inStride, const int32_t* lens, int nBlocks, int firstBlockLen,
                                 uint8_t* out, int64_t outStride, uint64_t* outBits, uint8_t* skipFlags)
{
    if (!ctx || !in || !lens || !out || !outBits || nBlocks < 0)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    for (int off = 0; off < nBlocks; off += ctx->maxBatch) {
        const int nb = (nBlocks - off < ctx->maxBatch) ? nBlocks - off : ctx->maxBatch;
        // small blocks can only be framed on the host; the rest go to the device
        int32_t glens[1];
        (void)glens;
        int ng = 0;
        int* map = (int*)malloc(sizeof(int) * (size_t)nb);
        int32_t* dl = (int32_t*)malloc(sizeof(int32_t) * (size_t)nb);
        for (int b = 0; b < nb; b++) {
            const int len = lens[off + b];
            if (len <= 0 || len > ctx->maxBlockSize) {
                free(map);
                free(dl);
                return KNZ_ERR_BLOCK_SIZE;
            }
            if (len <= 15) {
                outBits[off + b] = knz_frame_small_block(in + (i64)(off + b) * inStride, len, out + (i64)(off + b) * outStride,
                                                         ctx->checksumBits);
                if (skipFlags)
                    skipFlags[off + b] = 0x7F;
            } else {
                CK(cudaMemcpyAsync(ctx->dStageIn + (i64)ng * ctx->bstride, in + (i64)(off + b) * inStride, (size_t)len,
                                   cudaMemcpyHostToDevice, s));
                map[ng] = off + b;
                dl[ng] = len;
                ng++;
            }
        }
        int rc = KNZ_OK;
        u8* fl = (u8*)malloc((size_t)nb + 1);
        if (ng > 0)
            rc = knz_encode_batch(ctx, tType, eType, blockSize, ctx->dStageIn, ctx->bstride, dl, ng, firstBlockLen, ctx->dOut,
                              ctx->outStride, ctx->blockBits, fl);
        if (rc == KNZ_OK && ng > 0) {
            cudaMemcpyAsync(ctx->h_bits, ctx->blockBits, sizeof(u64) * ng, cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            for (int g = 0; g < ng && rc == KNZ_OK; g++) {
                const i64 nbytes = (i64)((ctx->h_bits[g] + 7) >> 3);
                if (nbytes > outStride) {
                    rc = KNZ_ERR_OUTPUT_TOO_SMALL;
                    break;
                }
                cudaMemcpyAsync(out + (i64)map[g] * outStride, ctx->dOut + (i64)g * ctx->outStride, (size_t)nbytes,
                                cudaMemcpyDeviceToHost, s);
                outBits[map[g]] = ctx->h_bits[g];
                if (skipFlags)
                    skipFlags[map[g]] = fl[g];
            }
            cudaStreamSynchronize(s);
        }
        free(fl);
        free(map);
        free(dl);
        if (rc != KNZ_OK)
            return rc;
    }
    return KNZ_OK;
}

extern "C" int knz_assemble_stream_dev(knz_ctx* ctx, const uint8_t* d_blockOut, int64_t outStride,
                                       const uint64_t* d_outBits, int nBlocks, uint8_t* d_stream, int64_t streamCap,
                                       uint64_t startBit, uint64_t* endBit)
{
    if (!ctx || nBlocks < 0 || ((uintptr_t)d_stream & 3))
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ctx->h_pos[0] = startBit;
    CK(cudaMemcpyAsync(ctx->streamPos, ctx->h_pos, sizeof(u64), cudaMemcpyHostToDevice, s));
    if (nBlocks > 0) {
        if (nBlocks > ctx->maxBatch)
            return KNZ_ERR_INVALID_PARAM;
        // the assembled bits must fit the caller's buffer: sum the prefixes + payloads before writing
        CK(cudaMemcpyAsync(ctx->h_bits, d_outBits, sizeof(u64) * (size_t)nBlocks, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        u64 total = startBit;
        for (int b = 0; b < nBlocks; b++) {
            int lw = 3;
            while (lw < 34 && (ctx->h_bits[b] >> lw) != 0)
                lw++;
            total += 5ull + (u64)lw + ctx->h_bits[b];
        }
        if (streamCap < 0 || ((total + 8 + 31) >> 5) * 4 > (u64)streamCap)
            return KNZ_ERR_OUTPUT_TOO_SMALL;
        launch_stream_assemble(d_blockOut, outStride, d_outBits, nBlocks, ctx->streamPos, ctx->blockOff, ctx->streamPos,
                               d_stream, s, &ctx->launches);
    }
    CK(cudaMemcpyAsync(ctx->h_pos, ctx->streamPos, sizeof(u64), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (endBit)
        *endBit = ctx->h_pos[0];
    return KNZ_OK;
}

// Stream header (io/CompressedOutputStream.cpp:277-342), written with a tiny
// host-side MSB-first packer.
struct HostBits {
    u8* p;
    u64 pos;
    void put(u64 v, int n)
    {
        for (int k = n - 1; k >= 0; k--) {
            if ((v >> k) & 1)
                p[pos >> 3] |= (u8)(0x80 >> (pos & 7));
            pos++;
        }
    }
};

extern "C" int knz_stream_header_ex(uint64_t tType, int eType, int blockSize, int64_t inputSize, int checksumBits,
                                    uint8_t out[32])
{
    const u32 ckSize = (checksumBits == 32) ? 1u : (checksumBits == 64) ? 2u : 0u; // :290-296
    memset(out, 0, 32);
    HostBits w = { out, 0 };
    w.put(0x4B414E5A, 32);
    w.put(6, 4);
    w.put(ckSize, 2);
    w.put((u64)eType, 5);
    w.put(tType, 48);
    w.put((u64)(blockSize >> 4), 28);
    int szMask = 0;
    if (inputSize != 0 && inputSize < ((i64)1 << 48)) {
        int lg = 0;
        for (u64 x = (u64)inputSize; x > 1; x >>= 1)
            lg++;
        szMask = (lg >> 4) + 1;
    }
    w.put((u64)szMask, 2);
    if (szMask)
        w.put((u64)inputSize, 16 * szMask);
    w.put(0, 15);
    const u32 HASH = 0x1E35A7BDu;
    u32 ck = HASH * (0x01030507u * 6u);
    ck ^= HASH * (u32)~ckSize;
    ck ^= HASH * (u32)~(u32)eType;
    ck ^= HASH * (u32)((~tType) >> 32);
    ck ^= HASH * (u32)(~tType);
    ck ^= HASH * (u32)~(u32)blockSize;
    if (szMask) {
        ck ^= HASH * (u32)((~(u64)inputSize) >> 32);
        ck ^= HASH * (u32)(~(u64)inputSize);
    }
    ck = (ck >> 23) ^ (ck >> 3);
    w.put(ck & 0xFFFFFFu, 24);
    return (int)(w.pos >> 3);
}

extern "C" int knz_stream_header(uint64_t tType, int eType, int blockSize, int64_t inputSize, uint8_t out[32])
{
    return knz_stream_header_ex(tType, eType, blockSize, inputSize, 0, out);
}

int knz_grow(knz_ctx* ctx, u8** buf, i64* cap, i64 need)
{
    if (*cap >= need)
        return KNZ_OK;
    if (*buf)
        cudaFree(*buf);
    *buf = NULL;
    *cap = 0;
    CK(cudaMalloc((void**)buf, (size_t)need));
    *cap = need;
    return KNZ_OK;
}

// ---- host stages in front of / behind the device stages (pre.cu) ------------------------------------
// Leading host stages of a sequence; -1 when a host stage follows a device stage (not supported: the
// reference's levels put them first).
int knz_host_prefix_len(const int* types, int nt)
{
    int hs = 0;
    while (hs < nt && knz_is_host_stage(types[hs]))
        hs++;
    for (int i = hs; i < nt; i++)
        if (knz_is_host_stage(types[i]))
            return -1;
    return hs;
}

template <class F>
static void host_parallel_for(int n, F&& body)
{
    int nt = (int)std::thread::hardware_concurrency();
    nt = (nt < 1) ? 1 : (nt > 64 ? 64 : nt);
    if (nt > n)
        nt = n;
    std::atomic<int> next(0);
    auto run = [&]() {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1))
            body(i);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++)
        pool.emplace_back(run);
    run();
    for (std::thread& t : pool)
        t.join();
}

int knz_ensure_h_pre(knz_ctx* ctx)
{
    if (ctx->h_pre == NULL && cudaMallocHost((void**)&ctx->h_pre, (size_t)ctx->maxBatch * (size_t)ctx->bstride) != cudaSuccess) {
        ctx->h_pre = NULL;
        snprintf(ctx->err, sizeof(ctx->err), "cannot allocate the host staging buffer");
        return KNZ_ERR_PROCESS_BLOCK;
    }
    return KNZ_OK;
}

// Forward: every block of a batch through the leading host stages, one block per task (what the reference's
// worker threads do in TransformSequence::forward, transform/TransformSequence.hpp:88-162, for those stages).
// Leaves the bytes in ctx->h_pre (bstride apart) and the state of every block in ctx->h_init / h_dtype / h_hash.
int knz_host_prefix_forward(knz_ctx* ctx, const int* types, int hs, int eType, int blockSize, const u8* in, i64 inStride,
                            const int32_t* lens, int nb)
{
    if (knz_ensure_h_pre(ctx) != KNZ_OK)
        return KNZ_ERR_PROCESS_BLOCK;
    const int cap = ctx->stageCap;
    host_parallel_for(nb, [&](int b) {
        std::vector<u8> tmp[2];
        const u8* cur = in + (i64)b * inStride;
        int len = lens[b], swaps = 0, flags = 0xFF, w = 0;
        // what EncodingTask leaves in the Context before the sequence runs (io/CompressedOutputStream.cpp:722-731)
        KnzPreCtx pc = { knz_magic_data_type(cur, len), blockSize, eType };
        if (ctx->checksumBits)
            ctx->h_hash[b] = knz_xxhash_host(cur, len, ctx->checksumBits);
        for (int i = 0; i < hs; i++) {
            if (tmp[w].empty())
                tmp[w].resize((size_t)cap + 64);
            int ol = 0;
            if (knz_pre_forward(types[i], cur, len, tmp[w].data(), cap, &ol, &pc)) {
                cur = tmp[w].data();
                w ^= 1;
                len = ol;
                swaps++;
                flags &= ~(1 << (7 - i));
            }
        }
        memcpy(ctx->h_pre + (i64)b * ctx->bstride, cur, (size_t)len);
        ctx->h_init[b].len = len;
        ctx->h_init[b].cur = 2;
        ctx->h_init[b].swaps = swaps;
        ctx->h_init[b].flags = flags;
        ctx->h_dtype[b] = pc.dataType;
    });
    return KNZ_OK;
}

// Inverse: block b of a batch arrives from the device in ctx->h_pre (len[b] bytes) with the leading `hs` stages
// still to undo, last one first (TransformSequence::inverse, transform/TransformSequence.hpp:165-247);
// the result goes to out[b].  Returns false in ok[b] when a stage rejects its input.
void knz_host_prefix_inverse(knz_ctx* ctx, const int* types, int hs, int eType, int blockSize, const u8* flags,
                             const int* lens, int nb, u8* const* out, const int* outCap, int* outLen, u8* ok)
{
    const int cap = ctx->stageCap;
    const KnzPreCtx pc = { KDT_UNDEFINED, blockSize, eType };
    host_parallel_for(nb, [&](int b) {
        std::vector<u8> tmp[2];
        const u8* cur = ctx->h_pre + (i64)b * ctx->bstride;
        int len = lens[b], w = 0;
        ok[b] = 1;
        for (int i = hs - 1; i >= 0 && ok[b]; i--) {
            if (flags[b] & (1 << (7 - i)))
                continue; // skipped by the encoder
            if (tmp[w].empty())
                tmp[w].resize((size_t)cap + 64);
            int ol = 0;
            if (!knz_pre_inverse(types[i], cur, len, tmp[w].data(), cap, &ol, &pc)) {
                ok[b] = 0;
                break;
            }
            cur = tmp[w].data();
            w ^= 1;
            len = ol;
        }
        if (ok[b] && len > outCap[b])
            ok[b] = 0;
        if (ok[b])
            memcpy(out[b], cur, (size_t)len);
        outLen[b] = len;
    });
}

extern "C" int knz_compress(knz_ctx* ctx, const char* transform, const char* entropy, int blockSize,
                            const uint8_t* in, int64_t n, uint8_t* out, int64_t cap, int64_t* outLen)
{
    if (!ctx || !in || !out || !outLen || n < 0)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    const u64 tType = knz_transform_type(transform);
    const int eType = knz_entropy_type(entropy);
    if (tType == (u64)-1 || eType < 0) {
        snprintf(ctx->err, sizeof(ctx->err), "unsupported pipeline %s / %s", transform ? transform : "?",
                 entropy ? entropy : "?");
        return KNZ_ERR_INVALID_CODEC;
    }
    if (blockSize < 1024 || blockSize > ctx->maxBlockSize || (blockSize & 15))
        return KNZ_ERR_BLOCK_SIZE;
    int types[8];
    const int ntAll = knz_split_types(tType, types);
    const int hs = knz_host_prefix_len(types, ntAll);
    if (hs < 0 || (hs > 0 && ctx->skipBlocks)) {
        snprintf(ctx->err, sizeof(ctx->err), "host stages (PACK, DNA, MM, UTF) must lead the sequence and do not combine with skipBlocks");
        return KNZ_ERR_INVALID_CODEC;
    }
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    u8 hdr[32];
    const int hdrBytes = knz_stream_header_ex(tType, eType, blockSize, n, ctx->checksumBits, hdr);
    const i64 nBlocks = (n + blockSize - 1) / blockSize;
    // Worst case per block: a transform sequence may legally expand a block up to the reference's task
    // buffer, max(bs + bs/8, 256 KiB) (ZRLT at odd swap parity on 0xFF-heavy data); the entropy stage adds at
    // most 25 % plus, under ANS1, the context headers of every order-1 chunk.
    const i64 refCapBlk = ((i64)blockSize + (blockSize >> 3) > 262144) ? (i64)blockSize + (blockSize >> 3) : 262144;
    const i64 worstBlk = ((2 * (i64)blockSize < refCapBlk) ? 2 * (i64)blockSize : refCapBlk);
    const i64 perBlk = worstBlk + (worstBlk >> 2) + 1024 + ((eType == E_ANS1) ? 131072 * (i64)((blockSize >> 22) + 1) : 0);
    const i64 streamCap = round_up(nBlocks * perBlk + 65536, 256);
    int rc = knz_grow(ctx, &ctx->dStream, &ctx->dStreamCap, streamCap);
    if (rc != KNZ_OK)
        return rc;
    // Encode in sub-batches: the pinned-host -> device copy of sub-batch i+1 runs on the copy
    // stream while sub-batch i is being encoded, and the finished bytes of the stream go back
    // to the host on a third stream.  The first sub-batch is small (its copy is exposed),
    // later ones are large (GPU efficiency, fewer round trips).
    const int* sched = ctx->encSched; // KNZ_ENC_BATCH=a,b,c overrides the schedule (read once per context)
    auto batchSize = [&](int k, i64 left) -> int {
        const int want = sched[k < 2 ? k : 2];
        const int lim = (want < ctx->maxBatch) ? want : ctx->maxBatch;
        return (int)((left < lim) ? left : lim);
    };
    const int ebWant = (sched[1] > sched[2]) ? (sched[1] > sched[0] ? sched[1] : sched[0]) : (sched[2] > sched[0] ? sched[2] : sched[0]);
    const int eb = (ctx->maxBatch < ebWant) ? ctx->maxBatch : ebWant;
    const i64 plainBytes = round_up((i64)eb * blockSize + 256, 256);
    rc = knz_grow(ctx, &ctx->dPlain, &ctx->dPlainCap, plainBytes);
    if (rc != KNZ_OK)
        return rc;
    rc = knz_grow(ctx, &ctx->dPlain2, &ctx->dPlain2Cap, plainBytes);
    if (rc != KNZ_OK)
        return rc;
    u8* plain[2] = { ctx->dPlain, ctx->dPlain2 };
    CK(cudaMemsetAsync(ctx->dStream, 0, (size_t)streamCap, s));
    ctx->h_pos[0] = 8ull * (u64)hdrBytes;
    CK(cudaMemcpyAsync(ctx->streamPos, ctx->h_pos, sizeof(u64), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    const int firstLen = (int)((n < blockSize) ? n : blockSize);
    float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    auto issueCopy = [&](i64 b0, int nb, int slot) -> cudaError_t {
        const i64 off = b0 * blockSize;
        const i64 bytes = ((off + (i64)nb * blockSize) <= n) ? (i64)nb * blockSize : n - off;
        cudaError_t e = cudaMemcpyAsync(plain[slot], in + off, (size_t)bytes, cudaMemcpyHostToDevice, ctx->copyStream);
        if (e == cudaSuccess)
            e = cudaEventRecord(ctx->evCopy[slot], ctx->copyStream);
        return e;
    };
    int nbCur = batchSize(0, nBlocks);
    if (nBlocks > 0 && hs == 0)
        CK(issueCopy(0, nbCur, 0));
    int slot = 0;
    i64 sentBytes = 0; // bytes of the stream already on their way to the host
    for (i64 b0 = 0, k = 0; b0 < nBlocks; k++, slot ^= 1) {
        const int nb = nbCur;
        const i64 off = b0 * blockSize;
        const i64 bNext = b0 + nb;
        // the other staging buffer is free: encode_batch of the previous sub-batch has completed
        if (bNext < nBlocks) {
            nbCur = batchSize((int)k + 1, nBlocks - bNext);
            if (hs == 0)
                CK(issueCopy(bNext, nbCur, slot ^ 1));
        }
        if (hs == 0)
            CK(cudaStreamWaitEvent(s, ctx->evCopy[slot], 0));
        int32_t* lens = (int32_t*)malloc(sizeof(int32_t) * (size_t)nb);
        for (int i = 0; i < nb; i++) {
            const i64 rem = n - (off + (i64)i * blockSize);
            lens[i] = (int)((rem < blockSize) ? rem : blockSize);
        }
        int ng = nb;
        if (lens[nb - 1] <= 15)
            ng = nb - 1; // only the last block of a stream can be that small
        if (ng > 0 && hs > 0) {
            // leading host stages on the host threads, then what they left goes to the device in one copy per block
            rc = knz_host_prefix_forward(ctx, types, hs, eType, blockSize, in + off, blockSize, lens, ng);
            for (int b = 0; b < ng && rc == KNZ_OK; b++)
                if (cudaMemcpyAsync(ctx->dStageIn + (i64)b * ctx->bstride, ctx->h_pre + (i64)b * ctx->bstride,
                                    (size_t)ctx->h_init[b].len, cudaMemcpyHostToDevice, s) != cudaSuccess)
                    rc = KNZ_ERR_PROCESS_BLOCK;
            if (rc == KNZ_OK) {
                ctx->nHost = hs;
                rc = knz_encode_batch(ctx, tType, eType, blockSize, ctx->dStageIn, ctx->bstride, lens, ng, firstLen, ctx->dOut,
                                      ctx->outStride, ctx->blockBits, NULL);
                ctx->nHost = 0;
            }
            for (int i = 0; i < 8; i++)
                acc[i] += ctx->ms[i];
        } else if (ng > 0) {
            rc = knz_encode_batch(ctx, tType, eType, blockSize, plain[slot], blockSize, lens, ng, firstLen, ctx->dOut,
                              ctx->outStride, ctx->blockBits, NULL);
            for (int i = 0; i < 8; i++)
                acc[i] += ctx->ms[i];
        }
        if (rc == KNZ_OK && ng < nb) {
            u8 tmp[32];
            memset(tmp, 0, sizeof(tmp));
            ctx->h_bits[ng] = knz_frame_small_block(in + off + (i64)ng * blockSize, lens[ng], tmp, ctx->checksumBits);
            CK(cudaMemcpyAsync(ctx->dOut + (i64)ng * ctx->outStride, tmp, 32, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(ctx->blockBits + ng, ctx->h_bits + ng, sizeof(u64), cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));
            ctx->h_st[ng].len = lens[ng];
            ctx->h_st[ng].flags = 0x7F;
            if (ctx->checksumBits)
                ctx->h_hash[ng] = knz_xxhash_host(in + off + (i64)ng * blockSize, lens[ng], ctx->checksumBits);
        }
        if (rc == KNZ_OK && ctx->listener) {
            // Event re-emission (io/CompressedOutputStream.cpp:685-689, :768-772, :810-814, :871-881): the batch
            // has completed, so the events of a block come out together, blocks in stream order
            u64 pos = ctx->h_pos[0]; // bit position of the batch in the stream (OutputBitStream::tell)
            for (int g = 0; g < nb; g++) {
                const int id = (int)(b0 + g) + 1;
                const u64 wr = ctx->h_bits[g];
                const int post = ctx->h_st[g].len;
                const u64 hsh = ctx->checksumBits ? ctx->h_hash[g] : 0;
                emit_event(ctx, KNZ_EVT_BEFORE_TRANSFORM, id, lens[g], hsh, ctx->checksumBits, -1, 0);
                emit_event(ctx, KNZ_EVT_AFTER_TRANSFORM, id, post, hsh, ctx->checksumBits, -1, 0);
                emit_event(ctx, KNZ_EVT_BEFORE_ENTROPY, id, post, hsh, ctx->checksumBits, -1, 0);
                emit_event(ctx, KNZ_EVT_AFTER_ENTROPY, id, (i64)((wr + 7) >> 3), hsh, ctx->checksumBits, -1, 0);
                emit_event(ctx, KNZ_EVT_BLOCK_INFO, id, (i64)((wr + 7) >> 3), hsh, ctx->checksumBits, (i64)pos,
                           ctx->h_st[g].flags & 0xFF);
                const u32 lw = (wr < 8) ? 3u : (u32)ilog2_u32((u32)(wr >> 3)) + 4u;
                pos += 5 + lw + wr;
            }
        }
        free(lens);
        if (rc != KNZ_OK) {
            cudaStreamSynchronize(ctx->d2hStream);
            return rc;
        }
        CK(cudaEventRecord(ctx->ev[5], s));
        launch_stream_assemble(ctx->dOut, ctx->outStride, ctx->blockBits, nb, ctx->streamPos, ctx->blockOff,
                               ctx->streamPos, ctx->dStream, s, &ctx->launches);
        CK(cudaEventRecord(ctx->ev[6], s));
        CK(cudaMemcpyAsync(ctx->h_pos, ctx->streamPos, sizeof(u64), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]);
        acc[4] += ms;
        // whole bytes assembled so far are final: send them while the next sub-batch encodes
        const i64 fullBytes = (i64)(ctx->h_pos[0] >> 3);
        if (fullBytes > cap || fullBytes > streamCap) {
            cudaStreamSynchronize(ctx->d2hStream);
            return KNZ_ERR_OUTPUT_TOO_SMALL;
        }
        if (bNext < nBlocks && fullBytes > sentBytes) {
            CK(cudaMemcpyAsync(out + sentBytes, ctx->dStream + sentBytes, (size_t)(fullBytes - sentBytes),
                               cudaMemcpyDeviceToHost, ctx->d2hStream));
            sentBytes = fullBytes;
        }
        b0 = bNext;
    }
    CK(cudaGetLastError());
    const u64 endBit = ctx->h_pos[0] + 8; // end marker: 5 + 3 zero bits (:416-417)
    const i64 total = (i64)((endBit + 7) >> 3);
    if (total > cap || total > streamCap) {
        cudaStreamSynchronize(ctx->d2hStream);
        return KNZ_ERR_OUTPUT_TOO_SMALL;
    }
    CK(cudaMemcpyAsync(out + sentBytes, ctx->dStream + sentBytes, (size_t)(total - sentBytes), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(ctx->d2hStream));
    memcpy(out, hdr, (size_t)hdrBytes);
    *outLen = total;
    for (int i = 0; i < 8; i++)
        ctx->ms[i] = acc[i];
    return KNZ_OK;
}

// ------------------------------------------------------------------ decode

// Inverse transforms + entropy decode for a batch.  Block b's bit string starts
// at bit h_start[b] of d_in (+ b*inStride) and holds h_bits[b] bits.
int knz_decode_batch(knz_ctx* ctx, u64 tType, int eType, int blockSize, const u8* d_in, i64 inStride,
                        const u64* h_payStart, const u64* h_endBit, const int* h_preLen, const u8* h_flags, int nB,
                        u8* d_out, i64 outStride, int32_t* h_outLens,
                        u8* h_sink, int* h_sinkBlocks, const u64* h_expectHash, int ckBits)
{
    int types[8];
    const int nt = knz_split_types(tType, types);
    const int hs = ctx->nHost; // leading stages the caller undoes on the host afterwards (pre.cu)
    for (int i = 0; i < nt; i++)
        if (!(type_supported(types[i]) || (i < hs && knz_is_host_stage(types[i])))) {
            snprintf(ctx->err, sizeof(ctx->err), "transform id %d not implemented%s", types[i],
                     knz_is_host_stage(types[i]) ? " here (host stages: leading stages of knz_compress / knz_decompress only)" : "");
            return KNZ_ERR_INVALID_CODEC;
        }
    if (!entropy_supported(eType))
        return KNZ_ERR_INVALID_CODEC;
    {
        int rc1 = ensure_ans1(ctx, eType);
        if (rc1 == KNZ_OK)
            rc1 = ensure_bwt(ctx, types + hs, nt - hs);
        if (rc1 != KNZ_OK)
            return rc1;
    }
    if (nB > ctx->maxBatch)
        return KNZ_ERR_INVALID_PARAM;
    if (blockSize < 1 || blockSize > ctx->maxBlockSize)
        return KNZ_ERR_BLOCK_SIZE;
    cudaStream_t s = ctx->stream;
    // a decoded block may not exceed its destination slot nor the stream's block size (what the host stages
    // still have to undo may be as large as a stage buffer)
    const int outCap = (hs > 0) ? (int)((outStride < ctx->stageCap) ? outStride : ctx->stageCap)
                                : (int)((outStride > 0 && outStride < blockSize) ? outStride : blockSize);
    // capacity of the reference's task buffers on the decode side (io/CompressedInputStream.cpp:275)
    const int blkLen = blockSize + ((blockSize >> 4) > 512 ? (blockSize >> 4) : 512);
    int maxLen = 0;
    for (int b = 0; b < nB; b++) {
        // entropy-decoded length: whatever a stage buffer slot holds (the encoder emits <= max(bs + bs/8, 256 KiB))
        if (h_preLen[b] <= 0 || h_preLen[b] > ctx->stageCap)
            return KNZ_ERR_INVALID_FILE;
        ctx->h_st[b].len = h_preLen[b];
        ctx->h_st[b].cur = 0;
        ctx->h_st[b].swaps = 0;
        ctx->h_st[b].flags = h_flags[b];
        ctx->h_preLen[b] = h_preLen[b];
        ctx->h_payStart[b] = h_payStart[b];
        ctx->h_bits[b] = h_endBit[b];
        ctx->h_capEven[b] = blkLen;
        ctx->h_capOdd[b] = blkLen;
        if (h_preLen[b] > maxLen)
            maxLen = h_preLen[b];
    }
    CK(cudaMemcpyAsync(ctx->st, ctx->h_st, sizeof(BlkState) * nB, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dPreLen, ctx->h_preLen, sizeof(int) * nB, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dPayStart, ctx->h_payStart, sizeof(u64) * nB, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dInBits, ctx->h_bits, sizeof(u64) * nB, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->capEven, ctx->h_capEven, sizeof(int) * nB, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->capOdd, ctx->h_capOdd, sizeof(int) * nB, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->errFlag, 0, sizeof(int) * 4, s));
    for (int i = 0; i < 8; i++)
        ctx->ms[i] = 0.f;
    CK(cudaEventRecord(ctx->ev[0], s));
    DecodeLaunch D;
    D.in = d_in;
    D.inStride = inStride;
    D.inBits = ctx->dInBits;
    D.payStart = ctx->dPayStart;
    D.preLen = ctx->dPreLen;
    D.nBlocks = nB;
    D.maxChunks = ctx->maxChunks;
    D.eType = eType;
    D.chunkPos = ctx->chunkPos;
    D.dst = ctx->bufA;
    D.dstStride = ctx->bstride;
    D.errFlag = ctx->errFlag;
    D.evK0 = ctx->ev[8];
    D.evK1 = ctx->ev[9];
    D.a1 = &ctx->a1;
    BufTable bt;
    bt.base[0] = ctx->bufA;
    bt.base[1] = ctx->bufB;
    bt.base[2] = ctx->bufA;
    bt.stride[0] = bt.stride[1] = bt.stride[2] = ctx->bstride;

    // Overlapped mode: the batch is cut into block groups, each running the whole inverse chain on
    // its own stream over its own slice of the scratch arrays.  The inverse RANK/MTFT stage is one
    // dependency chain per block (one warp per block, ~140 ms for 4 MiB whatever the batch size):
    // while one group sits in it the SMs run the entropy / ZRLT / BWT stages of the other groups.
    const int G = (hs == 0 && ctx->decGroups > 1 && nB >= 4 * ctx->decGroups) ? ctx->decGroups : 1;
    if (G > 1) {
        const bool sink = (h_sink != NULL) && (outStride == blockSize);
        CK(cudaEventRecord(ctx->gEv[KNZ_MAX_GROUPS], s));
        for (int g = 0; g < G; g++) {
            const int g0 = (int)((i64)nB * g / G), g1 = (int)((i64)nB * (g + 1) / G);
            cudaStream_t sg = ctx->gStream[g];
            CK(cudaStreamWaitEvent(sg, ctx->gEv[KNZ_MAX_GROUPS], 0));
            DecodeLaunch Dg = D;
            Dg.in = d_in + (i64)g0 * inStride;
            Dg.inBits = ctx->dInBits + g0;
            Dg.payStart = ctx->dPayStart + g0;
            Dg.preLen = ctx->dPreLen + g0;
            Dg.nBlocks = g1 - g0;
            Dg.chunkPos = ctx->chunkPos + (i64)g0 * ctx->maxChunks;
            Dg.dst = ctx->bufA + (i64)g0 * ctx->bstride;
            Dg.evK0 = NULL;
            Dg.evK1 = NULL;
            launch_entropy_decode(Dg, sg, &ctx->launches);
            int stepg = 0;
            for (int i = nt - 1; i >= 0; i--, stepg++) {
                StageLaunch L;
                L.bt = bt;
                for (int k = 0; k < 3; k++)
                    L.bt.base[k] = bt.base[k] + (i64)g0 * bt.stride[k];
                L.stIn = ctx->st + (i64)stepg * ctx->maxBatch + g0;
                L.stOut = ctx->st + (i64)(stepg + 1) * ctx->maxBatch + g0;
                L.stageIdx = i;
                L.nBlocks = g1 - g0;
                L.maxLen = (maxLen > blkLen) ? maxLen : blkLen;
                L.capEven = ctx->capEven + g0;
                L.capOdd = ctx->capOdd + g0;
                L.errFlag = ctx->errFlag;
                L.wsBlock0 = g0;
                launch_inverse_stage(ctx, types[i], L, sg);
            }
            {
                BufTable btg = bt;
                for (int k = 0; k < 3; k++)
                    btg.base[k] = bt.base[k] + (i64)g0 * bt.stride[k];
                launch_copy_out(btg, ctx->st + (i64)nt * ctx->maxBatch + g0, g1 - g0, d_out + (i64)g0 * outStride,
                                outStride, outCap, ctx->errFlag, sg, &ctx->launches);
            }
            if (sink) { // full-size blocks go to the host as soon as their group is done
                const int last = (g1 == nB) ? g1 - 1 : g1; // the batch's last block may be short: the caller copies it
                if (last > g0)
                    CK(cudaMemcpyAsync(h_sink + (i64)g0 * blockSize, d_out + (i64)g0 * outStride,
                                       (size_t)(last - g0) * (size_t)blockSize, cudaMemcpyDeviceToHost, sg));
            }
            CK(cudaEventRecord(ctx->gEv[g], sg));
            CK(cudaStreamWaitEvent(s, ctx->gEv[g], 0));
        }
        if (sink)
            *h_sinkBlocks = nB - 1;
        const BlkState* stF = ctx->st + (i64)nt * ctx->maxBatch;
        if (ckBits && h_expectHash) { // DecodingTask::run verifies the hash of the decoded block (:1003-1022)
            memcpy(ctx->h_hash, h_expectHash, sizeof(u64) * (size_t)nB);
            CK(cudaMemcpyAsync(ctx->expectHash, ctx->h_hash, sizeof(u64) * nB, cudaMemcpyHostToDevice, s));
            launch_xxhash(bt, stF, nB, ckBits, NULL, ctx->expectHash, ctx->errFlag, s, &ctx->launches);
        }
        CK(cudaEventRecord(ctx->ev[4], s));
        CK(cudaMemcpyAsync(ctx->h_err, ctx->errFlag, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->h_st, stF, sizeof(BlkState) * nB, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        float msT = 0.f;
        cudaEventElapsedTime(&msT, ctx->ev[0], ctx->ev[4]);
        ctx->ms[5] = msT; // per-stage times are not separable when the groups overlap
        for (int b = 0; b < nB; b++)
            h_outLens[b] = ctx->h_st[b].len;
        return map_kerr(ctx, ctx->h_err[0] ? ctx->h_err[0] : ctx->h_err[1] ? ctx->h_err[1] : ctx->h_err[2]);
    }

    launch_entropy_decode(D, s, &ctx->launches);
    CK(cudaEventRecord(ctx->ev[1], s));
    int step = 0;
    for (int i = nt - 1; i >= hs; i--, step++) {
        StageLaunch L;
        L.bt = bt;
        L.stIn = ctx->st + (i64)step * ctx->maxBatch;
        L.stOut = ctx->st + (i64)(step + 1) * ctx->maxBatch;
        L.stageIdx = i;
        L.nBlocks = nB;
        L.maxLen = (maxLen > blkLen) ? maxLen : blkLen;
        L.capEven = ctx->capEven;
        L.capOdd = ctx->capOdd;
        L.errFlag = ctx->errFlag;
        L.wsBlock0 = 0;
        CK(cudaEventRecord(ctx->evStage[2 * step], s));
        // Last stage of a full batch with a host sink: inverse-BWT the blocks in four groups and
        // send each group's (full-size) blocks to the host while the next group is being walked.
        const bool grouped = (h_sink != NULL) && (i == 0) && (types[0] == T_BWT) && (nB >= 64) && (outStride == blockSize);
        if (grouped) {
            const int G = (ctx->decBwtGroups < nB / 8) ? ctx->decBwtGroups : nB / 8;
            for (int g = 0; g < G; g++) {
                const int g0 = (int)((i64)nB * g / G), g1 = (int)((i64)nB * (g + 1) / G);
                StageLaunch Lg = L;
                for (int k = 0; k < 3; k++)
                    Lg.bt.base[k] = L.bt.base[k] + (i64)g0 * L.bt.stride[k];
                Lg.stIn = L.stIn + g0;
                Lg.stOut = L.stOut + g0;
                Lg.capEven = L.capEven + g0;
                Lg.capOdd = L.capOdd + g0;
                Lg.nBlocks = g1 - g0;
                launch_bwt_inverse(Lg, ctx->ws, s, &ctx->launches);
                launch_copy_out(Lg.bt, Lg.stOut, g1 - g0, d_out + (i64)g0 * outStride, outStride, outCap, ctx->errFlag, s,
                                &ctx->launches);
                CK(cudaEventRecord(ctx->evDone[g & 1], s));
                CK(cudaStreamWaitEvent(ctx->d2hStream, ctx->evDone[g & 1], 0));
                const int last = (g1 == nB) ? g1 - 1 : g1; // the batch's last block may be short: the caller copies it
                if (last > g0)
                    CK(cudaMemcpyAsync(h_sink + (i64)g0 * blockSize, d_out + (i64)g0 * outStride,
                                       (size_t)(last - g0) * (size_t)blockSize, cudaMemcpyDeviceToHost, ctx->d2hStream));
            }
            *h_sinkBlocks = nB - 1;
            CK(cudaEventRecord(ctx->evStage[2 * step + 1], s));
            continue;
        }
        launch_inverse_stage(ctx, types[i], L, s);
        CK(cudaEventRecord(ctx->evStage[2 * step + 1], s));
    }
    const BlkState* stFinal = ctx->st + (i64)(nt - hs) * ctx->maxBatch;
    if (ckBits && h_expectHash) { // DecodingTask::run verifies the hash of the decoded block (:1003-1022)
        memcpy(ctx->h_hash, h_expectHash, sizeof(u64) * (size_t)nB);
        CK(cudaMemcpyAsync(ctx->expectHash, ctx->h_hash, sizeof(u64) * nB, cudaMemcpyHostToDevice, s));
        launch_xxhash(bt, stFinal, nB, ckBits, NULL, ctx->expectHash, ctx->errFlag, s, &ctx->launches);
    }
    if (!(h_sinkBlocks && *h_sinkBlocks > 0))
        launch_copy_out(bt, stFinal, nB, d_out, outStride, outCap, ctx->errFlag, s, &ctx->launches);
    CK(cudaEventRecord(ctx->ev[4], s));
    CK(cudaMemcpyAsync(ctx->h_err, ctx->errFlag, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_st, stFinal, sizeof(BlkState) * nB, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    float ms = 0.f;
    for (int k = 0; k < nt - hs; k++) { // stage brackets were recorded without stalling the stream
        cudaEventElapsedTime(&ms, ctx->evStage[2 * k], ctx->evStage[2 * k + 1]);
        add_stage_time(ctx, types[nt - 1 - k], ms);
    }
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->ms[3] = ms;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[4]);
    ctx->ms[5] = ms;
    if (eType != E_RAW) {
        cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]);
        ctx->ms[7] = ms;
    }
    for (int b = 0; b < nB; b++)
        h_outLens[b] = ctx->h_st[b].len;
    return map_kerr(ctx, ctx->h_err[0] ? ctx->h_err[0] : ctx->h_err[1] ? ctx->h_err[1] : ctx->h_err[2]);
}

// Parse one block's private header (mode byte, [skip flags], length) at r.pos.
// Returns 0 ok, 1 copy block, <0 error.
int knz_parse_block_header(HostBitReader& r, int blockSize, u8* flags, int* preLen, int ckBits, u64* checksum)
{
    const int mode = (int)r.get(8);
    int fl = 0;
    int copy = 0;
    if (mode & 0x80)
        copy = 1;
    else if (mode & 0x10)
        fl = (int)r.get(8);
    else
        fl = ((mode << 4) | 0x0F) & 0xFF;
    const int dataSize = 1 + ((mode >> 5) & 3);
    const int pre = (int)r.get(8 * dataSize);
    int maxT = blockSize + blockSize / 2; // io/CompressedInputStream.cpp:893-894 (maxTransformSize)
    if (maxT < 2048)
        maxT = 2048;
    if (r.bad || pre <= 0 || pre > maxT)
        return -1;
    if (ckBits) { // io/CompressedInputStream.cpp:913-921
        const u64 ck = (ckBits == 32) ? r.get(32) : ((r.get(32) << 32) | r.get(32));
        if (checksum)
            *checksum = ck;
        if (r.bad)
            return -1;
    }
    *flags = (u8)fl;
    *preLen = pre;
    return copy;
}

extern "C" int knz_decode_blocks_dev(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* d_in,
                                     int64_t inStride, const uint64_t* h_inBits, int nBlocks, uint8_t* d_out,
                                     int64_t outStride, int32_t* h_outLens)
{
    // The block headers live in device memory here: fetch the first 8 bytes of each block.
    if (!ctx || !d_in || !h_inBits || !d_out || !h_outLens)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int off = 0; off < nBlocks; off += ctx->maxBatch) {
        const int nb = (nBlocks - off < ctx->maxBatch) ? nBlocks - off : ctx->maxBatch;
        u8* heads = (u8*)malloc((size_t)nb * 16);
        u64* cks = (u64*)malloc(sizeof(u64) * (size_t)nb);
        CK(cudaMemcpy2DAsync(heads, 16, d_in + (i64)off * inStride, (size_t)inStride, 16, (size_t)nb,
                             cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        u64* pay = (u64*)malloc(sizeof(u64) * (size_t)nb);
        u64* endb = (u64*)malloc(sizeof(u64) * (size_t)nb);
        int* pre = (int*)malloc(sizeof(int) * (size_t)nb);
        u8* fl = (u8*)malloc((size_t)nb);
        int rc = KNZ_OK;
        for (int b = 0; b < nb && rc == KNZ_OK; b++) {
            HostBitReader r = { heads + 16 * b, 128, 0, false };
            const int k = knz_parse_block_header(r, blockSize, &fl[b], &pre[b], ctx->checksumBits, &cks[b]);
            if (k != 0)
                rc = (k < 0) ? KNZ_ERR_INVALID_FILE : KNZ_ERR_INVALID_CODEC; // device path: no copy blocks
            pay[b] = r.pos;
            endb[b] = h_inBits[off + b];
        }
        if (rc == KNZ_OK)
            rc = knz_decode_batch(ctx, tType, eType, blockSize, d_in + (i64)off * inStride, inStride, pay, endb, pre, fl, nb,
                              d_out + (i64)off * outStride, outStride, h_outLens + off, NULL, NULL, cks,
                              ctx->checksumBits);
        free(heads);
        free(cks);
        free(pay);
        free(endb);
        free(pre);
        free(fl);
        if (rc != KNZ_OK)
            return rc;
        for (int i = 0; i < 8; i++)
            acc[i] += ctx->ms[i];
    }
    for (int i = 0; i < 8; i++)
        ctx->ms[i] = acc[i];
    return KNZ_OK;
}

extern "C" int knz_decode_blocks(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* in,
                                 int64_t inStride, const uint64_t* inBits, int nBlocks, uint8_t* out,
                                 int64_t outStride, int32_t* outLens)
{
    if (!ctx || !in || !inBits || !out || !outLens)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    for (int off = 0; off < nBlocks; off += ctx->maxBatch) {
        const int nb = (nBlocks - off < ctx->maxBatch) ? nBlocks - off : ctx->maxBatch;
        u64* pay = (u64*)malloc(sizeof(u64) * (size_t)nb);
        u64* endb = (u64*)malloc(sizeof(u64) * (size_t)nb);
        int* pre = (int*)malloc(sizeof(int) * (size_t)nb);
        u8* fl = (u8*)malloc((size_t)nb);
        int* map = (int*)malloc(sizeof(int) * (size_t)nb);
        u64* cks = (u64*)malloc(sizeof(u64) * (size_t)nb);
        const int ckBits = ctx->checksumBits;
        int ng = 0, rc = KNZ_OK;
        for (int b = 0; b < nb && rc == KNZ_OK; b++) {
            const u8* p = in + (i64)(off + b) * inStride;
            const i64 nbytes = (i64)((inBits[off + b] + 7) >> 3);
            HostBitReader r = { p, inBits[off + b], 0, false };
            u8 f = 0;
            int pl = 0;
            u64 ck = 0;
            const int k = knz_parse_block_header(r, blockSize, &f, &pl, ckBits, &ck);
            if (k < 0 || nbytes + 16 > ctx->outStride) {
                rc = KNZ_ERR_INVALID_FILE;
            } else if (k == 1) { // copy block: raw bytes follow
                if (r.pos + 8ull * (u64)pl > inBits[off + b] || pl > outStride) {
                    rc = KNZ_ERR_INVALID_FILE;
                } else {
                    memcpy(out + (i64)(off + b) * outStride, p + (r.pos >> 3), (size_t)pl);
                    outLens[off + b] = pl;
                    if (ckBits && knz_xxhash_host(out + (i64)(off + b) * outStride, pl, ckBits) != ck)
                        rc = KNZ_ERR_CRC_CHECK;
                }
            } else {
                CK(cudaMemcpyAsync(ctx->dOut + (i64)ng * ctx->outStride, p, (size_t)nbytes, cudaMemcpyHostToDevice, s));
                pay[ng] = r.pos;
                endb[ng] = inBits[off + b];
                pre[ng] = pl;
                fl[ng] = f;
                cks[ng] = ck;
                map[ng] = off + b;
                ng++;
            }
        }
        int32_t* ol = (int32_t*)malloc(sizeof(int32_t) * (size_t)nb);
        if (rc == KNZ_OK && ng > 0)
            rc = knz_decode_batch(ctx, tType, eType, blockSize, ctx->dOut, ctx->outStride, pay, endb, pre, fl, ng,
                              ctx->dStageIn, ctx->bstride, ol, NULL, NULL, cks, ckBits);
        if (rc == KNZ_OK) {
            for (int g = 0; g < ng; g++) {
                if (ol[g] > outStride) {
                    rc = KNZ_ERR_OUTPUT_TOO_SMALL;
                    break;
                }
                cudaMemcpyAsync(out + (i64)map[g] * outStride, ctx->dStageIn + (i64)g * ctx->bstride, (size_t)ol[g],
                                cudaMemcpyDeviceToHost, s);
                outLens[map[g]] = ol[g];
            }
            cudaStreamSynchronize(s);
        }
        free(ol);
        free(pay);
        free(endb);
        free(pre);
        free(fl);
        free(map);
        free(cks);
        if (rc != KNZ_OK)
            return rc;
    }
    return KNZ_OK;
}

// Stream header: magic, version 6, checksum size, entropy id, transform word, block size, optional
// original size, padding, header checksum (io/CompressedInputStream.cpp:511-663).
int knz_parse_stream_header(knz_ctx* ctx, HostBitReader& r, KnzStreamInfo* info)
{
    if (r.get(32) != 0x4B414E5A)
        return KNZ_ERR_INVALID_FILE;
    if (r.get(4) != 6)
        return KNZ_ERR_STREAM_VERSION;
    const u32 ckSize = (u32)r.get(2);
    const int eType = (int)r.get(5);
    const u64 tType = r.get(48);
    const int blockSize = (int)r.get(28) << 4;
    const int szMask = (int)r.get(2);
    u64 origSize = 0;
    if (szMask)
        origSize = r.get(16 * szMask);
    r.get(15);
    const u32 ck1 = (u32)r.get(24);
    if (r.bad)
        return KNZ_ERR_INVALID_FILE;
    { // header checksum: the reference rejects a mismatch (io/CompressedInputStream.cpp:622-645)
        const u32 HASH = 0x1E35A7BDu;
        u32 ck = HASH * (0x01030507u * 6u);
        ck ^= HASH * (u32)~ckSize;
        ck ^= HASH * (u32)~(u32)eType;
        ck ^= HASH * (u32)((~tType) >> 32);
        ck ^= HASH * (u32)(~tType);
        ck ^= HASH * (u32)~(u32)blockSize;
        if (szMask) {
            ck ^= HASH * (u32)((~origSize) >> 32);
            ck ^= HASH * (u32)(~origSize);
        }
        ck = (ck >> 23) ^ (ck >> 3);
        if ((ck & 0xFFFFFFu) != ck1) {
            snprintf(ctx->err, sizeof(ctx->err), "invalid bitstream, header checksum mismatch");
            return KNZ_ERR_CRC_CHECK;
        }
    }
    if (ckSize > 2)
        return KNZ_ERR_INVALID_FILE;
    if (blockSize < 1024 || blockSize > ctx->maxBlockSize)
        return KNZ_ERR_BLOCK_SIZE;
    info->ckBits = (int)ckSize * 32;
    info->eType = eType;
    info->tType = tType;
    info->blockSize = blockSize;
    info->origSize = szMask ? (i64)origSize : -1;
    return KNZ_OK;
}

// Events of one decoded block (io/CompressedInputStream.cpp:924-932, :973-983, :380), in the order the
// reference's task and its consumer emit them.
static void emit_decode_events(knz_ctx* ctx, int id, i64 offset, i64 rBytes, int preLen, int decoded, u64 ck, int ckBits,
                               int flags)
{
    emit_event(ctx, KNZ_EVT_BLOCK_INFO, id, rBytes, ck, ckBits, offset, flags);
    emit_event(ctx, KNZ_EVT_BEFORE_ENTROPY, id, rBytes, ck, ckBits, -1, 0);
    emit_event(ctx, KNZ_EVT_AFTER_ENTROPY, id, preLen, ck, ckBits, -1, 0);
    emit_event(ctx, KNZ_EVT_BEFORE_TRANSFORM, id, preLen, ck, ckBits, -1, 0);
    emit_event(ctx, KNZ_EVT_AFTER_TRANSFORM, id, decoded, ck, ckBits, -1, 0);
}

static int decompress_impl(knz_ctx* ctx, const uint8_t* in, int64_t n, uint8_t* out, int64_t cap, int64_t* outLen,
                           int fromBlock, int toBlock, int64_t seekBit = -1)
{
    if (!ctx || !in || !out || !outLen || n < 20 || fromBlock < 1 || toBlock < fromBlock || seekBit >= 8 * n)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    HostBitReader r = { in, 8ull * (u64)n, 0, false };
    KnzStreamInfo info;
    {
        const int rch = knz_parse_stream_header(ctx, r, &info);
        if (rch != KNZ_OK)
            return rch;
    }
    const int eType = info.eType;
    const u64 tType = info.tType;
    const int blockSize = info.blockSize;
    if (seekBit >= 0) {
        // CompressedInputStream::seek (io/CompressedInputStream.hpp:329-375): decoding restarts at a block boundary
        // given as a bit position of the stream (what BLOCK_INFO events report); block ids count from there
        if ((u64)seekBit < r.pos)
            return KNZ_ERR_INVALID_PARAM;
        r.pos = (u64)seekBit;
    }
    int types[8];
    const int ntAll = knz_split_types(tType, types);
    const int hs = knz_host_prefix_len(types, ntAll);
    if (hs < 0) {
        snprintf(ctx->err, sizeof(ctx->err), "host stages behind device stages are not supported");
        return KNZ_ERR_INVALID_CODEC;
    }
    if (hs > 0 && knz_ensure_h_pre(ctx) != KNZ_OK)
        return KNZ_ERR_PROCESS_BLOCK;
    // whole compressed stream to the device; kernels read at bit offsets
    int rc = knz_grow(ctx, &ctx->dStream, &ctx->dStreamCap, round_up(n + 256, 256));
    if (rc != KNZ_OK)
        return rc;
    rc = knz_grow(ctx, &ctx->dPlain, &ctx->dPlainCap, round_up((i64)ctx->maxBatch * blockSize + 256, 256));
    if (rc != KNZ_OK)
        return rc;
    CK(cudaMemsetAsync(ctx->dStream + (n & ~(i64)255), 0, (size_t)(round_up(n + 256, 256) - (n & ~(i64)255)), s));
    // Block range (the `from` / `to` entries of the reference's context, io/CompressedInputStream.cpp:836-837,
    // :864-869: ids are 1-based, blocks with from <= id < to are decoded): only the bytes that hold those blocks
    // go to the device, found by walking the length prefixes first.
    i64 upLo = 0, upHi = n;
    if (fromBlock > 1 || toBlock != 0x7FFFFFFF) {
        HostBitReader w = r;
        u64 firstBit = ~0ull, lastBit = 0;
        for (int id = 1; id < toBlock; id++) {
            const int lr = 3 + (int)w.get(5);
            const u64 bits = w.get(lr);
            if (w.bad || bits == 0)
                break;
            if (id >= fromBlock) {
                if (firstBit == ~0ull)
                    firstBit = w.pos;
                lastBit = w.pos + bits;
            }
            w.pos += bits;
        }
        if (firstBit == ~0ull) {
            upLo = upHi = 0;
        } else {
            upLo = (i64)(firstBit >> 3) & ~(i64)255;
            upHi = (i64)((lastBit + 7) >> 3);
            if (upHi > n)
                upHi = n;
        }
    }
    if (upHi > upLo)
        CK(cudaMemcpyAsync(ctx->dStream + upLo, in + upLo, (size_t)(upHi - upLo), cudaMemcpyHostToDevice, s));
    const int mb = ctx->maxBatch;
    i64* evOff = (i64*)malloc(sizeof(i64) * (size_t)mb); // listener: bit position and byte size of each block
    i64* evBytes = (i64*)malloc(sizeof(i64) * (size_t)mb);
    int* evId = (int*)malloc(sizeof(int) * (size_t)mb);
    int nextId = 1; // id of the next block of the stream
    u64* pay = (u64*)malloc(sizeof(u64) * (size_t)mb);
    u64* endb = (u64*)malloc(sizeof(u64) * (size_t)mb);
    int* pre = (int*)malloc(sizeof(int) * (size_t)mb);
    u8* fl = (u8*)malloc((size_t)mb);
    int32_t* ol = (int32_t*)malloc(sizeof(int32_t) * (size_t)mb);
    u64* cks = (u64*)malloc(sizeof(u64) * (size_t)mb);
    const int ckBits = info.ckBits;
    i64 produced = 0;
    bool done = false;
    float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    while (!done && rc == KNZ_OK) {
        int ng = 0;
        // gather up to maxBatch device blocks; copy blocks are resolved on the host in order
        i64 batchOut = produced;
        while (ng < mb) {
            const int lr = 3 + (int)r.get(5);
            const u64 bits = r.get(lr);
            if (r.bad) {
                rc = KNZ_ERR_INVALID_FILE;
                break;
            }
            if (bits == 0 || nextId >= toBlock) {
                done = true;
                break;
            }
            const u64 start = r.pos;
            if (start + bits > r.nbits) {
                rc = KNZ_ERR_INVALID_FILE;
                break;
            }
            if (nextId < fromBlock) { // skipped: its bits are consumed, nothing is decoded (:864-866)
                r.pos = start + bits;
                nextId++;
                continue;
            }
            HostBitReader hb = { in, start + bits, start, false };
            u8 f = 0;
            int pl = 0;
            u64 ck = 0;
            const int k = knz_parse_block_header(hb, blockSize, &f, &pl, ckBits, &ck);
            if (k < 0 || start + bits > r.nbits) {
                rc = KNZ_ERR_INVALID_FILE;
                break;
            }
            if (k == 1) {
                if (ng > 0) { // keep output order simple: flush the device batch first
                    r.pos = start - (u64)(5 + lr);
                    break;
                }
                if (hb.pos + 8ull * (u64)pl > start + bits || produced + pl > cap) {
                    rc = KNZ_ERR_INVALID_FILE;
                    break;
                }
                for (int i = 0; i < pl; i++)
                    out[produced + i] = (u8)hb.get(8);
                if (ckBits && knz_xxhash_host(out + produced, pl, ckBits) != ck) {
                    snprintf(ctx->err, sizeof(ctx->err), "corrupted bitstream: block checksum mismatch");
                    rc = KNZ_ERR_CRC_CHECK;
                    break;
                }
                if (ctx->listener)
                    emit_decode_events(ctx, nextId, (i64)start - (5 + lr), (i64)((bits + 7) >> 3), pl, pl, ck, ckBits, 0);
                produced += pl;
                batchOut = produced;
                r.pos = start + bits;
                nextId++;
                continue;
            }
            pay[ng] = hb.pos;
            endb[ng] = start + bits;
            pre[ng] = pl;
            fl[ng] = f;
            cks[ng] = ck;
            evOff[ng] = (i64)start - (5 + lr);
            evBytes[ng] = (i64)((bits + 7) >> 3);
            evId[ng] = nextId;
            ng++;
            r.pos = start + bits;
            nextId++;
        }
        if (rc != KNZ_OK || ng == 0)
            continue;
        if (hs > 0) {
            // Device stages first (into stage-sized slots), then the leading host stages are undone on host threads
            // (TransformSequence::inverse walks the sequence backwards, transform/TransformSequence.hpp:165-247).
            ctx->nHost = hs;
            rc = knz_decode_batch(ctx, tType, eType, blockSize, ctx->dStream, 0, pay, endb, pre, fl, ng, ctx->dStageIn,
                                  ctx->bstride, ol, NULL, NULL, NULL, 0);
            ctx->nHost = 0;
            if (rc != KNZ_OK)
                break;
            for (int i = 0; i < 8; i++)
                acc[i] += ctx->ms[i];
            for (int g = 0; g < ng; g++)
                cudaMemcpyAsync(ctx->h_pre + (i64)g * ctx->bstride, ctx->dStageIn + (i64)g * ctx->bstride, (size_t)ol[g],
                                cudaMemcpyDeviceToHost, s);
            CK(cudaStreamSynchronize(s));
            std::vector<u8*> dstp((size_t)ng);
            std::vector<int> dcap((size_t)ng), dlen((size_t)ng), ilen(ol, ol + ng);
            std::vector<u8> good((size_t)ng);
            for (int g = 0; g < ng; g++) { // every block but the last of a stream is full: decode in place
                const i64 at = batchOut + (i64)g * blockSize;
                dstp[g] = out + ((at < cap) ? at : cap);
                const i64 room = cap - at;
                dcap[g] = (int)((room < 0) ? 0 : (room < blockSize ? room : blockSize));
            }
            knz_host_prefix_inverse(ctx, types, hs, eType, blockSize, fl, ilen.data(), ng, dstp.data(), dcap.data(),
                                    dlen.data(), good.data());
            for (int g = 0; g < ng && rc == KNZ_OK; g++) {
                if (!good[g]) {
                    snprintf(ctx->err, sizeof(ctx->err), "transform inverse failed (host stage) in block %d", evId[g]);
                    rc = (dlen[g] > dcap[g]) ? KNZ_ERR_OUTPUT_TOO_SMALL : KNZ_ERR_PROCESS_BLOCK;
                    break;
                }
                if (dstp[g] != out + batchOut) // a short block in the middle of the batch: close the gap
                    memmove(out + batchOut, dstp[g], (size_t)dlen[g]);
                if (ckBits && knz_xxhash_host(out + batchOut, dlen[g], ckBits) != cks[g]) {
                    snprintf(ctx->err, sizeof(ctx->err), "corrupted bitstream: block checksum mismatch");
                    rc = KNZ_ERR_CRC_CHECK;
                    break;
                }
                ol[g] = dlen[g];
                batchOut += dlen[g];
            }
            if (rc != KNZ_OK)
                break;
            produced = batchOut;
            if (ctx->listener)
                for (int g = 0; g < ng; g++)
                    emit_decode_events(ctx, evId[g], evOff[g], evBytes[g], pre[g], ol[g], cks[g], ckBits, fl[g]);
            continue;
        }
        // blocks the batch already sent to the host (full-size blocks, overlapped with the last stage)
        int sunk = 0;
        const bool roomy = batchOut + (i64)ng * blockSize <= cap;
        rc = knz_decode_batch(ctx, tType, eType, blockSize, ctx->dStream, 0, pay, endb, pre, fl, ng, ctx->dPlain, blockSize,
                          ol, roomy ? out + batchOut : NULL, &sunk, cks, ckBits);
        if (rc != KNZ_OK) {
            cudaStreamSynchronize(ctx->d2hStream);
            break;
        }
        for (int i = 0; i < 8; i++)
            acc[i] += ctx->ms[i];
        for (int g = 0; g < sunk; g++)
            if (ol[g] != blockSize)
                sunk = 0; // a short block in the middle: positions shift, copy everything again
        // decoded blocks are contiguous when every block but the last is full
        for (int g = 0; g < ng; g++) {
            if (ol[g] > blockSize || batchOut + ol[g] > cap) {
                rc = KNZ_ERR_OUTPUT_TOO_SMALL;
                break;
            }
            if (g == 0 && sunk == 0)
                cudaStreamSynchronize(ctx->d2hStream); // nothing of ours may land after the fresh copies
            if (g >= sunk)
                cudaMemcpyAsync(out + batchOut, ctx->dPlain + (i64)g * blockSize, (size_t)ol[g], cudaMemcpyDeviceToHost, s);
            batchOut += ol[g];
        }
        cudaStreamSynchronize(s);
        cudaStreamSynchronize(ctx->d2hStream);
        produced = batchOut;
        if (rc == KNZ_OK && ctx->listener)
            for (int g = 0; g < ng; g++)
                emit_decode_events(ctx, evId[g], evOff[g], evBytes[g], pre[g], ol[g], cks[g], ckBits, fl[g]);
    }
    free(evOff);
    free(evBytes);
    free(evId);
    free(cks);
    free(pay);
    free(endb);
    free(pre);
    free(fl);
    free(ol);
    if (rc != KNZ_OK)
        return rc;
    *outLen = produced;
    for (int i = 0; i < 8; i++)
        ctx->ms[i] = acc[i];
    return KNZ_OK;
}

extern "C" int knz_decompress(knz_ctx* ctx, const uint8_t* in, int64_t n, uint8_t* out, int64_t cap, int64_t* outLen)
{
    return decompress_impl(ctx, in, n, out, cap, outLen, 1, 0x7FFFFFFF);
}

extern "C" int knz_decompress_range(knz_ctx* ctx, const uint8_t* in, int64_t n, int fromBlock, int toBlock, uint8_t* out,
                                    int64_t cap, int64_t* outLen)
{
    return decompress_impl(ctx, in, n, out, cap, outLen, fromBlock, toBlock);
}

extern "C" int knz_decompress_seek(knz_ctx* ctx, const uint8_t* in, int64_t n, int64_t bitPos, int nBlocks, uint8_t* out,
                                   int64_t cap, int64_t* outLen)
{
    if (bitPos < 0 || nBlocks < 1)
        return KNZ_ERR_INVALID_PARAM;
    return decompress_impl(ctx, in, n, out, cap, outLen, 1, (nBlocks >= 0x7FFFFFFE) ? 0x7FFFFFFF : nBlocks + 1, bitPos);
}

// ------------------------------------------------------------------ stage-level API
static int run_single_stage(knz_ctx* ctx, int type, bool inverse, const u8* in, int n, u8* out, int cap, int* outLen,
                            int* applied)
{
    if (!ctx || !in || !out || !outLen || !applied || n < 0)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    *applied = 0;
    *outLen = 0;
    if (n == 0) {
        *applied = 1;
        return KNZ_OK;
    }
    if (knz_is_host_stage(type)) { // host stage (pre.cu): nothing for the device to do
        if (cap < 0)
            return KNZ_ERR_BLOCK_SIZE;
        // a stage on its own sees an empty Context: no data type, no "blockSize"; the text codec variant is the
        // one of an ANS0 stream (what the reference-side shim of the tests sets)
        KnzPreCtx pc = { KDT_UNDEFINED, 0, E_ANS0 };
        int len = 0;
        const bool ok = inverse ? knz_pre_inverse(type, in, n, out, cap, &len, &pc) : knz_pre_forward(type, in, n, out, cap, &len, &pc);
        if (ok) {
            *outLen = len;
            *applied = 1;
        }
        return KNZ_OK;
    }
    if (!type_supported(type))
        return KNZ_ERR_INVALID_CODEC;
    if (n > ctx->stageCap || cap < 0)
        return KNZ_ERR_BLOCK_SIZE;
    const int devCap = (cap < ctx->stageCap) ? cap : ctx->stageCap; // what a stage buffer slot can hold
    {
        const int rcb = ensure_bwt(ctx, &type, 1);
        if (rcb != KNZ_OK)
            return rcb;
    }
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->dStageIn, in, (size_t)n, cudaMemcpyHostToDevice, s));
    ctx->h_st[0].len = n;
    ctx->h_st[0].cur = 2;
    ctx->h_st[0].swaps = 0;
    ctx->h_st[0].flags = inverse ? 0x00 : 0xFF;
    ctx->h_capEven[0] = devCap;
    ctx->h_capOdd[0] = devCap;
    CK(cudaMemcpyAsync(ctx->st, ctx->h_st, sizeof(BlkState), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->capEven, ctx->h_capEven, sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->capOdd, ctx->h_capOdd, sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->errFlag, 0, sizeof(int) * 4, s));
    StageLaunch L;
    L.bt.base[0] = ctx->bufA;
    L.bt.base[1] = ctx->bufB;
    L.bt.base[2] = ctx->dStageIn;
    L.bt.stride[0] = L.bt.stride[1] = L.bt.stride[2] = ctx->bstride;
    L.stIn = ctx->st;
    L.stOut = ctx->st + ctx->maxBatch;
    L.stageIdx = 0;
    L.nBlocks = 1;
    L.maxLen = (n > devCap ? n : devCap) + 64;
    L.capEven = ctx->capEven;
    L.capOdd = ctx->capOdd;
    L.errFlag = ctx->errFlag;
    L.wsBlock0 = 0;
    CK(cudaEventRecord(ctx->ev[1], s));
    if (inverse)
        launch_inverse_stage(ctx, type, L, s);
    else
        launch_forward_stage(ctx, type, L, s);
    CK(cudaEventRecord(ctx->ev[2], s));
    CK(cudaMemcpyAsync(ctx->h_err, ctx->errFlag, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_st, L.stOut, sizeof(BlkState), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
        for (int k = 0; k < 8; k++)
            ctx->ms[k] = 0.f;
        add_stage_time(ctx, type, ms);
        ctx->ms[5] = ms;
    }
    if (ctx->h_err[0] != 0) {
        if (inverse && ctx->h_err[0] == KERR_BAD_STREAM)
            return KNZ_OK; // inverse() returns false on malformed input: applied stays 0
        return map_kerr(ctx, ctx->h_err[0] ? ctx->h_err[0] : ctx->h_err[1] ? ctx->h_err[1] : ctx->h_err[2]);
    }
    const BlkState r = ctx->h_st[0];
    if (r.swaps == 0)
        return KNZ_OK; // stage refused
    if (r.len > cap)
        return KNZ_OK;
    const u8* src = (r.cur == 0) ? ctx->bufA : (r.cur == 1) ? ctx->bufB : ctx->dStageIn;
    CK(cudaMemcpyAsync(out, src, (size_t)r.len, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *outLen = r.len;
    *applied = 1;
    return KNZ_OK;
}

extern "C" int knz_transform_forward(knz_ctx* ctx, int type, const uint8_t* in, int n, uint8_t* out, int cap,
                                     int* outLen, int* applied)
{
    return run_single_stage(ctx, type, false, in, n, out, cap, outLen, applied);
}

extern "C" int knz_transform_inverse(knz_ctx* ctx, int type, const uint8_t* in, int n, uint8_t* out, int cap,
                                     int* outLen, int* applied)
{
    return run_single_stage(ctx, type, true, in, n, out, cap, outLen, applied);
}

extern "C" int knz_entropy_encode(knz_ctx* ctx, int type, const uint8_t* in, int n, uint8_t* out, int64_t cap,
                                  int64_t* outBits)
{
    if (!ctx || !in || !out || !outBits || n <= 0)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    if (!entropy_supported(type))
        return KNZ_ERR_INVALID_CODEC;
    if ((i64)n + 64 > ctx->bstride)
        return KNZ_ERR_BLOCK_SIZE;
    cudaSetDevice(ctx->device);
    {
        const int rc1 = ensure_ans1(ctx, type);
        if (rc1 != KNZ_OK)
            return rc1;
    }
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->dStageIn, in, (size_t)n, cudaMemcpyHostToDevice, s));
    ctx->h_st[0].len = n;
    ctx->h_st[0].cur = 2;
    ctx->h_st[0].swaps = 0;
    ctx->h_st[0].flags = 0xFF;
    CK(cudaMemcpyAsync(ctx->st, ctx->h_st, sizeof(BlkState), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->errFlag, 0, sizeof(int) * 4, s));
    EncodeLaunch E;
    E.bt.base[0] = ctx->bufA;
    E.bt.base[1] = ctx->bufB;
    E.bt.base[2] = ctx->dStageIn;
    E.bt.stride[0] = E.bt.stride[1] = E.bt.stride[2] = ctx->bstride;
    E.st = ctx->st;
    E.nBlocks = 1;
    E.maxChunks = ctx->maxChunks;
    E.eType = type;
    E.nTransforms = 1;
    E.slots = ctx->slots;
    E.hdrBits = ctx->hdrBits;
    E.payBytes = ctx->payBytes;
    E.payOff = ctx->payOff;
    E.chunkOff = ctx->chunkOff;
    E.blockBits = ctx->blockBits;
    E.out = ctx->dOut;
    E.outStride = ctx->outStride;
    E.errFlag = ctx->errFlag;
    E.evK0 = E.evK1 = NULL;
    E.a1 = &ctx->a1;
    launch_entropy_encode(E, s, &ctx->launches);
    CK(cudaMemcpyAsync(ctx->h_err, ctx->errFlag, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_bits, ctx->blockBits, sizeof(u64), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (ctx->h_err[0])
        return map_kerr(ctx, ctx->h_err[0] ? ctx->h_err[0] : ctx->h_err[1] ? ctx->h_err[1] : ctx->h_err[2]);
    // strip the block header the block-level path put in front (mode + length bytes)
    const int dataSize = (n < 256) ? 1 : (ilog2_u32((u32)n) >> 3) + 1;
    const int hdr = 1 + dataSize;
    const u64 bits = ctx->h_bits[0] - 8ull * (u64)hdr;
    const i64 nbytes = (i64)((bits + 7) >> 3);
    if (nbytes > cap)
        return KNZ_ERR_OUTPUT_TOO_SMALL;
    CK(cudaMemcpyAsync(out, ctx->dOut + hdr, (size_t)nbytes, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *outBits = (i64)bits;
    return KNZ_OK;
}

extern "C" int knz_entropy_decode(knz_ctx* ctx, int type, const uint8_t* in, int64_t inBits, uint8_t* out, int n)
{
    if (!ctx || !in || !out || n <= 0 || inBits < 0)
        return KNZ_ERR_INVALID_PARAM;
    std::lock_guard<std::recursive_mutex> lock_(ctx->mtx);
    if (!entropy_supported(type))
        return KNZ_ERR_INVALID_CODEC;
    const i64 nbytes = (inBits + 7) >> 3;
    if ((i64)n + 64 > ctx->bstride || nbytes + 16 > ctx->outStride)
        return KNZ_ERR_BLOCK_SIZE;
    cudaSetDevice(ctx->device);
    {
        const int rc1 = ensure_ans1(ctx, type);
        if (rc1 != KNZ_OK)
            return rc1;
    }
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->dOut + (nbytes & ~(i64)15), 0, 32, s));
    CK(cudaMemcpyAsync(ctx->dOut, in, (size_t)nbytes, cudaMemcpyHostToDevice, s));
    ctx->h_preLen[0] = n;
    ctx->h_payStart[0] = 0;
    ctx->h_bits[0] = (u64)inBits;
    CK(cudaMemcpyAsync(ctx->dPreLen, ctx->h_preLen, sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dPayStart, ctx->h_payStart, sizeof(u64), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dInBits, ctx->h_bits, sizeof(u64), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->errFlag, 0, sizeof(int) * 4, s));
    DecodeLaunch D;
    D.in = ctx->dOut;
    D.inStride = ctx->outStride;
    D.inBits = ctx->dInBits;
    D.payStart = ctx->dPayStart;
    D.preLen = ctx->dPreLen;
    D.nBlocks = 1;
    D.maxChunks = ctx->maxChunks;
    D.eType = type;
    D.chunkPos = ctx->chunkPos;
    D.dst = ctx->bufA;
    D.dstStride = ctx->bstride;
    D.errFlag = ctx->errFlag;
    D.evK0 = D.evK1 = NULL;
    D.a1 = &ctx->a1;
    launch_entropy_decode(D, s, &ctx->launches);
    CK(cudaMemcpyAsync(ctx->h_err, ctx->errFlag, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(out, ctx->bufA, (size_t)n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return map_kerr(ctx, ctx->h_err[0] ? ctx->h_err[0] : ctx->h_err[1] ? ctx->h_err[1] : ctx->h_err[2]);
}
