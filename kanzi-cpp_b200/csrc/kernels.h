// kernels.h -- internal launch interfaces between api.cu and the kernel files.
#pragma once
#include "common.cuh"

#define ANS_CHUNK 16384
#define ANS0_LR 12
// Per-chunk staging slot: [0,512) header bit string (+varint, 4 states);
// renormalisation words grow downwards from ANS_WEND; <= 3 tail bytes follow it.
#define ANS_HDR_AREA 512
#define ANS_WEND (ANS_HDR_AREA + 2 * ANS_CHUNK)
#define ANS_SLOT (ANS_WEND + 512)

#define E_RAW 0
#define E_HUF 1
#define E_FPAQ 2
#define E_ANS0 5
#define E_ANS1 8

// order-1 rANS: chunk = 16384 << 8 bytes, logRange 11 (entropy/ANSRangeEncoder.cpp:59-67)
#define ANS1_CHUNK (16384 << 8)
#define ANS1_LR 11

#define T_NONE 0
#define T_BWT 1
#define T_ZRLT 6
#define T_MTFT 7
#define T_RANK 8
#define T_SRT 13
#define T_LZ 3
#define T_LZP 14
#define T_LZX 16

// device-side error flags (first one wins)
#define KERR_OUT_OVERFLOW 1
#define KERR_BAD_STREAM 2
#define KERR_UNSUPPORTED 3
#define KERR_INTERNAL 4
#define KERR_CRC 5 // block checksum mismatch (reported in errFlag[2])

// Scratch of the order-1 rANS coder (ans1.cu), allocated on first use of ANS1 by a context.
struct Ans1Work {
    int maxBlocks, stageCap, cpb; // cpb = order-1 chunks per block
    i64 payRegion, payStride, recStride;
    u32* freq;     // [blocks*cpb][256][256] pair counts (decode: per-context symbol lists)
    u64* tenc;     // [blocks*cpb][256][256] packed encoder entries
    u64* rec;      // [blocks][recStride]    every position's encoder entry (pre-mapped records)
    u8* hdr;       // [blocks*cpb][256][512] context header bit strings
    u32* hbits;    // [blocks*cpb][256]      their lengths (decode: alphabet sizes)
    u8* pay;       // [blocks][cpb][payRegion] renormalisation words (backwards) + tail bytes
    void* trailer; // [blocks*cpb] payload size + final states
    u64* pieceOff; // [blocks*cpb][258] bit offsets of the pieces of a chunk
    u32* tdec;     // [blocks*cpb][256][2048] decoder slot tables
    void* dmeta;   // [blocks*cpb] decoder: payload position, size, initial states
};
bool ans1_work_alloc(Ans1Work& W, int maxBlocks, int stageCap);
void ans1_work_free(Ans1Work& W);

struct EncodeLaunch {
    BufTable bt;
    const BlkState* st; // state after the last transform stage
    int nBlocks, maxChunks, eType;
    int nTransforms;      // number of transforms of the sequence | (block checksum bytes << 8), see knz_hdr_bytes
    const u64* blockHash; // per block XXHash (when checksum bytes != 0)
    u8* slots;
    u32 *hdrBits, *payBytes, *payOff;
    u64 *chunkOff, *blockBits;
    u8* out;
    i64 outStride;
    int* errFlag;
    cudaEvent_t evK0, evK1; // optional: bracket the rANS kernel alone (NULL = off)
    Ans1Work* a1;           // ANS1 only
};
void launch_entropy_encode(const EncodeLaunch& L, cudaStream_t s, u64* launches);
void launch_ans1_encode(const EncodeLaunch& L, cudaStream_t s, u64* launches);
void launch_fpaq_encode(const EncodeLaunch& L, cudaStream_t s, u64* launches);
// zero the output words of every block and write the block headers (shared by all entropy coders)
void launch_out_prepare_and_header(const EncodeLaunch& L, cudaStream_t s, u64* launches);
// startBit / endBit are device scalars (may alias): batches chain without a host round trip.
void launch_stream_assemble(const u8* blockOut, i64 outStride, const u64* blockBits, int nBlocks,
                            const u64* startBit, u64* blockOff, u64* endBit, u8* stream, cudaStream_t s,
                            u64* launches, const int* srcIndex = NULL);

// Entropy decode: block b's bit string is at in + b*inStride; its entropy payload
// starts at bit payStart[b] and must yield preLen[b] bytes into dst (buffer A).
struct DecodeLaunch {
    const u8* in;
    i64 inStride;
    const u64* inBits;    // per block total bits
    const u64* payStart;  // per block: bit offset of the entropy payload
    const int* preLen;    // per block: bytes to decode
    int nBlocks, maxChunks, eType;
    u64* chunkPos;        // [nBlocks][maxChunks] bit offset of each chunk header
    u8* dst;
    i64 dstStride;
    int* errFlag;
    cudaEvent_t evK0, evK1; // optional: bracket the rANS decode kernel alone
    Ans1Work* a1;           // ANS1 only
};
void launch_entropy_decode(const DecodeLaunch& L, cudaStream_t s, u64* launches);
void launch_ans1_decode(const DecodeLaunch& L, cudaStream_t s, u64* launches);
void launch_fpaq_decode(const DecodeLaunch& L, cudaStream_t s, u64* launches);
void launch_huffman_encode_chunks(const EncodeLaunch& L, cudaStream_t s, u64* launches);
void launch_huffman_decode(const DecodeLaunch& L, cudaStream_t s, u64* launches);

// Transform stages.  Every launcher reads st[s] and writes st[s+1].
struct StageLaunch {
    BufTable bt;
    const BlkState* stIn;
    BlkState* stOut;
    int stageIdx;   // position in the sequence (skip-flag bit 7 - stageIdx)
    int nBlocks;
    int maxLen;     // upper bound of len over the batch (grid sizing)
    const int* capEven; // per block: capacity of the reference's `output` buffer
    const int* capOdd;  // per block: capacity of the reference's `input`/private buffer
    int* errFlag;
    int wsBlock0;       // first workspace slot of this launch (block groups decoded concurrently on
                        // separate streams use disjoint slices of the per-block scratch arrays)
    const int* dtype = nullptr; // per block Global::DataType left by the host stages (NULL: undefined); read by LZ / LZX
};
struct Workspace;
void launch_none_forward(const StageLaunch& L, cudaStream_t s, u64* launches);
void launch_zrlt_forward(const StageLaunch& L, Workspace& ws, cudaStream_t s, u64* launches);
void launch_zrlt_inverse(const StageLaunch& L, Workspace& ws, cudaStream_t s, u64* launches);
void launch_sbrt_forward(const StageLaunch& L, int mode, Workspace& ws, cudaStream_t s, u64* launches);
void launch_sbrt_inverse(const StageLaunch& L, int mode, Workspace& ws, cudaStream_t s, u64* launches);
// MTFT / RANK ranks of nBlocks buffers without the stage bookkeeping (stIn[b].len bytes from
// bt.base[0] to bt.base[1]; blocks with stOut[b].swaps == stIn[b].swaps are left alone): SRT's ranks
// are the MTFT ranks of the block relabelled by first appearance.
void launch_sbrt_rank_only(const BufTable& bt, const BlkState* stIn, const BlkState* stOut, int nBlocks, int maxLen,
                           int mode, Workspace& ws, cudaStream_t s, u64* launches);
// Sorted Rank Transform scratch (srt.cu), allocated on the first SRT stage of a context.
struct SrtWork {
    u32* cntT;       // [blocks][tiles][256] per-tile symbol counts -> exclusive prefix over tiles
    u32* firstT;     // [blocks][tiles][256] first position of each symbol in the tile
    u8* relabel;     // [blocks][256] symbol -> first-appearance index
    u32* bucketBase; // [blocks][256] header size + bucket start of each symbol
    BlkState *fakeIn, *fakeOut; // states handed to the rank kernels
    u8 *tmp1, *tmp2; // relabelled block, its MTFT ranks
    i64 tmpStride;
};
bool srt_work_alloc(SrtWork& W, int maxBlocks, int capN);
void srt_work_free(SrtWork& W);
void launch_srt_forward(const StageLaunch& L, Workspace& ws, SrtWork& W, cudaStream_t s, u64* launches);
void launch_srt_inverse(const StageLaunch& L, cudaStream_t s, u64* launches);
// LZ / LZX / LZP scratch (lz.cu), allocated on the first LZ-family stage of a context.
struct LzWork {
    int* hashes;    // [blocks][2^19] last position of every hash (LZ and LZP use the first 2^16)
    i64 hashStride; // ints per block
    u8* side;       // [blocks][2][sideStride] tokens + match lengths | distances
    i64 sideStride;
};
bool lz_work_alloc(LzWork& W, int maxBlocks, i64 stageStride);
void lz_work_free(LzWork& W);
void launch_lz_forward(const StageLaunch& L, int type, LzWork& W, cudaStream_t s, u64* launches);
void launch_lz_inverse(const StageLaunch& L, int type, LzWork& W, cudaStream_t s, u64* launches);
void launch_bwt_forward(const StageLaunch& L, Workspace& ws, cudaStream_t s, u64* launches);
void launch_bwt_inverse(const StageLaunch& L, Workspace& ws, cudaStream_t s, u64* launches);

// Scratch memory shared by the stage launchers (allocated once per context).
struct Workspace {
    int maxBlocks;   // batch capacity
    int capN;        // per-block element capacity (maxBlockSize + slack)
    // generic per-tile scratch (ZRLT / SBRT / scans)
    u32* tileA;      // [maxBlocks * maxTiles * 4]
    u32* tileB;
    i64 tileWords;
    // SBRT per-tile last-two-occurrence tables: [maxBlocks][tiles][256][2]
    u32* occ;
    i64 occWords;
    u32* zlen;       // [maxBlocks] ZRLT: output length per block (or "refused")
    // suffix sorting
    u64 *keyA, *keyB;   // [maxBlocks * capN]
    u32 *valA, *valB;
    u32 *grpA, *grpB;
    u32* isa;           // [maxBlocks * capN]
    u32* hist;          // [maxBlocks * sortTiles * 256]
    u32* digitBase;     // [maxBlocks * 256]
    u32* totals;        // [maxBlocks * 8 * 256]
    int* trivial;       // [8][maxBlocks]
    int* which;         // [16][maxBlocks]
    int* cnt;           // [maxBlocks] elements in play per block
    int* cntNext;
    int* h_cnt;         // pinned host mirror
    u32* scanA;         // [maxBlocks * sortTiles * 4]
    u32* scanB;         // [maxBlocks * sortTiles * 2]
    u64 *xkeyA, *xkeyB; // side sort of tile-crossing groups
    u32 *xvalA, *xvalB;
    int* whichX;        // [9][maxBlocks]
    int* cntX;          // [maxBlocks]
    int* pidx;          // [maxBlocks * 8] primary indexes
    int* bwtOk;         // [maxBlocks]
    bool bwtReady;      // suffix-sort arrays allocated (workspace_alloc_bwt)
};

bool workspace_alloc(Workspace& ws, int maxBlocks, int capN);
bool workspace_alloc_bwt(Workspace& ws); // on the first BWT stage of a context
void workspace_free(Workspace& ws);
// Copies each block's final bytes to out + b*outStride; a block longer than outCap is not copied
// and raises KERR_OUT_OVERFLOW (a crafted stream may not write past its destination slot).
// Block checksums (xxhash.cu): hash of st[b].len bytes of every block; expect == NULL stores it,
// else compares and raises KERR_CRC in errFlag[2].  bits = 32 or 64.
void launch_xxhash(const BufTable& bt, const BlkState* st, int nBlocks, int bits, u64* hash, const u64* expect,
                   int* errFlag, cudaStream_t s, u64* launches);
u64 knz_xxhash_host(const u8* data, int length, int bits);
// skipBlocks (blockscan.cu): entropy / signature test of every block, and the copy-block framing that
// overrides the private buffer of the flagged ones.  log2tab: 257 ints, round(4096 * log2(i)).
void launch_skip_decide(const BufTable& bt, const BlkState* st, int nBlocks, const int* log2tab, int* skip,
                        cudaStream_t s, u64* launches);
void launch_copy_frame(const BufTable& bt, const BlkState* st0, int nBlocks, const int* skip, int ckBytes,
                       const u64* blockHash, u8* out, i64 outStride, u64* blockBits, cudaStream_t s, u64* launches);
void launch_copy_out(const BufTable& bt, const BlkState* st, int nBlocks, u8* out, i64 outStride, int outCap,
                     int* errFlag, cudaStream_t s, u64* launches);
