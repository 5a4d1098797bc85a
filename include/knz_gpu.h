/* knz_gpu.h -- C ABI of libknzgpu.so: the B200-native replacement for kanzi's
 * per-block transform -> entropy path (EncodingTask::run / DecodingTask::run and
 * the Transform<byte> / EntropyEncoder / EntropyDecoder stages they instantiate).
 *
 * Plain pointers and sizes only; no exceptions cross this boundary; the caller
 * owns every host buffer, the library owns device memory, streams and kernels.
 * All functions return 0 on success or a kanzi error code (src/Error.hpp:27-49;
 * the ones used here are repeated below).  There is NO CPU fallback: if no CUDA
 * device is usable knz_create() fails with KNZ_ERR_CREATE_COMPRESSOR.
 *
 * Each entry point names the reference interface it replaces (file:line under
 * /root/reference/src).  INTEGRATION.md shows the reference-side bindings.
 */
#ifndef KNZ_GPU_H
#define KNZ_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error.hpp:27-49 */
#define KNZ_OK 0
#define KNZ_ERR_MISSING_PARAM 1
#define KNZ_ERR_BLOCK_SIZE 2
#define KNZ_ERR_INVALID_CODEC 3
#define KNZ_ERR_CREATE_COMPRESSOR 4
#define KNZ_ERR_CREATE_DECOMPRESSOR 5
#define KNZ_ERR_OUTPUT_TOO_SMALL 12 /* ERR_WRITE_FILE: the caller's output buffer is too small */
#define KNZ_ERR_PROCESS_BLOCK 13
#define KNZ_ERR_INVALID_FILE 15
#define KNZ_ERR_STREAM_VERSION 16
#define KNZ_ERR_INVALID_PARAM 18
#define KNZ_ERR_CRC_CHECK 19
#define KNZ_ERR_UNKNOWN 127

/* Transform ids are wire-visible: transform/TransformFactory.hpp:49-73 */
#define KNZ_T_NONE 0
#define KNZ_T_BWT 1
#define KNZ_T_ZRLT 6
#define KNZ_T_MTFT 7
#define KNZ_T_RANK 8
#define KNZ_T_SRT 13
#define KNZ_T_LZ 3
#define KNZ_T_LZP 14
#define KNZ_T_LZX 16
/* Entropy ids: entropy/EntropyEncoderFactory.hpp:37-52 */
#define KNZ_E_NONE 0
#define KNZ_E_HUFFMAN 1
#define KNZ_E_FPAQ 2
#define KNZ_E_ANS0 5
#define KNZ_E_ANS1 8

typedef struct knz_ctx knz_ctx;

/* Library lifetime.  `device` is the CUDA ordinal; maxBlockSize is the stream
 * block size the context is provisioned for (multiple of 16, <= 64 MiB here);
 * maxBatchBlocks bounds how many blocks are processed per device batch (working
 * memory is ~48 B per input byte of a batch for the suffix sorter).             */
int knz_create(int device, int maxBlockSize, int maxBatchBlocks, knz_ctx** out);
void knz_destroy(knz_ctx* ctx);
const char* knz_last_error(const knz_ctx* ctx);

/* "BWT+RANK+ZRLT" -> 48-bit type word, first stage in the top 6 bits
 * (TransformFactory<T>::getType, transform/TransformFactory.hpp:100-137);
 * returns (uint64_t)-1 for a name this library does not implement.              */
uint64_t knz_transform_type(const char* name);
/* EntropyEncoderFactory::getType (entropy/EntropyEncoderFactory.hpp:131);  -1 if unsupported */
int knz_entropy_type(const char* name);

/* ---- Block level: what EncodingTask<T>::run (io/CompressedOutputStream.cpp:652-898)
 * builds in its private buffer for each block -- mode byte, [skip-flag byte],
 * post-transform length, entropy payload -- for nBlocks blocks at once.
 *   in          nBlocks blocks laid out at `inStride` bytes from each other
 *   lens[i]     bytes in block i (1 .. maxBlockSize)
 *   out         block i's bytes are written at out + i*outStride
 *   outBits[i]  exact bit count (`written` in the reference, :830)
 *   skipFlags[i] TransformSequence skip flags (:745)
 * The ping-pong buffer capacities that make ZRLT's accept/refuse decision
 * (TransformSequence.hpp:88-162) follow the reference's jobs=1 buffer model with
 * `firstBlockLen` = length of the first block of the stream and `blockSize` the
 * stream block size (io/CompressedOutputStream.cpp:138-146, :733-739).            */
int knz_encode_blocks(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* in,
                      int64_t inStride, const int32_t* lens, int nBlocks, int firstBlockLen, uint8_t* out,
                      int64_t outStride, uint64_t* outBits, uint8_t* skipFlags);

/* DecodingTask<T>::run (io/CompressedInputStream.cpp:791-1041), payload part:
 * block i = inBits[i] bits at in + i*inStride (byte aligned, as the reference's
 * per-task copy :843-856); decoded bytes go to out + i*outStride; outLens[i] =
 * decoded length.  blockSize = stream block size (sizes the task buffers).       */
int knz_decode_blocks(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* in,
                      int64_t inStride, const uint64_t* inBits, int nBlocks, uint8_t* out, int64_t outStride,
                      int32_t* outLens);

/* ---- Stream level: CompressedOutputStream::write + close
 * (io/CompressedOutputStream.cpp:361-440, header :277-342, per-block length
 * prefixes :852-864, end marker :416-417) and CompressedInputStream::read
 * (io/CompressedInputStream.cpp:428-510, readHeader :511-663), host buffers.
 * Output is byte-identical to the reference's stream for the same parameters
 * (block checksums: knz_set_checksum; skipBlocks: knz_set_skip_blocks).           */
int knz_compress(knz_ctx* ctx, const char* transform, const char* entropy, int blockSize, const uint8_t* in,
                 int64_t n, uint8_t* out, int64_t cap, int64_t* outLen);
int knz_decompress(knz_ctx* ctx, const uint8_t* in, int64_t n, uint8_t* out, int64_t cap, int64_t* outLen);

/* Same work with the data already resident in device memory (bench `value` leg;
 * also what a multi-GPU rank calls on its shard).  d_in / d_out are device
 * pointers in the context's device.  compress: blocks [firstBlock, firstBlock+nBlocks)
 * of a stream of total length streamLen whose block 0 has length min(blockSize,
 * streamLen); d_blockOut receives each block's private buffer (outStride apart),
 * d_outBits (device, uint64) the bit counts.                                      */
int knz_encode_blocks_dev(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* d_in,
                          int64_t inStride, const int32_t* lens, int nBlocks, int firstBlockLen,
                          uint8_t* d_blockOut, int64_t outStride, uint64_t* d_outBits,
                          uint8_t* h_skipFlags /* may be NULL */);
int knz_decode_blocks_dev(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* d_in,
                          int64_t inStride, const uint64_t* h_inBits, int nBlocks, uint8_t* d_out,
                          int64_t outStride, int32_t* h_outLens);
/* Bit-concatenate nBlocks block buffers (device) into the final stream body on
 * the device: for each block `lw-3`(5) | bits(lw) | payload, starting at bit
 * `startBit`; returns the end bit position.  (CompressedOutputStream.cpp:852-864) */
int knz_assemble_stream_dev(knz_ctx* ctx, const uint8_t* d_blockOut, int64_t outStride, const uint64_t* d_outBits,
                            int nBlocks, uint8_t* d_stream, int64_t streamCap, uint64_t startBit,
                            uint64_t* endBit);
/* Stream header bytes (io/CompressedOutputStream.cpp:277-342); returns byte count (20..26). */
int knz_stream_header(uint64_t tType, int eType, int blockSize, int64_t inputSize, uint8_t out[32]);
int knz_stream_header_ex(uint64_t tType, int eType, int blockSize, int64_t inputSize, int checksumBits, uint8_t out[32]);
/* Block checksums (the `checksum` parameter of CompressedOutputStream, io/CompressedOutputStream.cpp:99-117):
 * bits = 0, 32 or 64.  Encoders of this context then hash every block with kanzi's XXHash32 / XXHash64
 * (util/XXHash.hpp, seed "KANZ") before the transforms and write the value behind the block length (:674-682,
 * :804-807); the block-level decoders expect it there.  knz_decompress / knz_decompress_dist read the
 * checksum size from the stream header and verify every decoded block (io/CompressedInputStream.cpp:
 * 1003-1022), failing with KNZ_ERR_CRC_CHECK.                                                            */
int knz_set_checksum(knz_ctx* ctx, int bits);

/* `skipBlocks` of the reference's context (io/CompressedOutputStream.cpp:697-715): when on, a block whose
 * first four bytes are the signature of an already compressed format (Magic.hpp:69-131) or whose order-0
 * entropy reaches EntropyUtils::INCOMPRESSIBLE_THRESHOLD (973/1024 bits per byte, Global.cpp:313-329) is
 * written as a copy block (mode 0x80, no transform, no entropy coding).  Applies to every encoder of the
 * context; the decoders need nothing (copy blocks are part of the format).                               */
int knz_set_skip_blocks(knz_ctx* ctx, int on);

/* Block range of CompressedInputStream (context entries "from" / "to", io/CompressedInputStream.cpp:836-837,
 * :864-869): block ids are 1-based; blocks with fromBlock <= id < toBlock are decoded and concatenated in
 * `out`, the others are stepped over through their length prefixes (only the bytes of the stream that
 * hold the wanted blocks are sent to the device).  knz_decompress == range [1, INT_MAX).                */
int knz_decompress_range(knz_ctx* ctx, const uint8_t* in, int64_t n, int fromBlock, int toBlock, uint8_t* out,
                         int64_t cap, int64_t* outLen);

/* CompressedInputStream::seek (io/CompressedInputStream.hpp:329-375): restart decoding at a block boundary given
 * as a BIT position of the stream -- the `offset` of a BLOCK_INFO event -- and decode up to nBlocks blocks from
 * there.  The stream header is still read (it carries the pipeline and the block size).                  */
int knz_decompress_seek(knz_ctx* ctx, const uint8_t* in, int64_t n, int64_t bitPos, int nBlocks, uint8_t* out,
                        int64_t cap, int64_t* outLen);

/* Listener events (Event.hpp:28-88).  The reference's tasks notify their listeners around every stage of
 * every block (io/CompressedOutputStream.cpp:685-689, :768-772, :810-814, :871-881;
 * io/CompressedInputStream.cpp:380, :924-932, :973-983); here a batch of blocks goes through the device
 * as one unit, so knz_compress / knz_decompress re-emit the same events, with the same ids, sizes, hashes,
 * bit offsets and skip flags, when the batch has completed -- per block in the reference's order, blocks
 * in stream order, on the calling thread.  type values are those of Event::Type.                         */
#define KNZ_EVT_BEFORE_TRANSFORM 2
#define KNZ_EVT_AFTER_TRANSFORM 3
#define KNZ_EVT_BEFORE_ENTROPY 4
#define KNZ_EVT_AFTER_ENTROPY 5
#define KNZ_EVT_BLOCK_INFO 9
typedef struct knz_event {
    int type;          /* KNZ_EVT_* */
    int blockId;       /* 1-based */
    int64_t size;      /* bytes: block / post-transform / coded size as in the reference's event */
    uint64_t hash;     /* block checksum when hashBits != 0 */
    int hashBits;      /* 0, 32 or 64 (Event::HashType) */
    int64_t offset;    /* BLOCK_INFO: bit position of the block in the stream, else -1 */
    uint8_t skipFlags; /* BLOCK_INFO: TransformSequence skip flags */
} knz_event;
typedef void (*knz_event_fn)(void* user, const knz_event* evt);
int knz_set_listener(knz_ctx* ctx, knz_event_fn fn, void* user); /* fn == NULL: off */

/* ---- Stage level (what the Transform<byte> / EntropyEncoder adapters call).
 * knz_transform_forward == Transform<byte>::forward (src/Transform.hpp:38):
 * returns KNZ_OK and *applied=1 when the stage produced output (*outLen bytes),
 * *applied=0 when the stage refuses (reference returns false, stage skipped).
 * type is a single transform id; cap = destination capacity (_length - _index).  */
int knz_transform_forward(knz_ctx* ctx, int type, const uint8_t* in, int n, uint8_t* out, int cap, int* outLen,
                          int* applied);
int knz_transform_inverse(knz_ctx* ctx, int type, const uint8_t* in, int n, uint8_t* out, int cap, int* outLen,
                          int* applied);
/* EntropyEncoder::encode (src/EntropyEncoder.hpp:30) into a private bit buffer:
 * out receives ceil(*outBits/8) bytes, MSB-first.                                */
int knz_entropy_encode(knz_ctx* ctx, int type, const uint8_t* in, int n, uint8_t* out, int64_t cap,
                       int64_t* outBits);
/* EntropyDecoder::decode (src/EntropyDecoder.hpp:30): n = number of bytes to produce. */
int knz_entropy_decode(knz_ctx* ctx, int type, const uint8_t* in, int64_t inBits, uint8_t* out, int n);

/* ---- Multi-GPU: one process per GPU, blocks sharded round-robin (block i -> rank i % world).
 * Replaces the task pool of CompressedOutputStream / CompressedInputStream across devices: the ordered
 * append to the shared bitstream (io/CompressedOutputStream.cpp:836-868) becomes an all-gather of the
 * per-block bit counts, a gather of the block payloads to rank 0 over NCCL and one bit-concatenation
 * kernel there; on decode the host walks the length prefixes (io/CompressedInputStream.cpp:823-856)
 * and ships only its own blocks' bit ranges to its GPU.
 *   knz_dist_unique_id   rank 0 creates the NCCL id, the launcher hands it to every rank
 *   knz_dist_init        collective: joins the context to an NCCL communicator of `world` ranks
 *   knz_dist_init_transport   same with caller-supplied collectives over device buffers (tests; gloo)  */
typedef int (*knz_allgather_fn)(void* user, const void* d_send, int64_t bytes, void* d_recv);  /* recv: world * bytes, rank order */
typedef int (*knz_gather_fn)(void* user, const void* d_send, int64_t bytes, void* d_recv);     /* recv used on rank 0 only */
typedef int (*knz_bcast_fn)(void* user, void* d_buf, int64_t bytes);                            /* from rank 0 */
int knz_dist_unique_id(uint8_t id[128]);
int knz_dist_init(knz_ctx* ctx, int rank, int world, const uint8_t id[128]);
int knz_dist_init_transport(knz_ctx* ctx, int rank, int world, knz_allgather_fn allgather, knz_gather_fn gather,
                            knz_bcast_fn bcast, void* user);
/* Stream level, host buffers, collective.  compress: every rank passes the whole input, encodes the
 * blocks it owns, rank 0 receives the stream (*outLen = 0 elsewhere); byte-identical to knz_compress.
 * decompress: every rank passes the whole stream and a buffer for the whole original and fills the
 * blocks it owns (block i at out + i * blockSize; pass a shared mapping to collect them in one place).  */
int knz_compress_dist(knz_ctx* ctx, const char* transform, const char* entropy, int blockSize, const uint8_t* in,
                      int64_t n, uint8_t* out, int64_t cap, int64_t* outLen);
int knz_decompress_dist(knz_ctx* ctx, const uint8_t* in, int64_t n, uint8_t* out, int64_t cap, int64_t* outLen);
/* Same with the data resident in device memory.  encode: this rank's nbOwn blocks (block k of the rank is
 * block rank + k * world of the stream) -> stream body assembled in d_stream on rank 0 from bit `startBit`;
 * h_allBits (every rank, may be NULL) receives the bit count of all nBlocks blocks in stream order.
 * decode: d_stream holds the stream on rank 0 and is a receive buffer elsewhere (broadcast inside); the
 * rank's blocks are decoded to d_out + k * outStride.                                                    */
int knz_dist_encode_dev(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, const uint8_t* d_in, int64_t inStride,
                        const int32_t* lens, int nbOwn, int nBlocks, int firstBlockLen, uint8_t* d_stream,
                        int64_t streamCap, uint64_t startBit, uint64_t* h_allBits, uint64_t* endBit);
int knz_dist_decode_dev(knz_ctx* ctx, uint64_t tType, int eType, int blockSize, uint8_t* d_stream, int64_t streamBytes,
                        uint64_t startBit, const uint64_t* h_allBits, int nBlocks, uint8_t* d_out, int64_t outStride,
                        int32_t* h_outLens);

/* Instrumentation for bench.py: kernels launched by this context since creation,
 * and the CUDA stream (cudaStream_t) the library launches on.                    */
uint64_t knz_launch_count(const knz_ctx* ctx);
void* knz_stream(const knz_ctx* ctx);
/* Device time (ms) of the last call, split by stage group, measured with CUDA
 * events on the library stream: [0]=BWT [1]=RANK/MTFT [2]=ZRLT [3]=entropy stage
 * [4]=bit assembly [5]=total [6]=rANS encode kernel alone [7]=rANS decode kernel alone. */
void knz_last_timings(const knz_ctx* ctx, float ms[8]);
/* Decode scheduling: the blocks of a batch are decoded as `groups` (1..8) independent groups on
 * separate CUDA streams, so the serial inverse RANK/MTFT chain of one group (the counterpart of
 * SBRT::inverse, src/transform/SBRT.cpp:99-145, one dependency chain per block) overlaps the
 * entropy / ZRLT / BWT stages of the others -- the role DecodingTask concurrency plays in
 * src/io/CompressedInputStream.cpp:436-470.  1 = one group: knz_last_timings is split per stage;
 * with more groups only the total [5] is meaningful.  Default 1 (env KNZ_DEC_OVERLAP): on B200 the
 * chain warps slow down when they share SMs with the other groups' kernels.      */
int knz_set_decode_groups(knz_ctx* ctx, int groups);

#ifdef __cplusplus
}
#endif
#endif /* KNZ_GPU_H */
