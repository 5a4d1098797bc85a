// integration/kanzi_gpu_adapters.hpp -- reference-side bindings of libknzgpu.so.
//
// This header is what a kanzi maintainer adds to the reference tree: three adapter
// classes deriving from the reference's own plugin interfaces and forwarding to the
// C ABI (include/knz_gpu.h).  It is compile-checked against the unmodified reference
// headers by __graft_entry__.build() when /root/reference is present; it is not part
// of the product library (the reference headers do not travel to the GPU box).
//
//   GpuTransform      : kanzi::Transform<byte>   (src/Transform.hpp:31-48)
//   GpuEntropyEncoder : kanzi::EntropyEncoder    (src/EntropyEncoder.hpp:25-40)
//   GpuEntropyDecoder : kanzi::EntropyDecoder    (src/EntropyDecoder.hpp:25-40)
//
// Stage-level drop-in: return these from the three factory switches
//   TransformFactory<T>::newToken        transform/TransformFactory.hpp:225-308
//   EntropyEncoderFactory::newEncoder    entropy/EntropyEncoderFactory.hpp:62-97
//   EntropyDecoderFactory::newDecoder    entropy/EntropyDecoderFactory.hpp:62-97
// Block-level drop-in (the one that performs): EncodingTask<T>::run /
// DecodingTask<T>::run call knz_encode_blocks / knz_decode_blocks for a batch of
// blocks instead of building a TransformSequence + codec per block.
#pragma once
#include <stdexcept>
#include <vector>

#include "EntropyDecoder.hpp"
#include "EntropyEncoder.hpp"
#include "InputBitStream.hpp"
#include "OutputBitStream.hpp"
#include "SliceArray.hpp"
#include "Transform.hpp"
#include "knz_gpu.h"

namespace kanzi {

// One process-wide context per device; the C ABI serialises calls per context.
inline knz_ctx* knzSharedContext(int maxBlockSize = 4 * 1024 * 1024)
{
    static knz_ctx* ctx = nullptr;
    if (ctx == nullptr && knz_create(0, maxBlockSize, 64, &ctx) != KNZ_OK)
        throw std::runtime_error("knz_create failed: no usable CUDA device (there is no CPU fallback)");
    return ctx;
}

class GpuTransform FINAL : public Transform<byte> {
public:
    // type = wire id (TransformFactory.hpp:49-73): BWT_TYPE, RANK_TYPE, MTFT_TYPE, ZRLT_TYPE
    GpuTransform(int type, knz_ctx* ctx) : _type(type), _ctx(ctx) {}
    ~GpuTransform() {}

    bool forward(SliceArray<byte>& src, SliceArray<byte>& dst, int length) { return run(src, dst, length, false); }
    bool inverse(SliceArray<byte>& src, SliceArray<byte>& dst, int length) { return run(src, dst, length, true); }
    int getMaxEncodedLength(int srcLen) const { return (_type == KNZ_T_BWT) ? srcLen + 33 : srcLen; }

private:
    int _type;
    knz_ctx* _ctx;

    bool run(SliceArray<byte>& src, SliceArray<byte>& dst, int length, bool inv)
    {
        if (length == 0)
            return true;
        if (!SliceArray<byte>::isValid(src))
            throw std::invalid_argument("GpuTransform: Invalid input block");
        if (!SliceArray<byte>::isValid(dst))
            throw std::invalid_argument("GpuTransform: Invalid output block");
        if (length < 0 || length > src._length - src._index)
            return false;
        int outLen = 0, applied = 0;
        const uint8_t* in = reinterpret_cast<const uint8_t*>(&src._array[src._index]);
        uint8_t* out = reinterpret_cast<uint8_t*>(&dst._array[dst._index]);
        const int cap = dst._length - dst._index;
        const int rc = inv ? knz_transform_inverse(_ctx, _type, in, length, out, cap, &outLen, &applied)
                           : knz_transform_forward(_ctx, _type, in, length, out, cap, &outLen, &applied);
        if (rc != KNZ_OK || applied == 0)
            return false; // "stage not applicable": the sequence skips it (TransformSequence.hpp:132-138)
        src._index += length;
        dst._index += outLen;
        return true;
    }
};

class GpuEntropyEncoder FINAL : public EntropyEncoder {
public:
    GpuEntropyEncoder(OutputBitStream& obs, int type, knz_ctx* ctx) : _obs(obs), _type(type), _ctx(ctx) {}
    ~GpuEntropyEncoder() {}

    // Encodes the whole block on the device and appends the resulting bit string
    // to the task's bitstream (bit-aligned, like ANSRangeEncoder::encode).
    int encode(const byte block[], uint blkptr, uint len)
    {
        if (len == 0)
            return 0;
        _buf.resize(size_t(len) + (len >> 2) + 8192);
        int64_t bits = 0;
        if (knz_entropy_encode(_ctx, _type, reinterpret_cast<const uint8_t*>(&block[blkptr]), int(len), _buf.data(),
                               int64_t(_buf.size()), &bits) != KNZ_OK)
            return -1;
        _obs.writeBits(reinterpret_cast<const byte*>(_buf.data()), uint(bits));
        return int(len);
    }
    OutputBitStream& getBitStream() const { return _obs; }
    void dispose() {}

private:
    OutputBitStream& _obs;
    int _type;
    knz_ctx* _ctx;
    std::vector<uint8_t> _buf;
};

class GpuEntropyDecoder FINAL : public EntropyDecoder {
public:
    // `availableBits`: bits left in the task's private stream (DecodingTask copies each
    // block into its own buffer first: io/CompressedInputStream.cpp:843-856).
    GpuEntropyDecoder(InputBitStream& ibs, int type, knz_ctx* ctx, uint64 availableBits)
        : _ibs(ibs), _type(type), _ctx(ctx), _avail(availableBits)
    {
    }
    ~GpuEntropyDecoder() {}

    int decode(byte block[], uint blkptr, uint len)
    {
        if (len == 0)
            return 0;
        const size_t nbytes = size_t((_avail + 7) >> 3);
        _buf.assign(nbytes + 16, 0);
        _ibs.readBits(reinterpret_cast<byte*>(_buf.data()), uint(_avail));
        if (knz_entropy_decode(_ctx, _type, _buf.data(), int64_t(_avail), reinterpret_cast<uint8_t*>(&block[blkptr]),
                               int(len)) != KNZ_OK)
            return -1;
        return int(len);
    }
    InputBitStream& getBitStream() const { return _ibs; }
    void dispose() {}

private:
    InputBitStream& _ibs;
    int _type;
    knz_ctx* _ctx;
    uint64 _avail;
    std::vector<uint8_t> _buf;
};

} // namespace kanzi
