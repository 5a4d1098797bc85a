// integration/kanzi_gpu_adapters.hpp -- reference-side bindings of libknzgpu.so.
//
// This header is what a kanzi maintainer adds to the reference tree: three adapter
// classes deriving from the reference's own plugin interfaces and forwarding to the
// C ABI (include/knz_gpu.h).  It is compiled against the unmodified reference headers
// by oracle/Makefile (target refgpu: the reference's stream classes rebuilt with the
// three factory switches routed here, integration/knz_reference_hooks.hpp) and executed
// on the GPU box by tests/test_gpu_adapters.py; it is not part of the product library.
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
#include <mutex>
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

// One process-wide context per device, shared by every worker thread of the stream classes
// (up to 64 tasks, io/CompressedOutputStream.cpp:40,512-525): every C-ABI entry point locks its
// context, so concurrent stage calls serialise on the device.  Stage-level calls handle one
// block at a time: a batch capacity of 2 is enough.  Provisioned for the largest block size seen.
inline knz_ctx* knzSharedContext(int maxBlockSize = 4 * 1024 * 1024)
{
    static std::mutex mtx;
    static knz_ctx* ctx = nullptr;
    static int provisioned = 0;
    std::lock_guard<std::mutex> lock(mtx);
    maxBlockSize = (maxBlockSize + 15) & ~15;
    if (maxBlockSize < 65536)
        maxBlockSize = 65536;
    if (ctx != nullptr && maxBlockSize > provisioned) {
        knz_destroy(ctx); // nobody holds a stage call across block sizes: streams are created one after another
        ctx = nullptr;
    }
    if (ctx == nullptr) {
        if (knz_create(0, maxBlockSize, 2, &ctx) != KNZ_OK)
            throw std::runtime_error("knz_create failed: no usable CUDA device (there is no CPU fallback)");
        provisioned = maxBlockSize;
    }
    return ctx;
}

class GpuTransform FINAL : public Transform<byte> {
public:
    // type = wire id (TransformFactory.hpp:49-73): BWT_TYPE, RANK_TYPE, MTFT_TYPE, ZRLT_TYPE, SRT_TYPE
    GpuTransform(int type, knz_ctx* ctx) : _type(type), _ctx(ctx) {}
    ~GpuTransform() {}

    bool forward(SliceArray<byte>& src, SliceArray<byte>& dst, int length) { return run(src, dst, length, false); }
    bool inverse(SliceArray<byte>& src, SliceArray<byte>& dst, int length) { return run(src, dst, length, true); }
    int getMaxEncodedLength(int srcLen) const
    {
        if (_type == KNZ_T_LZ || _type == KNZ_T_LZX || _type == KNZ_T_LZP) // transform/LZCodec.hpp:91-95, :158-161
            return ((srcLen <= 1024) ? srcLen + 16 : srcLen + (srcLen / 64)) + ((_type == KNZ_T_LZP) ? 0 : 2);
        return (_type == KNZ_T_BWT) ? srcLen + 33 : (_type == KNZ_T_SRT) ? srcLen + 1024 : srcLen;
    }

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
        _buf.resize(size_t(len) + (len >> 2) + 8192 + 140000 * size_t((len >> 22) + 1));
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
    // The device decodes a whole block payload at once, so the adapter hands it every bit the block
    // has left.  `availableBits` > 0: that many bits are read (a one-line patch in DecodingTask::run can
    // publish them through the Context).  availableBits == 0: the stream is the
    // task's PRIVATE per-block stream (DecodingTask copies each block into its own buffer when jobs > 1:
    // io/CompressedInputStream.cpp:843-856,870-872) and is read to its end; trailing padding bits are
    // harmless because every codec knows its own lengths.
    GpuEntropyDecoder(InputBitStream& ibs, int type, knz_ctx* ctx, uint64 availableBits = 0)
        : _ibs(ibs), _type(type), _ctx(ctx), _avail(availableBits)
    {
    }
    ~GpuEntropyDecoder() {}

    int decode(byte block[], uint blkptr, uint len)
    {
        if (len == 0)
            return 0;
        uint64 bits = _avail;
        if (bits > 0) {
            _buf.assign(size_t((bits + 7) >> 3) + 16, 0);
            _ibs.readBits(reinterpret_cast<byte*>(_buf.data()), uint(bits));
        } else {
            _buf.clear();
            while (_ibs.hasMoreToRead())
                _buf.push_back(uint8_t(_ibs.readBits(8)));
            bits = uint64(_buf.size()) * 8;
            _buf.resize(_buf.size() + 16, 0);
        }
        if (knz_entropy_decode(_ctx, _type, _buf.data(), int64_t(bits), reinterpret_cast<uint8_t*>(&block[blkptr]),
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
