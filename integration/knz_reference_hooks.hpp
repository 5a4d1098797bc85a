// integration/knz_reference_hooks.hpp -- routes the reference's three factory switches to the GPU
// adapters WITHOUT editing a reference file.
//
// oracle/Makefile (target refgpu) compiles the reference's stream classes where they lie
// (io/CompressedOutputStream.cpp, io/CompressedInputStream.cpp) and the test shim with
//     g++ ... -include integration/knz_reference_hooks.hpp
// This header pulls in the reference's own factory headers first (their include guards then keep the
// .cpp files from seeing them again), declares drop-in factory classes with the same static
// interface, and renames the factory identifiers for the rest of the translation unit.  The effect
// is exactly the patch INTEGRATION.md section 2 describes -- "return the adapter from the switch for
// the ids the GPU implements, keep the reference's own codec for the rest":
//   TransformFactory<T>::newTransform    transform/TransformFactory.hpp:208-223 (newToken :225-308)
//   EntropyEncoderFactory::newEncoder    entropy/EntropyEncoderFactory.hpp:62-97
//   EntropyDecoderFactory::newDecoder    entropy/EntropyDecoderFactory.hpp:62-97
// EncodingTask::run / DecodingTask::run then drive the GPU stages through the reference's own
// TransformSequence, block framing, bitstream and task pool.
#pragma once
#include "transform/TransformFactory.hpp"
#include "entropy/EntropyEncoderFactory.hpp"
#include "entropy/EntropyDecoderFactory.hpp"
#include "kanzi_gpu_adapters.hpp"

namespace kanzi {

inline bool knzGpuTransformId(uint64 t)
{
    return t == KNZ_T_BWT || t == KNZ_T_RANK || t == KNZ_T_MTFT || t == KNZ_T_ZRLT || t == KNZ_T_SRT || t == KNZ_T_LZ ||
           t == KNZ_T_LZX || t == KNZ_T_LZP;
}

inline bool knzGpuEntropyId(short e)
{
    return e == KNZ_E_ANS0 || e == KNZ_E_ANS1 || e == KNZ_E_HUFFMAN || e == KNZ_E_FPAQ;
}

template <class T>
class GpuTransformFactory {
public:
    enum { NONE_TYPE = TransformFactory<T>::NONE_TYPE };
    static uint64 getType(const char* name) { return TransformFactory<T>::getType(name); }
    static std::string getName(uint64 type) { return TransformFactory<T>::getName(type); }

    // Same slot logic as TransformFactory<T>::newTransform (:208-223): slot 0 always, later NONE slots dropped.
    static TransformSequence<T>* newTransform(Context& ctx, uint64 functionType)
    {
        Transform<T>* transforms[8];
        int n = 0;
        const int blockSize = ctx.getInt("blockSize", 4 * 1024 * 1024);
        for (int i = 0; i < 8; i++) {
            transforms[i] = nullptr;
            const uint64 t = (functionType >> (42 - 6 * i)) & 63;
            if ((t == TransformFactory<T>::NONE_TYPE) && (i != 0))
                continue;
            if (knzGpuTransformId(t))
                transforms[n++] = new GpuTransform(int(t), knzSharedContext(blockSize));
            else if (t == TransformFactory<T>::NONE_TYPE)
                transforms[n++] = new NullTransform(ctx);
            else
                throw std::invalid_argument("transform not routed in the GPU check build");
        }
        return new TransformSequence<T>(transforms, true);
    }
};

class GpuEntropyEncoderFactory {
public:
    static const short NONE_TYPE = 0;
    static short getType(const char* name) { return EntropyEncoderFactory::getType(name); }
    static const char* getName(short type) { return EntropyEncoderFactory::getName(type); }
    static EntropyEncoder* newEncoder(OutputBitStream& obs, Context& ctx, short type)
    {
        if (knzGpuEntropyId(type))
            return new GpuEntropyEncoder(obs, type, knzSharedContext(ctx.getInt("blockSize", 4 * 1024 * 1024)));
        return EntropyEncoderFactory::newEncoder(obs, ctx, type);
    }
};

class GpuEntropyDecoderFactory {
public:
    static const short NONE_TYPE = 0;
    static short getType(const char* name) { return EntropyDecoderFactory::getType(name); }
    static const char* getName(short type) { return EntropyDecoderFactory::getName(type); }
    static EntropyDecoder* newDecoder(InputBitStream& ibs, Context& ctx, short type)
    {
        if (knzGpuEntropyId(type)) // bits of the block, when the host published them; else the private stream is read out
            return new GpuEntropyDecoder(ibs, type, knzSharedContext(ctx.getInt("blockSize", 4 * 1024 * 1024)),
                                         uint64(ctx.getLong("knzBlockBits", 0)));
        return EntropyDecoderFactory::newDecoder(ibs, ctx, type);
    }
};

} // namespace kanzi

#define TransformFactory GpuTransformFactory
#define EntropyEncoderFactory GpuEntropyEncoderFactory
#define EntropyDecoderFactory GpuEntropyDecoderFactory
