// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Thin extern "C" shim over the UNMODIFIED reference sources under
// /root/reference/src (compiled where they lie by oracle/Makefile into
// oracle/_ref/libkanzi_ref.so).  It exposes the reference's own classes for
// the hot path so that (a) the C restatement in oracle/kanzi_oracle.c can be
// pinned against the real thing, (b) golden fixtures under tests/golden/ can
// be generated, and (c) bench.py --impl reference / cpu_baseline can time the
// reference's multi-threaded CPU path.  Nothing in the product library links
// or loads this file.
//
// Reference entry points used:
//   kanzi::CompressedOutputStream  src/io/CompressedOutputStream.hpp:136
//   kanzi::CompressedInputStream   src/io/CompressedInputStream.hpp:177
//   kanzi::TransformFactory<byte>  src/transform/TransformFactory.hpp:208
//   kanzi::EntropyEncoderFactory   src/entropy/EntropyEncoderFactory.hpp:62
//   kanzi::EntropyDecoderFactory   src/entropy/EntropyDecoderFactory.hpp:62
//   kanzi::BWT                     src/transform/BWT.hpp:59
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#include <streambuf>
#include <cstdio>

#include "types.hpp"
#include "Context.hpp"
#include "SliceArray.hpp"
#include "io/CompressedOutputStream.hpp"
#include "io/CompressedInputStream.hpp"
#include "transform/TransformFactory.hpp"
#include "transform/BWT.hpp"
#include "entropy/EntropyEncoderFactory.hpp"
#include "entropy/EntropyDecoderFactory.hpp"
#include "bitstream/DefaultOutputBitStream.hpp"
#include "bitstream/DefaultInputBitStream.hpp"

using namespace kanzi;

namespace {

// Fixed-capacity output streambuf over caller memory (no reallocation).
class FixedOut : public std::streambuf {
public:
    FixedOut(char* p, size_t n) { setp(p, p + n); }
    size_t count() const { return size_t(pptr() - pbase()); }
};

class FixedIn : public std::streambuf {
public:
    FixedIn(const char* p, size_t n)
    {
        char* q = const_cast<char*>(p);
        setg(q, q, q + n);
    }
};

} // namespace

extern "C" {

// Whole-stream compression through the reference's CompressedOutputStream.
// Returns 0 on success, <0 on exception / overflow.
int kref_stream_compress(const uint8_t* in, int64_t n, const char* transform, const char* entropy,
                         int blockSize, int jobs, int checksum, uint8_t* out, int64_t cap, int64_t* outLen)
{
    try {
        FixedOut ob(reinterpret_cast<char*>(out), size_t(cap));
        std::ostream os(&ob);
        {
            CompressedOutputStream cos(os, jobs, entropy, transform, blockSize, checksum, uint64(n));
            int64_t off = 0;

            while (off < n) {
                const int64_t len = (n - off < (int64_t(1) << 24)) ? (n - off) : (int64_t(1) << 24);
                cos.write(reinterpret_cast<const char*>(in + off), std::streamsize(len));
                off += len;
            }

            cos.close();
        }

        if (os.fail())
            return -2;

        *outLen = int64_t(ob.count());
        return 0;
    }
    catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] %s\n", e.what());
        return -1;
    }
}

int kref_stream_decompress(const uint8_t* in, int64_t n, int jobs, uint8_t* out, int64_t cap, int64_t* outLen)
{
    try {
        FixedIn ib(reinterpret_cast<const char*>(in), size_t(n));
        std::istream is(&ib);
        CompressedInputStream cis(is, jobs);
        int64_t off = 0;

        while (off < cap) {
            const int64_t want = (cap - off < (int64_t(1) << 24)) ? (cap - off) : (int64_t(1) << 24);
            cis.read(reinterpret_cast<char*>(out + off), std::streamsize(want));
            const int64_t got = int64_t(cis.gcount());
            off += got;

            if (got < want)
                break;
        }

        cis.close();
        *outLen = off;
        return 0;
    }
    catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] %s\n", e.what());
        return -1;
    }
}

// Transform sequence forward ("BWT+RANK+ZRLT", "ZRLT", ...) exactly as
// EncodingTask builds it (TransformFactory::newTransform).  `cap` is the
// destination capacity handed to the sequence (SliceArray::_length).
// srcCap is the capacity of the source SliceArray (it is used as a ping-pong
// buffer by the sequence, so it matters for capacity-dependent stages).
// Returns 1 if forward() returned true, 0 if false, <0 on exception.
int kref_transform_forward(const char* name, const uint8_t* in, int n, int srcCap, uint8_t* out, int cap,
                           int* outLen, int* skipFlags)
{
    try {
        Context ctx;
        ctx.putInt("bsVersion", 6);
        ctx.putInt("size", n);
        ctx.putString("entropy", "ANS0");
        const uint64 tt = TransformFactory<byte>::getType(name);
        TransformSequence<byte>* seq = TransformFactory<byte>::newTransform(ctx, tt);
        std::vector<byte> src(size_t(srcCap < n ? n : srcCap));
        memcpy(src.data(), in, size_t(n));
        SliceArray<byte> sa1(src.data(), int(src.size()), 0);
        SliceArray<byte> sa2(reinterpret_cast<byte*>(out), cap, 0);
        const bool ok = seq->forward(sa1, sa2, n);
        *outLen = sa2._index;
        *skipFlags = int(seq->getSkipFlags());
        delete seq;
        return ok ? 1 : 0;
    }
    catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] %s\n", e.what());
        return -1;
    }
}

int kref_transform_inverse(const char* name, int skipFlags, const uint8_t* in, int n, uint8_t* out, int cap, int* outLen)
{
    try {
        Context ctx;
        ctx.putInt("bsVersion", 6);
        ctx.putInt("size", n);
        ctx.putString("entropy", "ANS0");
        const uint64 tt = TransformFactory<byte>::getType(name);
        TransformSequence<byte>* seq = TransformFactory<byte>::newTransform(ctx, tt);
        seq->setSkipFlags(byte(skipFlags));
        std::vector<byte> src(size_t(n) + 64);
        memcpy(src.data(), in, size_t(n));
        SliceArray<byte> sa1(src.data(), int(src.size()), 0);
        SliceArray<byte> sa2(reinterpret_cast<byte*>(out), cap, 0);
        const bool ok = seq->inverse(sa1, sa2, n);
        *outLen = sa2._index;
        delete seq;
        return ok ? 1 : 0;
    }
    catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] %s\n", e.what());
        return -1;
    }
}

// Raw BWT (no block header): output bytes + the 8 primary indexes.
int kref_bwt_forward(const uint8_t* in, int n, uint8_t* out, int* primaryIndexes)
{
    try {
        BWT bwt(1);
        std::vector<byte> src(size_t(n) + 8);
        memcpy(src.data(), in, size_t(n));
        SliceArray<byte> sa1(src.data(), n, 0);
        SliceArray<byte> sa2(reinterpret_cast<byte*>(out), n, 0);
        const bool ok = bwt.forward(sa1, sa2, n);

        for (int i = 0; i < 8; i++)
            primaryIndexes[i] = bwt.getPrimaryIndex(i);

        return ok ? 1 : 0;
    }
    catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] %s\n", e.what());
        return -1;
    }
}

// Entropy stage alone: encoder writes into a private DefaultOutputBitStream,
// exactly like EncodingTask::run does after the block header.  outBits gets
// obs.written() after close().
int kref_entropy_encode(const char* name, const uint8_t* in, int n, uint8_t* out, int64_t cap, int64_t* outBits)
{
    try {
        FixedOut ob(reinterpret_cast<char*>(out), size_t(cap));
        std::ostream os(&ob);
        DefaultOutputBitStream obs(os, 65536);
        Context ctx;
        ctx.putInt("bsVersion", 6);
        const short et = EntropyEncoderFactory::getType(name);
        EntropyEncoder* ee = EntropyEncoderFactory::newEncoder(obs, ctx, et);
        const int res = ee->encode(reinterpret_cast<const byte*>(in), 0, uint(n));
        ee->dispose();
        delete ee;
        obs.close();
        *outBits = int64_t(obs.written());

        if (os.fail())
            return -2;

        return (res == n) ? 1 : 0;
    }
    catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] %s\n", e.what());
        return -1;
    }
}

int kref_entropy_decode(const char* name, const uint8_t* in, int64_t nbytes, uint8_t* out, int n, int64_t* bitsRead)
{
    try {
        FixedIn ib(reinterpret_cast<const char*>(in), size_t(nbytes));
        std::istream is(&ib);
        DefaultInputBitStream ibs(is, 65536);
        Context ctx;
        ctx.putInt("bsVersion", 6);
        const short et = EntropyDecoderFactory::getType(name);
        EntropyDecoder* ed = EntropyDecoderFactory::newDecoder(ibs, ctx, et);
        const int res = ed->decode(reinterpret_cast<byte*>(out), 0, uint(n));
        ed->dispose();
        delete ed;
        *bitsRead = int64_t(ibs.read());
        return (res == n) ? 1 : 0;
    }
    catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] %s\n", e.what());
        return -1;
    }
}

} // extern "C"
