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
#include <mutex>

#include <chrono>
#include <map>
#include <iomanip>

#include "types.hpp"
#include "util/WallTimer.hpp"
// Event::_skipFlags has no accessor (it only appears in Event::toString()); the shim reads the member
// itself.  Access control does not change the layout of the unmodified class.
#define private public
#include "Event.hpp"
#undef private
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

protected:
    // tellp(): the reference reports bit offsets of blocks through OutputBitStream::tell()
    pos_type seekoff(off_type off, std::ios_base::seekdir dir, std::ios_base::openmode) override
    {
        return (dir == std::ios_base::cur && off == 0) ? pos_type(pptr() - pbase()) : pos_type(off_type(-1));
    }
};

class FixedIn : public std::streambuf {
public:
    FixedIn(const char* p, size_t n)
    {
        char* q = const_cast<char*>(p);
        setg(q, q, q + n);
    }

protected:
    pos_type seekoff(off_type off, std::ios_base::seekdir dir, std::ios_base::openmode) override
    {
        return (dir == std::ios_base::cur && off == 0) ? pos_type(gptr() - eback()) : pos_type(off_type(-1));
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

// The same two calls through the Context constructors, which is how the reference's application sets
// `skipBlocks`, `from` / `to` and attaches listeners (app/BlockCompressor.cpp, app/BlockDecompressor.cpp).
// Every event a listener receives is recorded as seven int64: type, id, size, hash, hash type, offset; the
// skip flags of BLOCK_INFO events are the seventh.
struct RecListener : public Listener<Event> {
    int64_t* buf;
    int cap, n;
    std::mutex m;
    RecListener(int64_t* b, int c) : buf(b), cap(c), n(0) {}
    void processEvent(const Event& e)
    {
        std::lock_guard<std::mutex> g(m);
        if (buf == nullptr || n >= cap)
            return;
        int64_t* r = buf + 7 * n++;
        r[0] = int64_t(e.getType());
        r[1] = e.getId();
        r[2] = e.getSize();
        r[3] = int64_t(e.getHash());
        r[4] = int64_t(e.getHashType());
        r[5] = e.getOffset();
        r[6] = 0;
        if (e.getType() == Event::BLOCK_INFO)
            r[6] = int64_t(e._skipFlags);
    }
};

int kref_stream_compress_ctx(const uint8_t* in, int64_t n, const char* transform, const char* entropy, int blockSize,
                             int jobs, int checksum, int skipBlocks, uint8_t* out, int64_t cap, int64_t* outLen,
                             int64_t* events, int evCap, int* evCount)
{
    try {
        FixedOut ob(reinterpret_cast<char*>(out), size_t(cap));
        std::ostream os(&ob);
        RecListener rec(events, evCap);
        {
            Context ctx;
            ctx.putInt("jobs", jobs);
            ctx.putInt("blockSize", blockSize);
            ctx.putInt("checksum", checksum);
            ctx.putInt("skipBlocks", skipBlocks);
            ctx.putInt("verbosity", 5); // BLOCK_INFO events are only built above 4
            ctx.putLong("fileSize", n);
            ctx.putString("entropy", entropy);
            ctx.putString("transform", transform);
            CompressedOutputStream cos(os, ctx);
            if (events != nullptr)
                cos.addListener(rec);
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
        if (evCount)
            *evCount = rec.n;
        return 0;
    }
    catch (const std::exception& e) {
        fprintf(stderr, "[ref_shim] %s\n", e.what());
        return -1;
    }
}

int kref_stream_decompress_ctx(const uint8_t* in, int64_t n, int jobs, int from, int to, uint8_t* out, int64_t cap,
                               int64_t* outLen, int64_t* events, int evCap, int* evCount)
{
    try {
        FixedIn ib(reinterpret_cast<const char*>(in), size_t(n));
        std::istream is(&ib);
        RecListener rec(events, evCap);
        Context ctx;
        ctx.putInt("jobs", jobs);
        ctx.putInt("verbosity", 5);
        if (from > 0)
            ctx.putInt("from", from);
        if (to > 0)
            ctx.putInt("to", to);
        CompressedInputStream cis(is, ctx);
        if (events != nullptr)
            cis.addListener(rec);
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
        if (evCount)
            *evCount = rec.n;
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
