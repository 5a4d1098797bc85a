// cusim.h -- TEST INFRASTRUCTURE: a tiny single-process CUDA execution-model
// emulator so the kernels' *logic* (indexing, scans, ballots, bit packing) can be
// exercised on the CPU-only build box before GPU time is spent.  The product
// library never includes this file; tests/sim/Makefile compiles the very same
// .cu sources as C++ with -DKNZ_SIM -include cusim.h into tests/sim/libknzsim.so.
//
// Model: CTAs run one after another; inside a CTA every CUDA thread is a fiber
// (hand-rolled x86-64 context switch) scheduled round-robin.  Fibers only yield
// inside __syncthreads / warp collectives, so everything else is sequentially
// consistent (races between barriers are NOT detected -- that is what
// compute-sanitizer on the B200 box is for).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint4 {
    unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{ a, b, c, d }; }
struct uint2 {
    unsigned x, y;
};
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{ a, b }; }
struct ulonglong2 {
    unsigned long long x, y;
};
inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { return ulonglong2{ a, b }; }

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
#define cudaSuccess 0
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaHostAllocDefault = 0 };

namespace cusim {

struct Fiber {
    void* rsp = nullptr;
    char* stack = nullptr;
    dim3 tIdx;
    int lin = 0;
    bool done = false;
};

struct WarpState {
    uint32_t alive = 0, arrived = 0, draining = 0, snapMask = 0;
    uint64_t gen = 0;
    uint64_t vals[32], snap[32];
};

struct Global {
    Fiber* cur = nullptr;
    dim3 bIdx, bDim, gDim;
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    void* schedRsp = nullptr;
    int alive = 0, barArrived = 0;
    uint64_t barGen = 0;
    const std::function<void()>* body = nullptr;
    size_t stackSize = 256 * 1024;
};

extern Global g;
extern "C" void cusim_switch(void** from, void* to);
void yield();
void launch(dim3 grid, dim3 block, const std::function<void()>& body);

inline int lane() { return g.cur->lin & 31; }
inline WarpState& warp() { return g.warps[g.cur->lin >> 5]; }

// Generic warp rendezvous: deposit a value, wait for every live lane in `mask`,
// then evaluate `f` on the snapshot.
template <class F>
inline uint64_t collective(uint32_t mask, uint64_t v, F f)
{
    WarpState& W = warp();
    const int ln = lane();
    const uint32_t bit = 1u << ln;
    while (W.draining != 0)
        yield();
    W.vals[ln] = v;
    W.arrived |= bit;
    const uint32_t need = mask & W.alive;
    if ((W.arrived & need) == need) {
        memcpy(W.snap, W.vals, sizeof(W.vals));
        W.snapMask = need;
        W.draining = need;
        W.arrived = 0;
        W.gen++;
    } else {
        const uint64_t gen0 = W.gen;
        while (W.gen == gen0)
            yield();
    }
    const uint64_t r = f(W.snap, W.snapMask, ln);
    W.draining &= ~bit;
    return r;
}

} // namespace cusim

#define threadIdx (cusim::g.cur->tIdx)
#define blockIdx (cusim::g.bIdx)
#define blockDim (cusim::g.bDim)
#define gridDim (cusim::g.gDim)

inline void __syncthreads()
{
    using namespace cusim;
    const uint64_t gen0 = g.barGen;
    g.barArrived++;
    if (g.barArrived >= g.alive) {
        g.barArrived = 0;
        g.barGen++;
    } else {
        while (g.barGen == gen0)
            yield();
    }
}

inline void __syncwarp(uint32_t mask = 0xFFFFFFFFu)
{
    cusim::collective(mask, 0, [](const uint64_t*, uint32_t, int) { return (uint64_t)0; });
}

inline uint32_t __ballot_sync(uint32_t mask, int pred)
{
    return (uint32_t)cusim::collective(mask, pred ? 1 : 0, [](const uint64_t* s, uint32_t m, int) {
        uint32_t r = 0;
        for (int i = 0; i < 32; i++)
            if (((m >> i) & 1) && s[i])
                r |= 1u << i;
        return (uint64_t)r;
    });
}
inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(uint32_t mask, int pred)
{
    return (int)cusim::collective(mask, pred ? 1 : 0, [](const uint64_t* s, uint32_t m, int) {
        for (int i = 0; i < 32; i++)
            if (((m >> i) & 1) && !s[i])
                return (uint64_t)0;
        return (uint64_t)1;
    });
}
inline uint32_t __match_any_sync(uint32_t mask, uint64_t v)
{
    return (uint32_t)cusim::collective(mask, v, [](const uint64_t* s, uint32_t m, int ln) {
        uint32_t r = 0;
        for (int i = 0; i < 32; i++)
            if (((m >> i) & 1) && s[i] == s[ln])
                r |= 1u << i;
        return (uint64_t)r;
    });
}

template <class T>
inline T __shfl_sync(uint32_t mask, T v, int src)
{
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t r = cusim::collective(mask, raw, [src](const uint64_t* s, uint32_t, int) { return s[src & 31]; });
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <class T>
inline T __shfl_up_sync(uint32_t mask, T v, unsigned d)
{
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t r = cusim::collective(mask, raw, [d](const uint64_t* s, uint32_t, int ln) {
        return (ln >= (int)d) ? s[ln - d] : s[ln];
    });
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <class T>
inline T __shfl_down_sync(uint32_t mask, T v, unsigned d)
{
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t r = cusim::collective(mask, raw, [d](const uint64_t* s, uint32_t, int ln) {
        return (ln + (int)d < 32) ? s[ln + d] : s[ln];
    });
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <class T>
inline T __shfl_xor_sync(uint32_t mask, T v, int x)
{
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t r = cusim::collective(mask, raw, [x](const uint64_t* s, uint32_t, int ln) { return s[(ln ^ x) & 31]; });
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}

inline uint32_t __reduce_or_sync(uint32_t mask, uint32_t v)
{
    return (uint32_t)cusim::collective(mask, v, [](const uint64_t* s, uint32_t m, int) {
        uint64_t r = 0;
        for (int i = 0; i < 32; i++)
            if ((m >> i) & 1)
                r |= s[i];
        return r;
    });
}
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __popcll(uint64_t x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t __brev(uint32_t x)
{
    uint32_t r = 0;
    for (int i = 0; i < 32; i++)
        if ((x >> i) & 1)
            r |= 1u << (31 - i);
    return r;
}
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const int k = (sel >> (4 * i)) & 7;
        r |= (uint32_t)((v >> (8 * k)) & 0xFF) << (8 * i);
    }
    return r;
}
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh)
{
    sh &= 31;
    return sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    sh &= 31;
    return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
}
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T min(T a, T b) { return a < b ? a : b; }
template <class T>
inline T max(T a, T b) { return a > b ? a : b; }

template <class T>
inline T atomicAdd(T* p, T v)
{
    T o = *p;
    *p = o + v;
    return o;
}
template <class T>
inline T atomicOr(T* p, T v)
{
    T o = *p;
    *p = o | v;
    return o;
}
template <class T>
inline T atomicMax(T* p, T v)
{
    T o = *p;
    if (v > o)
        *p = v;
    return o;
}
template <class T>
inline T atomicMin(T* p, T v)
{
    T o = *p;
    if (v < o)
        *p = v;
    return o;
}
template <class T>
inline T atomicExch(T* p, T v)
{
    T o = *p;
    *p = v;
    return o;
}
template <class T>
inline T atomicCAS(T* p, T cmp, T v)
{
    T o = *p;
    if (o == cmp)
        *p = v;
    return o;
}
inline void __threadfence() {}
inline void __threadfence_block() {}

// ---- CUDA runtime stubs (host memory plays the device)
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, int, cudaStream_t)
{
    for (size_t i = 0; i < h; i++)
        memmove((char*)d + i * dp, (const char*)s + i * sp, w);
    return 0;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemset2DAsync(void* d, size_t pitch, int v, size_t w, size_t h, cudaStream_t)
{
    for (size_t r = 0; r < h; r++)
        memset((char*)d + r * pitch, v, w);
    return 0;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "cusim"; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
#define cudaEventDisableTiming 2
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }

#define KLAUNCH(kernel, grid, block, stream, ...) cusim::launch(grid, block, [&]() { kernel(__VA_ARGS__); })
#define KLAUNCH_DYN(kernel, grid, block, smem, stream, ...) cusim::launch(grid, block, [&]() { kernel(__VA_ARGS__); })
#define KNZ_DYN_SMEM(name) static __attribute__((aligned(16))) unsigned char name[232 * 1024]
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
