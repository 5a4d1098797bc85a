// lat.cu -- dependent-chain latencies of the warp primitives the SBRT replay is built from
// (one warp, cycles per link).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32;
#define N 4096
#define FULL 0xffffffffu

template <int WHICH>
__global__ void chain(u32* out, long long* cyc, u32 seed, u32 c1, u32 c2)
{
    __shared__ u32 sm[1024];
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32)
        sm[i] = (i * 37 + 11) & 1023;
    __syncwarp();
    u32 x = seed + lane, y = seed * 3 + lane;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (WHICH == 0) x = __shfl_sync(FULL, x, x & 31) + 1;               // SHFL.IDX + IADD
        if (WHICH == 1) x = __shfl_up_sync(FULL, x, 1) + 1;                 // SHFL.UP + IADD
        if (WHICH == 2) x = x + c1;                                         // IADD
        if (WHICH == 3) x = (x <= y) ? c1 : c2;                             // ISETP -> SEL (values alternate)
        if (WHICH == 4) x = max(y, min(x + 1, c2));                         // IADD, IMNMX, IMNMX
        if (WHICH == 5) x = (x >> 1) ^ c1;                                  // SHF + LOP (or LOP3)
        if (WHICH == 6) x = __popc(__ballot_sync(FULL, x > (u32)lane)) + c1; // ISETP VOTE POPC IADD
        if (WHICH == 7) x = sm[x & 1023];                                   // LDS chase
        if (WHICH == 8) { if (x & 1) x = x * 3 + 1; else x = x >> 1; if (x == 1) x = c2; } // branches (uniform when seed uniform)
        if (WHICH == 9) x = __reduce_or_sync(FULL, x) + 1;                  // REDUX
        if (WHICH == 10) { u32 e = __shfl_sync(FULL, y, x & 31); u32 Y = (e >> 8) + i; bool a = (x <= Y), b = (y <= Y); if (a) { x = b ? y : (Y & ~1u); y = b ? e : (u32)(i << 8); } } // model step
        if (WHICH == 11) x = __shfl_sync(FULL, x, 0) + 1;                   // SHFL.IDX const lane
    }
    long long t1 = clock64();
    out[lane] = x + y;
    if (lane == 0)
        cyc[WHICH] = t1 - t0;
}

int main()
{
    u32* out;
    long long* cyc;
    cudaMalloc(&out, 128);
    cudaMallocManaged(&cyc, 16 * 8);
    const char* names[] = { "SHFL.IDX+IADD", "SHFL.UP+IADD", "IADD", "ISETP+SEL", "IADD+2xIMNMX", "SHF+LOP", "ISETP+VOTE+POPC+IADD",
                            "LDS chase", "collatz branches", "REDUX.OR+IADD", "model step", "SHFL.IDX lane0 + IADD" };
    for (int rep = 0; rep < 2; rep++) {
        chain<0><<<1, 32>>>(out, cyc, 5, 7, 9);
        chain<1><<<1, 32>>>(out, cyc, 5, 7, 9);
        chain<2><<<1, 32>>>(out, cyc, 5, 7, 9);
        chain<3><<<1, 32>>>(out, cyc, 5, 7, 900);
        chain<4><<<1, 32>>>(out, cyc, 5, 7, 900000);
        chain<5><<<1, 32>>>(out, cyc, 5, 0x5a5a5a5a, 9);
        chain<6><<<1, 32>>>(out, cyc, 5, 7, 9);
        chain<7><<<1, 32>>>(out, cyc, 5, 7, 9);
        chain<8><<<1, 32>>>(out, cyc, 27, 7, 27);
        chain<9><<<1, 32>>>(out, cyc, 5, 7, 9);
        chain<10><<<1, 32>>>(out, cyc, 5, 7, 9);
        chain<11><<<1, 32>>>(out, cyc, 5, 7, 9);
        cudaDeviceSynchronize();
    }
    for (int k = 0; k < 12; k++)
        printf("%-28s %.2f cycles/link\n", names[k], (double)cyc[k] / N);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
