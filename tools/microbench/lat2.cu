// lat2.cu -- shuffle throughput / overlap, taken-branch cost and candidate SBRT step bodies
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32;
#define N 2048
#define FULL 0xffffffffu

template <int WHICH>
__global__ void k(u32* out, long long* cyc, u32 seed, u32 c1, const u32* __restrict__ ranks)
{
    const int lane = threadIdx.x;
    u32 x = seed + lane, y = seed * 3 + lane, z = lane * 7 + 1;
    u32 K = 0, P = lane; // list state for the step candidates
    long long t0 = clock64();
    if (WHICH == 0) { // 8 independent SHFL.UP then fold (throughput)
#pragma unroll 4
        for (int i = 0; i < N; i++) {
            u32 a0 = __shfl_up_sync(FULL, x, 1), a1 = __shfl_up_sync(FULL, x + 1, 1), a2 = __shfl_up_sync(FULL, x + 2, 1), a3 = __shfl_up_sync(FULL, x + 3, 1);
            u32 a4 = __shfl_up_sync(FULL, x + 4, 1), a5 = __shfl_up_sync(FULL, x + 5, 1), a6 = __shfl_up_sync(FULL, x + 6, 1), a7 = __shfl_up_sync(FULL, x + 7, 1);
            x = (a0 ^ a1) + (a2 ^ a3) + (a4 ^ a5) + (a6 ^ a7);
        }
    }
    if (WHICH == 1) { // IDX + 2 UP in parallel
#pragma unroll 4
        for (int i = 0; i < N; i++) {
            u32 a0 = __shfl_sync(FULL, x, x & 31), a1 = __shfl_up_sync(FULL, x + 1, 1), a2 = __shfl_up_sync(FULL, x + 2, 1);
            x = a0 + a1 + a2;
        }
    }
    if (WHICH == 2) { // 2 IDX in parallel
#pragma unroll 4
        for (int i = 0; i < N; i++) {
            u32 a0 = __shfl_sync(FULL, x, x & 31), a1 = __shfl_sync(FULL, x + 1, (x >> 5) & 31);
            x = a0 + a1;
        }
    }
    if (WHICH == 3) { // taken branches: 4-way dispatch on data, bodies too different to predicate
        for (int i = 0; i < N; i++) {
            switch (x & 3) {
            case 0: x = x * 3 + 1; asm volatile(""); break;
            case 1: x = (x >> 2) ^ c1; asm volatile(""); x += 5; break;
            case 2: x = __popc(x) + c1; asm volatile(""); x ^= 77; break;
            default: x = x + (x << 3); asm volatile(""); x -= 9; break;
            }
        }
    }
    if (WHICH == 4 || WHICH == 5) { // step candidates, 4 steps per word straight line
        for (int i = 0; i < N; i += 4) {
            const u32 w4 = ranks[(i >> 2) & 255];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int r = (w4 >> (8 * s)) & 31;
                const u32 ii = i + s;
                if (WHICH == 4) { // old formulation: q = (i+p)>>1, three predicate hops
                    const int nq = __shfl_up_sync(FULL, (int)K, 1);
                    const u32 npb = __shfl_up_sync(FULL, P, 1);
                    const u32 e = __shfl_sync(FULL, P, r);
                    const u32 c = e & 0xFF;
                    const int qc = (int)((ii + (e >> 8)) >> 1);
                    const bool below = lane <= r;
                    const bool mv = below && lane != 0 && (nq <= qc);
                    const bool ins = below && ((int)K <= qc) && (lane == 0 || (nq > qc));
                    const u32 ne = (ii << 8) | c;
                    K = ins ? qc : (mv ? nq : K);
                    P = ins ? ne : (mv ? npb : P);
                    y += c;
                } else { // even keys, one predicate hop
                    const u32 e = __shfl_sync(FULL, P, r);
                    u32 nK = __shfl_up_sync(FULL, K, 1);
                    const u32 nP = __shfl_up_sync(FULL, P, 1);
                    if (lane == 0) nK = 0xFFFFFFFFu;
                    const u32 c = e & 0xFF;
                    const u32 Y = ii + (e >> 8);
                    const u32 Yn = Y & ~1u, ne = (ii << 8) | c;
                    if (lane <= r && K <= Y) {
                        const bool up = nK <= Y;
                        K = up ? nK : Yn;
                        P = up ? nP : ne;
                    }
                    y += c;
                }
            }
        }
        x = K + P;
    }
    if (WHICH == 6) { // 8 ballots + popc + add (position search of the generic path)
#pragma unroll 2
        for (int i = 0; i < N; i++) {
            int rp = 0;
#pragma unroll
            for (int s = 0; s < 8; s++)
                rp += __popc(__ballot_sync(FULL, (x + s * 31 + lane) > y));
            x = x + rp + 1;
        }
    }
    if (WHICH == 7) { // 64-bit packed entry: 2 SHFL.IDX
#pragma unroll 4
        for (int i = 0; i < N; i++) {
            unsigned long long v = ((unsigned long long)x << 32) | z;
            v = __shfl_sync(FULL, v, x & 31);
            x = (u32)(v >> 32) + (u32)v;
        }
    }
    long long t1 = clock64();
    out[lane] = x + y + z;
    if (lane == 0)
        cyc[WHICH] = t1 - t0;
}

int main()
{
    u32 *out, *ranks;
    long long* cyc;
    cudaMalloc(&out, 128);
    cudaMallocManaged(&ranks, 1024);
    cudaMallocManaged(&cyc, 16 * 8);
    for (int i = 0; i < 256; i++)
        ranks[i] = (u32)((i * 2654435761u) >> 3) & 0x1f1f1f1fu;
    const char* names[] = { "8 indep SHFL.UP + fold (per iter)", "IDX + 2 UP parallel", "2 IDX parallel", "4-way taken branch", "old step (per step)",
                            "one-hop step (per step)", "8 ballots+popc", "64-bit SHFL.IDX" };
    for (int rep = 0; rep < 2; rep++) {
        k<0><<<1, 32>>>(out, cyc, 5, 7, ranks);
        k<1><<<1, 32>>>(out, cyc, 5, 7, ranks);
        k<2><<<1, 32>>>(out, cyc, 5, 7, ranks);
        k<3><<<1, 32>>>(out, cyc, 5, 7, ranks);
        k<4><<<1, 32>>>(out, cyc, 5, 7, ranks);
        k<5><<<1, 32>>>(out, cyc, 5, 7, ranks);
        k<6><<<1, 32>>>(out, cyc, 5, 7, ranks);
        k<7><<<1, 32>>>(out, cyc, 5, 7, ranks);
        cudaDeviceSynchronize();
    }
    for (int i = 0; i < 8; i++)
        printf("%-36s %.2f cycles\n", names[i], (double)cyc[i] / N);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
