#include "cusim.h"
__global__ void k(int* out, int n) {
    __shared__ int s[64];
    int t = threadIdx.x;
    s[t] = t; __syncthreads();
    int v = s[(t+1)%64];
    unsigned b = __ballot_sync(0xFFFFFFFFu, v & 1);
    int sh = __shfl_xor_sync(0xFFFFFFFFu, v, 1);
    int sum = v; for (int o=16;o>0;o>>=1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    if (t < n) out[blockIdx.x*64+t] = sum + (b==0x55555555u || b==0xAAAAAAAAu) + sh*0;
}
int main(){ int out[128]; KLAUNCH(k, dim3(2), dim3(64), 0, out, 64); printf("%d %d %d\n", out[0], out[33], out[127]); return 0; }
