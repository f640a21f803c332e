// mma_rate.cu — issue rate and latency of mma.sync.m16n8k16 (bf16, fp32 accumulate) on one SM of a B200, per warp count.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int ILP>
__global__ void k(long long* out, int iters) {
    uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x3c003c00u;
    float d[ILP][4];
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma(d[i], a, b0, b1);
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    if (s == 123.456f) out[1] = 1;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    const int iters = 2000;
    for (int warps : {1, 4, 8, 16}) {
        long long h;
        k<1><<<1, warps * 32>>>(d, iters); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d  dependent chain (ILP 1): %.1f cycles per MMA per warp\n", warps, (double)h / iters);
        k<2><<<1, warps * 32>>>(d, iters); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d  ILP 2: %.2f cycles per MMA per warp\n", warps, (double)h / iters / 2);
        k<6><<<1, warps * 32>>>(d, iters); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d  ILP 6: %.2f cycles per MMA per warp\n", warps, (double)h / iters / 6);
        k<12><<<1, warps * 32>>>(d, iters); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("warps/SM %2d  ILP 12: %.2f cycles per MMA per warp  (= %.2f per MMA per SM sub-partition)\n", warps, (double)h / iters / 12,
               (double)h / iters / 12 / (warps < 4 ? 1 : warps / 4));
    }
    return 0;
}
