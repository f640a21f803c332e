// scatter_bw.cu — what the memory system of a B200 gives to the access pattern of the power-law sweep (config 5):
// 32-byte messages streamed in, streamed out, and scattered / gathered at random 32-byte granularity.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_bw scatter_bw.cu ; run: ./scatter_bw [n_slots]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// two lanes per 32-byte message (16 bytes each), like k_pw_*
template <int MODE>  // 0: stream -> stream, 1: stream -> scatter, 2: gather -> stream, 3: sweep mix (read 1, write 1 stream + 1 scatter)
__global__ void k(const uint4* __restrict__ src, uint4* __restrict__ dst, uint4* __restrict__ dst2, const uint32_t* __restrict__ perm, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x / 2;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 2; i < n; i += stride) {
        int h = threadIdx.x & 1;
        uint32_t p = perm[i];
        if (MODE == 0) { uint4 v = __ldcs(src + 2 * i + h); __stcs(dst + 2 * i + h, v); }
        if (MODE == 1) { uint4 v = __ldcs(src + 2 * i + h); dst[2 * (size_t)p + h] = v; }
        if (MODE == 2) { uint4 v = src[2 * (size_t)p + h]; __stcs(dst + 2 * i + h, v); }
        if (MODE == 3) { uint4 v = __ldcs(src + 2 * i + h); __stcs(dst2 + 2 * i + h, v); v.x ^= 1; dst[2 * (size_t)p + h] = v; }
    }
}
// scatter only: nothing is read but the index
__global__ void k_scatter_only(uint4* __restrict__ dst, const uint32_t* __restrict__ perm, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x / 2;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 2; i < n; i += stride) {
        int h = threadIdx.x & 1;
        uint32_t p = perm[i];
        dst[2 * (size_t)p + h] = make_uint4(p, h, 0, 0);
    }
}
int main(int argc, char** argv) {
    size_t n = argc > 1 ? atoll(argv[1]) : 40000000;
    std::vector<uint32_t> perm(n);
    for (size_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
    std::mt19937_64 rng(1);
    std::shuffle(perm.begin(), perm.end(), rng);
    // a second permutation with locality: random inside windows of W slots (W * 32 B = the span that stays in L2)
    uint4 *src, *dst, *dst2; uint32_t *dperm, *dperm_w;
    CK(cudaMalloc(&src, n * 32)); CK(cudaMalloc(&dst, n * 32)); CK(cudaMalloc(&dst2, n * 32));
    CK(cudaMalloc(&dperm, n * 4)); CK(cudaMalloc(&dperm_w, n * 4));
    CK(cudaMemset(src, 1, n * 32)); CK(cudaMemset(dst, 0, n * 32)); CK(cudaMemset(dst2, 0, n * 32));
    CK(cudaMemcpy(dperm, perm.data(), n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch, double bytes) {
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int R = 10;
        for (int i = 0; i < R; ++i) launch();
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= R;
        printf("%-44s %8.3f ms  %8.1f GB/s  (%.1f G msgs/s)\n", name, ms, bytes / ms * 1e-6, n / ms * 1e-6);
    };
    for (int blocks_per_sm : {4, 8}) {
        int grid = 148 * blocks_per_sm;
        printf("grid %d x 256, n = %zu slots of 32 B\n", grid, n);
        run("stream -> stream (64 B per slot)", [&] { k<0><<<grid, 256>>>(src, dst, dst2, dperm, n); }, n * 68.0);
        run("stream -> random 32 B scatter", [&] { k<1><<<grid, 256>>>(src, dst, dst2, dperm, n); }, n * 68.0);
        run("random 32 B gather -> stream", [&] { k<2><<<grid, 256>>>(src, dst, dst2, dperm, n); }, n * 68.0);
        run("scatter only (index stream + 32 B scatter)", [&] { k_scatter_only<<<grid, 256>>>(dst, dperm, n); }, n * 36.0);
        run("sweep mix: read 32, write 32, scatter 32", [&] { k<3><<<grid, 256>>>(src, dst, dst2, dperm, n); }, n * 100.0);
    }
    // windowed permutations: destinations random within windows of W slots
    for (size_t W : {(size_t)1 << 16, (size_t)1 << 20, (size_t)1 << 21, (size_t)1 << 22}) {
        for (size_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
        for (size_t w0 = 0; w0 < n; w0 += W) std::shuffle(perm.begin() + w0, perm.begin() + std::min(n, w0 + W), rng);
        CK(cudaMemcpy(dperm_w, perm.data(), n * 4, cudaMemcpyHostToDevice));
        char nm[96]; snprintf(nm, sizeof nm, "scatter within windows of %zu MB", W * 32 >> 20);
        run(nm, [&] { k<1><<<148 * 8, 256>>>(src, dst, dst2, dperm_w, n); }, n * 68.0);
    }
    return 0;
}
