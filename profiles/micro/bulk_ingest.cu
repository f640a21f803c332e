// bulk_ingest.cu — how fast can ONE SM pull L2-resident data into shared memory with cp.async.bulk (the operand ingest of
// the K = 512 HMM step kernel: 393 KB message + 98 KB table slice per CTA and step through a ~200 KB ring)?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o bulk_ingest bulk_ingest.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t n, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n), "r"(bar) : "memory");
}
// every CTA streams `total` bytes of the region [src + group * region, + region) (wrapping) through a ring of `stages` x `chunk` bytes;
// each stage is filled by `split` bulk copies
__global__ void k(const unsigned char* src, size_t region, int ctas_per_group, int total_chunks, int chunk, int stages, int split, long long* cyc,
                  int first_copy, int delay) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)stages * chunk);
    uint64_t* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(s32(&full[s]), 1); mbar_init(s32(&empty[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned char* base = src + (size_t)(blockIdx.x / ctas_per_group) * region;
    long long t0 = clock64();
    if (threadIdx.x == 0) {  // producer
        size_t off = 0;
        for (int c = 0; c < total_chunks; ++c) {
            int s = c % stages;
            mbar_wait(s32(&empty[s]), ((c / stages) & 1) ^ 1);
            mbar_expect(s32(&full[s]), chunk);
            if (first_copy > 0) {  // one copy of first_copy bytes, the rest in (split - 1) equal copies
                bulk(s32(sm + (size_t)s * chunk), base + off, first_copy, s32(&full[s]));
                const int rest = (chunk - first_copy) / (split - 1);
                for (int q = 0; q < split - 1; ++q)
                    bulk(s32(sm + (size_t)s * chunk + first_copy + (size_t)q * rest), base + off + first_copy + (size_t)q * rest, rest, s32(&full[s]));
                off += chunk;
                if (off + chunk > region) off = 0;
            } else
            for (int q = 0; q < split; ++q) {
                bulk(s32(sm + (size_t)s * chunk + (size_t)q * (chunk / split)), base + off, chunk / split, s32(&full[s]));
                off += chunk / split;
                if (off + chunk / split > region) off = 0;
            }
        }
    } else if (threadIdx.x == 32) {  // consumer: frees the stage as soon as it has landed
        for (int c = 0; c < total_chunks; ++c) {
            int s = c % stages;
            mbar_wait(s32(&full[s]), (c / stages) & 1);
            if (delay) { long long t = clock64(); while (clock64() - t < delay) { } }
            mbar_arrive(s32(&empty[s]));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}
int main() {
    const size_t region = 491520;  // 480 KB: message + table slice of one CTA
    const int max_groups = 148;
    unsigned char* src; long long* cyc;
    CK(cudaMalloc(&src, region * max_groups)); CK(cudaMemset(src, 1, region * max_groups));
    CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    struct Cfg { int grid, cpg, chunk, stages, split, first, delay; };
    Cfg cfgs[] = {{64, 16, 61440, 3, 6, 0, 0}, {64, 16, 61440, 3, 1, 0, 0}, {64, 16, 61440, 3, 2, 49152, 0}, {64, 16, 61440, 3, 2, 49152, 400},
                  {64, 16, 61440, 3, 1, 0, 400}, {64, 16, 61440, 3, 1, 0, 1000}, {64, 16, 49152, 4, 1, 0, 0}, {64, 16, 98304, 2, 1, 0, 0}, {64, 16, 98304, 2, 1, 0, 400},
                  {64, 16, 16384, 12, 1, 0, 0}, {64, 16, 32768, 6, 1, 0, 0}, {64, 16, 122880, 1, 1, 0, 0}, {64, 16, 196608, 1, 1, 0, 0}};
    for (auto c : cfgs) {
        int total_chunks = (int)(region * 40 / c.chunk);  // 40 steps' worth
        size_t smem = (size_t)c.stages * c.chunk + 2 * c.stages * 8 + 64;
        for (int rep = 0; rep < 2; ++rep) k<<<c.grid, 64, smem>>>(src, region, c.cpg, total_chunks, c.chunk, c.stages, c.split, cyc, c.first, c.delay);
        CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, cyc, c.grid * 8, cudaMemcpyDeviceToHost));
        double mx = 0, av = 0; for (int i = 0; i < c.grid; ++i) { av += h[i]; if (h[i] > mx) mx = h[i]; } av /= c.grid;
        double bytes = (double)total_chunks * c.chunk;
        printf("grid %3d, ring %2d x %6d B, %d copies per stage (first %5d B), consumer holds a stage %4d cycles: %6.1f B/clk/SM, %5.0f cycles per stage, %.2f us per 480 KB\n",
               c.grid, c.stages, c.chunk, c.split, c.first, c.delay, bytes / av, av / total_chunks, av / 40 / 1.9e3);
    }
    return 0;
}
