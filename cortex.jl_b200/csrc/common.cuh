// common.cuh — shared helpers of the sm_100a kernels (error handling, device buffers, launch counting).
#pragma once
#include <cstdio>
#include <exception>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cortex_b200.h"

namespace cxb {

// every kernel launch of the library goes through this counter (bench.py reports it as gpu_launches)
extern unsigned long long g_kernel_launches;
#define CXB_LAUNCH(kernel, grid, block, smem, stream, ...)      \
    do {                                                        \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
        ++::cxb::g_kernel_launches;                             \
    } while (0)

inline std::string cuda_msg(cudaError_t e, const char* what) {
    return std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
}

#define CXB_CUDA(expr)                                  \
    do {                                                \
        cudaError_t e__ = (expr);                       \
        if (e__ != cudaSuccess) {                       \
            this->err = ::cxb::cuda_msg(e__, #expr);    \
            return CXB_ERR_CUDA;                        \
        }                                               \
    } while (0)

// growable device buffer
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) cap = n ? n : 1;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    ~DBuf() { release(); }
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
};

// pinned host buffer
template <class T>
struct HBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost((void**)&p, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) cap = n ? n : 1;
        return e;
    }
    ~HBuf() {
        if (p) cudaFreeHost(p);
    }
    HBuf() = default;
    HBuf(const HBuf&) = delete;
    HBuf& operator=(const HBuf&) = delete;
};

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace cxb

// No C++ exception may cross the C ABI (the callers are Julia ccall / Python ctypes / C): every entry point that does host
// allocation is a function-try-block closed by this handler. The engine's last-error text is left as it was.
#define CXB_ABI_CATCH(ret)                                                             \
    catch (const std::exception& ex) {                                                 \
        fprintf(stderr, "cortex_b200: C++ exception at the ABI boundary: %s\n", ex.what()); \
        return ret;                                                                    \
    }                                                                                  \
    catch (...) {                                                                      \
        fprintf(stderr, "cortex_b200: unknown C++ exception at the ABI boundary\n");   \
        return ret;                                                                    \
    }

namespace cxb {

}  // namespace cxb
