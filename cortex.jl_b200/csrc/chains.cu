// chains.cu — structured engine for a batch of independent linear-Gaussian random-walk chains
// (BASELINE configs 1-2).  One call = update_marginals!(engine, x[1:T]) of every chain of the graph
// of test/inference_engine_tests.jl:436-462 under the DefaultDependencyResolver: the forward round
// computes m2v(x_t,lik_t), m2v(x_t,tr_{t-1}), m2f(x_t,tr_t); the reverse round m2v(x_t,tr_t),
// m2f(x_t,tr_{t-1}); the final phase the marginals — 6T-4 message updates per chain,
// every one of them materialised in HBM ("materialise-all").
//
// Rules (canonical form (precision L, precision-mean h)):
//   observation factor, variance r :  (1/r, y/r)
//   random-walk factor, variance q :  (L, h) -> (L/(1+qL), h/(1+qL))
//   m2f / marginal                 :  component-wise sum of the dependencies, left to right
//
// Data layout: time-major [T][B] so that the 32 chains of a warp touch one contiguous 128 B (y) /
// 256 B (float2 message) / 512 B (double2) segment per step. One thread owns one chain; the loads of
// a whole time tile are issued before the sequential recursion consumes them (the recursion itself
// is a ~40-cycle dependent chain per step, hidden by ~14 resident warps per SM). Stores are
// streaming (st.global.cs): nothing written is re-read before ~2 GB of other traffic.
// HBM-bound: 64 B (fp32) / 128 B (fp64) of algorithmic traffic per variable.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace cxb {

template <class T>
struct Vec2;
template <>
struct Vec2<float> {
    using type = float2;
};
template <>
struct Vec2<double> {
    using type = double2;
};

template <class T>
__device__ __forceinline__ typename Vec2<T>::type mk2(T a, T b) {
    typename Vec2<T>::type v;
    v.x = a;
    v.y = b;
    return v;
}


// (Replacing the two IEEE divisions per step by MUFU.RCP + Newton was measured SLOWER here — 0.749 ms vs 0.728 ms at the
// best tile / block of a fresh sweep — although the same change gained 6.5 % in the Potts kernel; the divisions stay.)
// msg: [6][T][B] of (L,h) pairs. Classes: 0 m2v(x_t,lik_t)  1 m2v(x_t,tr_{t-1})  2 m2f(x_t,tr_t)
//                                          3 m2v(x_t,tr_t)   4 m2f(x_t,tr_{t-1})  5 marginal(x_t)
template <class T, int TILE, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 512 / BLOCK)
k_chains_fwd_bwd(const T* __restrict__ y, const T* __restrict__ qv, const T* __restrict__ rv,
                 typename Vec2<T>::type* __restrict__ msg, long long B, long long Tn, long long b0, long long b1) {
    using V = typename Vec2<T>::type;
    const long long b = b0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;  // chains [b0, b1) of the batch
    if (b >= b1) return;
    const T q = qv[b], r = rv[b];
    const T inv_r = T(1) / r;
    const size_t plane = (size_t)Tn * (size_t)B;
    V* __restrict__ m_obs = msg;
    V* __restrict__ m_pred = msg + plane;
    V* __restrict__ m_fwd = msg + 2 * plane;
    V* __restrict__ m_bwd = msg + 3 * plane;
    V* __restrict__ m_back = msg + 4 * plane;
    V* __restrict__ m_marg = msg + 5 * plane;

    // ---- forward round: t = 0 .. T-1 -------------------------------------------------------------------
    // software pipeline: the loads of tile n+1 are issued before the recursion consumes tile n.
    // Addresses advance by running offsets (one 64-bit add per step) to keep the register count at 128.
    const size_t sB = (size_t)B;
    T L = 0, h = 0;  // m2f(x_{t-1}, tr_{t-1}) carried in registers
    T ynext[TILE];
    size_t lidx = (size_t)b;  // load cursor
#pragma unroll
    for (int k = 0; k < TILE; ++k) {
        ynext[k] = k < Tn ? __ldcs(y + lidx) : T(0);
        lidx += sB;
    }
    size_t sidx = (size_t)b;  // store cursor
    for (long long t0 = 0; t0 < Tn; t0 += TILE) {
        T yy[TILE];
#pragma unroll
        for (int k = 0; k < TILE; ++k) yy[k] = ynext[k];
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
            ynext[k] = (t0 + TILE + k) < Tn ? __ldcs(y + lidx) : T(0);
            lidx += sB;
        }
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
            long long t = t0 + k;
            if (t < Tn) {
                T oL = inv_r, oh = yy[k] * inv_r;  // m2v(x_t, lik_t) = (1/r, y/r)
                T pL = 0, ph = 0;
                if (t > 0) {  // m2v(x_t, tr_{t-1}) = RW(m2f(x_{t-1}, tr_{t-1}))
                    T den = T(1) + q * L;
                    pL = L / den;
                    ph = h / den;
                    L = oL + pL;  // m2f(x_t, tr_t) = lik (+) tr_{t-1}, dependency order
                    h = oh + ph;
                } else {
                    L = oL;
                    h = oh;
                }
                __stcs(m_obs + sidx, mk2<T>(oL, oh));
                __stcs(m_pred + sidx, mk2<T>(pL, ph));
                __stcs(m_fwd + sidx, mk2<T>(L, h));
                sidx += sB;
            }
        }
    }
    // ---- reverse round + final phase: t = T-1 .. 0 --------------------------------------------------------
    L = 0;
    h = 0;  // m2f(x_{t+1}, tr_t)
    V pnext[TILE];
    lidx = (size_t)(Tn - 1) * sB + (size_t)b;
#pragma unroll
    for (int k = 0; k < TILE; ++k) {
        bool ok = Tn - 1 - k >= 0;
        ynext[k] = ok ? __ldcs(y + lidx) : T(0);
        pnext[k] = ok ? __ldcs(m_pred + lidx) : mk2<T>(T(0), T(0));
        lidx -= sB;
    }
    sidx = (size_t)(Tn - 1) * sB + (size_t)b;
    for (long long t1 = Tn - 1; t1 >= 0; t1 -= TILE) {
        T yy[TILE];
        V pr[TILE];
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
            yy[k] = ynext[k];
            pr[k] = pnext[k];
        }
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
            bool ok = t1 - TILE - k >= 0;
            ynext[k] = ok ? __ldcs(y + lidx) : T(0);
            pnext[k] = ok ? __ldcs(m_pred + lidx) : mk2<T>(T(0), T(0));
            lidx -= sB;
        }
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
            long long t = t1 - k;
            if (t >= 0) {
                T oL = inv_r, oh = yy[k] * inv_r;
                T bL = 0, bh = 0;
                if (t < Tn - 1) {  // m2v(x_t, tr_t) = RW(m2f(x_{t+1}, tr_t))
                    T den = T(1) + q * L;
                    bL = L / den;
                    bh = h / den;
                    L = oL + bL;  // m2f(x_t, tr_{t-1}) = lik (+) tr_t
                    h = oh + bh;
                } else {
                    L = oL;
                    h = oh;
                }
                // marginal(x_t) = lik (+) tr_{t-1} (+) tr_t, left to right
                T gL = oL, gh = oh;
                if (t > 0) {
                    gL = gL + pr[k].x;
                    gh = gh + pr[k].y;
                }
                if (t < Tn - 1) {
                    gL = gL + bL;
                    gh = gh + bh;
                }
                __stcs(m_bwd + sidx, mk2<T>(bL, bh));
                __stcs(m_back + sidx, mk2<T>(L, h));
                __stcs(m_marg + sidx, mk2<T>(gL, gh));
                sidx -= sB;
            }
        }
    }
}

// Variant without the explicit double buffer: the loads of a time tile are issued together at the top of the tile and
// the recursion then consumes them (latency is hidden by the other resident warps only). Kept because it measured
// FASTER for fp32 on B200 (0.739 ms vs 0.79 ms, profiles/r01_sweep_chains.txt): CXB_CHAINS_PIPE=0 selects it.
template <class T, int TILE, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_chains_fwd_bwd_np(const T* __restrict__ y, const T* __restrict__ qv, const T* __restrict__ rv,
                    typename Vec2<T>::type* __restrict__ msg, long long B, long long Tn, long long b0, long long b1) {
    using V = typename Vec2<T>::type;
    const long long b = b0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;  // chains [b0, b1) of the batch
    if (b >= b1) return;
    const T q = qv[b], r = rv[b];
    const T inv_r = T(1) / r;
    const size_t plane = (size_t)Tn * (size_t)B;
    V* __restrict__ m_obs = msg;
    V* __restrict__ m_pred = msg + plane;
    V* __restrict__ m_fwd = msg + 2 * plane;
    V* __restrict__ m_bwd = msg + 3 * plane;
    V* __restrict__ m_back = msg + 4 * plane;
    V* __restrict__ m_marg = msg + 5 * plane;
    T L = 0, h = 0;
    for (long long t0 = 0; t0 < Tn; t0 += TILE) {
        T yy[TILE];
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
            long long t = t0 + k;
            yy[k] = t < Tn ? __ldcs(&y[(size_t)t * B + b]) : T(0);
        }
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
            long long t = t0 + k;
            if (t < Tn) {
                size_t idx = (size_t)t * B + b;
                T oL = inv_r, oh = yy[k] * inv_r;
                T pL = 0, ph = 0;
                if (t > 0) {
                    T den = T(1) + q * L;
                    pL = L / den;
                    ph = h / den;
                    L = oL + pL;
                    h = oh + ph;
                } else {
                    L = oL;
                    h = oh;
                }
                __stcs(&m_obs[idx], mk2<T>(oL, oh));
                __stcs(&m_pred[idx], mk2<T>(pL, ph));
                __stcs(&m_fwd[idx], mk2<T>(L, h));
            }
        }
    }
    L = 0;
    h = 0;
    for (long long t1 = Tn - 1; t1 >= 0; t1 -= TILE) {
        T yy[TILE];
        V pr[TILE];
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
            long long t = t1 - k;
            if (t >= 0) {
                yy[k] = __ldcs(&y[(size_t)t * B + b]);
                pr[k] = __ldcs(&m_pred[(size_t)t * B + b]);
            } else {
                yy[k] = T(0);
                pr[k] = mk2<T>(T(0), T(0));
            }
        }
#pragma unroll
        for (int k = 0; k < TILE; ++k) {
            long long t = t1 - k;
            if (t >= 0) {
                size_t idx = (size_t)t * B + b;
                T oL = inv_r, oh = yy[k] * inv_r;
                T bL = 0, bh = 0;
                if (t < Tn - 1) {
                    T den = T(1) + q * L;
                    bL = L / den;
                    bh = h / den;
                    L = oL + bL;
                    h = oh + bh;
                } else {
                    L = oL;
                    h = oh;
                }
                T gL = oL, gh = oh;
                if (t > 0) {
                    gL = gL + pr[k].x;
                    gh = gh + pr[k].y;
                }
                if (t < Tn - 1) {
                    gL = gL + bL;
                    gh = gh + bh;
                }
                __stcs(&m_bwd[idx], mk2<T>(bL, bh));
                __stcs(&m_back[idx], mk2<T>(L, h));
                __stcs(&m_marg[idx], mk2<T>(gL, gh));
            }
        }
    }
}

struct Chains {
    int device = 0, dtype = CXB_F32;
    long long B = 0, T = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the chunk-pipelined host entry point
    static constexpr int MAX_CHUNKS = 32;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_k[MAX_CHUNKS] = {}, ev_done = nullptr;
    long long rb0 = 0, rb1 = 0;  // chain range of the next launch
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    DBuf<unsigned char> y, q, r, msg;
    bool have_noise = false, have_obs = false, ran = false;
    size_t esz() const { return dtype == CXB_F32 ? 4 : 8; }
    size_t plane_bytes() const { return (size_t)T * B * 2 * esz(); }
    ~Chains() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (int i = 0; i < MAX_CHUNKS; ++i) {
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_k[i]) cudaEventDestroy(ev_k[i]);
        }
        if (ev_done) cudaEventDestroy(ev_done);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        if (stream) cudaStreamDestroy(stream);
    }
    int32_t init() {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            err = "no CUDA device available (cortex_b200 has no CPU fallback)";
            return CXB_ERR_CUDA;
        }
        if (device < 0 || device >= count || B <= 0 || T <= 0) {
            err = "bad device / shape";
            return CXB_ERR_BAD_ARG;
        }
        CXB_CUDA(cudaSetDevice(device));
        CXB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CXB_CUDA(cudaEventCreate(&ev0));
        CXB_CUDA(cudaEventCreate(&ev1));
        CXB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        CXB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < MAX_CHUNKS; ++i) {
            CXB_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            CXB_CUDA(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
        }
        CXB_CUDA(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
        rb0 = 0;
        rb1 = B;
        CXB_CUDA(y.reserve((size_t)T * B * esz()));
        CXB_CUDA(q.reserve((size_t)B * esz()));
        CXB_CUDA(r.reserve((size_t)B * esz()));
        CXB_CUDA(msg.reserve(6 * plane_bytes()));
        return CXB_OK;
    }
    int32_t set_noise(const double* qh, const double* rh) {
        CXB_CUDA(cudaSetDevice(device));
        std::vector<unsigned char> a((size_t)B * esz()), c((size_t)B * esz());
        for (long long i = 0; i < B; ++i) {
            if (!(qh[i] >= 0) || !(rh[i] > 0)) {
                err = "noise variances must satisfy q >= 0, r > 0";
                return CXB_ERR_BAD_ARG;
            }
            if (dtype == CXB_F32) {
                ((float*)a.data())[i] = (float)qh[i];
                ((float*)c.data())[i] = (float)rh[i];
            } else {
                ((double*)a.data())[i] = qh[i];
                ((double*)c.data())[i] = rh[i];
            }
        }
        CXB_CUDA(cudaMemcpyAsync(q.p, a.data(), a.size(), cudaMemcpyHostToDevice, stream));
        CXB_CUDA(cudaMemcpyAsync(r.p, c.data(), c.size(), cudaMemcpyHostToDevice, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        have_noise = true;
        return CXB_OK;
    }
    template <class T, int TILE, int BLOCK>
    int32_t launch_variant() {
        CXB_LAUNCH((k_chains_fwd_bwd<T, TILE, BLOCK>), cdiv((size_t)(rb1 - rb0), BLOCK), BLOCK, 0, stream, (const T*)y.p,
                   (const T*)q.p, (const T*)r.p, (typename Vec2<T>::type*)msg.p, B, this->T, rb0, rb1);
        return CXB_OK;
    }
    // tile = time steps in flight per thread (double buffered), block = chains per CTA. Defaults are the measured
    // best on B200 (profiles/); CXB_CHAINS_TILE / CXB_CHAINS_BLOCK override them for tuning runs.
    template <class T, int TILE, int BLOCK>
    int32_t launch_variant_np() {
        CXB_LAUNCH((k_chains_fwd_bwd_np<T, TILE, BLOCK>), cdiv((size_t)(rb1 - rb0), BLOCK), BLOCK, 0, stream, (const T*)y.p,
                   (const T*)q.p, (const T*)r.p, (typename Vec2<T>::type*)msg.p, B, this->T, rb0, rb1);
        return CXB_OK;
    }
    template <class T>
    int32_t dispatch() {
        // measured best on B200 (profiles/r01_sweep_chains*.txt): 64-thread CTAs (1,024 CTAs = 6.9 per SM), tile loads
        // issued together at the top of the tile (no explicit double buffer): fp32 tile 8 -> 0.725 ms (0.904 of the
        // measured copy peak), fp64 tile 4 -> 1.474 ms (0.889)
        int tile = sizeof(T) == 4 ? 8 : 4, block = 64, pipe = 0;
        if (const char* e = getenv("CXB_CHAINS_TILE")) tile = atoi(e);
        if (const char* e = getenv("CXB_CHAINS_BLOCK")) block = atoi(e);
        if (const char* e = getenv("CXB_CHAINS_PIPE")) pipe = atoi(e);
#define V_(TL, BL) if (pipe && tile == TL && block == BL) return launch_variant<T, TL, BL>();
        V_(2, 64) V_(4, 64) V_(8, 64) V_(16, 64) V_(2, 128) V_(4, 128) V_(8, 128) V_(16, 128) V_(4, 256) V_(8, 256) V_(4, 32) V_(8, 32)
#undef V_
#define V_(TL, BL) if (!pipe && tile == TL && block == BL) return launch_variant_np<T, TL, BL>();
        V_(4, 64) V_(8, 64) V_(16, 64) V_(4, 128) V_(8, 128) V_(16, 128) V_(8, 256) V_(16, 256)
#undef V_
        err = "unsupported CXB_CHAINS_TILE / CXB_CHAINS_BLOCK combination";
        return CXB_ERR_BAD_ARG;
    }
    int32_t launch() {
        if (!have_noise || !have_obs) {
            err = "set the noise variances and the observations first";
            return CXB_ERR_STATE;
        }
        CXB_CUDA(cudaSetDevice(device));
        rb0 = 0;
        rb1 = B;
        CXB_CUDA(cudaEventRecord(ev0, stream));
        int32_t st = dtype == CXB_F32 ? dispatch<float>() : dispatch<double>();
        if (st) return st;
        CXB_CUDA(cudaEventRecord(ev1, stream));
        CXB_CUDA(cudaGetLastError());
        ran = true;
        return CXB_OK;
    }
    // Host entry point: observations from (pinned) host memory, marginals back to host memory. The batch is cut into
    // chunks of chains; chunk c's observations travel host->device on s_in while chunk c-1 is computed on `stream` and
    // chunk c-2's marginals travel device->host on s_out, so that the PCIe link is busy in both directions at once.
    int32_t infer_host(const void* y_host, void* marg_out) {
        if (!have_noise) {
            err = "set the noise variances first";
            return CXB_ERR_STATE;
        }
        CXB_CUDA(cudaSetDevice(device));
        int n_chunks = 16;  // measured: 4..32 chunks are within 5 % of each other (the PCIe link is the bound); 16 was best
        if (const char* e = getenv("CXB_CHAINS_HOST_CHUNKS")) n_chunks = atoi(e);
        n_chunks = std::max(1, std::min<int>(n_chunks, MAX_CHUNKS));
        long long per = ((B + n_chunks - 1) / n_chunks + 63) / 64 * 64;  // whole CTAs per chunk
        n_chunks = (int)((B + per - 1) / per);
        const size_t e1 = esz(), e2 = 2 * esz();
        unsigned char* marg_dev = msg.p + 5 * plane_bytes();
        // the previous call may still be using the buffers on the other streams
        CXB_CUDA(cudaEventRecord(ev_done, stream));
        CXB_CUDA(cudaStreamWaitEvent(s_in, ev_done, 0));
        CXB_CUDA(cudaEventRecord(ev0, stream));
        for (int c = 0; c < n_chunks; ++c) {
            const long long c0 = c * per, c1 = std::min(B, c0 + per);
            CXB_CUDA(cudaMemcpy2DAsync(y.p + c0 * e1, (size_t)B * e1, (const unsigned char*)y_host + c0 * e1, (size_t)B * e1,
                                       (size_t)(c1 - c0) * e1, (size_t)T, cudaMemcpyHostToDevice, s_in));
            CXB_CUDA(cudaEventRecord(ev_in[c], s_in));
            CXB_CUDA(cudaStreamWaitEvent(stream, ev_in[c], 0));
            rb0 = c0;
            rb1 = c1;
            int32_t st = dtype == CXB_F32 ? dispatch<float>() : dispatch<double>();
            if (st) return st;
            CXB_CUDA(cudaEventRecord(ev_k[c], stream));
            CXB_CUDA(cudaStreamWaitEvent(s_out, ev_k[c], 0));
            CXB_CUDA(cudaMemcpy2DAsync((unsigned char*)marg_out + c0 * e2, (size_t)B * e2, marg_dev + c0 * e2, (size_t)B * e2,
                                       (size_t)(c1 - c0) * e2, (size_t)T, cudaMemcpyDeviceToHost, s_out));
        }
        rb0 = 0;
        rb1 = B;
        CXB_CUDA(cudaEventRecord(ev1, stream));
        CXB_CUDA(cudaEventRecord(ev_done, s_out));
        CXB_CUDA(cudaStreamWaitEvent(stream, ev_done, 0));  // later work on `stream` is ordered after the copies
        CXB_CUDA(cudaGetLastError());
        have_obs = true;
        ran = true;
        CXB_CUDA(cudaStreamSynchronize(stream));
        return CXB_OK;
    }
};

}  // namespace cxb

using cxb::Chains;
static inline Chains* CH(cxb_chains* c) { return reinterpret_cast<Chains*>(c); }

extern "C" {

int32_t cxb_chains_create(int32_t device, int32_t dtype, int64_t n_chains, int64_t n_steps, cxb_chains** out) try {
    if (!out || (dtype != CXB_F32 && dtype != CXB_F64)) return CXB_ERR_BAD_ARG;
    *out = nullptr;
    Chains* c = new Chains();
    c->device = device;
    c->dtype = dtype;
    c->B = n_chains;
    c->T = n_steps;
    int32_t st = c->init();
    if (st) {
        fprintf(stderr, "cxb_chains_create: %s\n", c->err.c_str());
        delete c;
        return st;
    }
    *out = reinterpret_cast<cxb_chains*>(c);
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
void cxb_chains_destroy(cxb_chains* c) {
    if (c) {
        cudaSetDevice(CH(c)->device);
        delete CH(c);
    }
}
const char* cxb_chains_last_error(cxb_chains* c) { return c ? CH(c)->err.c_str() : "null handle"; }
int32_t cxb_chains_set_noise(cxb_chains* c, const double* q, const double* r) try { return CH(c)->set_noise(q, r); } CXB_ABI_CATCH(CXB_ERR_INTERNAL)

#define CH_CUDA(c, expr)                                        \
    do {                                                        \
        cudaError_t e__ = (expr);                               \
        if (e__ != cudaSuccess) {                               \
            CH(c)->err = ::cxb::cuda_msg(e__, #expr);           \
            return CXB_ERR_CUDA;                                \
        }                                                       \
    } while (0)

int32_t cxb_chains_set_observations(cxb_chains* c, const void* y_host) try {
    Chains* h = CH(c);
    CH_CUDA(c, cudaSetDevice(h->device));
    CH_CUDA(c, cudaMemcpyAsync(h->y.p, y_host, (size_t)h->T * h->B * h->esz(), cudaMemcpyHostToDevice, h->stream));
    CH_CUDA(c, cudaStreamSynchronize(h->stream));
    h->have_obs = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_chains_set_observations_device(cxb_chains* c, const void* y_dev) try {
    Chains* h = CH(c);
    CH_CUDA(c, cudaSetDevice(h->device));
    if (y_dev != h->y.p)
        CH_CUDA(c, cudaMemcpyAsync(h->y.p, y_dev, (size_t)h->T * h->B * h->esz(), cudaMemcpyDeviceToDevice, h->stream));
    h->have_obs = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_chains_update_marginals(cxb_chains* c, int64_t* n_updates_out) try {
    Chains* h = CH(c);
    int32_t st = h->launch();
    if (st) return st;
    if (n_updates_out) *n_updates_out = h->B * (6 * h->T - 4);
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_chains_get_messages(cxb_chains* c, int32_t m, void* out_host) try {
    Chains* h = CH(c);
    if (m < 0 || m > 5) {
        h->err = "message class must be 0..5";
        return CXB_ERR_BAD_ARG;
    }
    if (!h->ran) {
        h->err = "no update has run yet";
        return CXB_ERR_STATE;
    }
    CH_CUDA(c, cudaSetDevice(h->device));
    CH_CUDA(c, cudaMemcpyAsync(out_host, h->msg.p + (size_t)m * h->plane_bytes(), h->plane_bytes(), cudaMemcpyDeviceToHost,
                               h->stream));
    CH_CUDA(c, cudaStreamSynchronize(h->stream));
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_chains_get_marginals(cxb_chains* c, void* out_host) try { return cxb_chains_get_messages(c, 5, out_host); } CXB_ABI_CATCH(CXB_ERR_INTERNAL)
void* cxb_chains_device_ptr(cxb_chains* c, int32_t which) {
    Chains* h = CH(c);
    if (which >= 0 && which <= 5) return h->msg.p + (size_t)which * h->plane_bytes();
    if (which == 6) return h->y.p;
    return nullptr;
}
int32_t cxb_chains_infer_host(cxb_chains* c, const void* y_host, void* marg_out, int64_t* n_updates_out) try {
    Chains* h = CH(c);
    int32_t st = h->infer_host(y_host, marg_out);
    if (st) return st;
    if (n_updates_out) *n_updates_out = h->B * (6 * h->T - 4);
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
void* cxb_chains_stream(cxb_chains* c) { return (void*)CH(c)->stream; }
int32_t cxb_chains_last_kernel_ms(cxb_chains* c, float* ms_out) try {
    Chains* h = CH(c);
    if (!h->ran) {
        h->err = "no update has run yet";
        return CXB_ERR_STATE;
    }
    CH_CUDA(c, cudaEventSynchronize(h->ev1));
    CH_CUDA(c, cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_chains_sync(cxb_chains* c) try {
    CH_CUDA(c, cudaStreamSynchronize(CH(c)->stream));
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)

}  // extern "C"
