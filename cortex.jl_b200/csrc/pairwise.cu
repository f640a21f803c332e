// pairwise.cu — structured engine for loopy BP on an ARBITRARY pairwise categorical graph by synchronous sweeps
// (BASELINE config 5: power-law graph, K = 8).  The irregular analogue of grid.cu.
//
// Graph (identical to tests/models.py:make_powerlaw_model): variables 0..n-1; one unary (leaf) factor per variable
// (ids below every pairwise factor); pairwise factor f = (u_f < v_f) with table psi_{t_f}[x_u][x_v]; for a variable
// the connected factors in ascending id order are  unary < its pairwise factors in ascending f.  One sweep =
// protocol B = re-assert the unary evidence + update_marginals!(engine, all):
//   m2v(v,f)[x_v]  = normalise( sum_{x_u} psi_f(x_u,x_v) * m2f(u,f)[x_u] )      from the PREVIOUS sweep's m2f
//   marginal(v)    = normalise( unary_v * prod_f m2v(v,f) )
//   m2f(v,f)       = normalise( unary_v * prod_{g != f} m2v(v,g) )
// which in the reference is 4m + n + (#ProductOfMessages nodes) signal updates (variables with more than 5 factors
// go through the segment tree of src/dependencies.jl:90-173; here their exclusive products are computed by a
// renormalised prefix/suffix scan — same values up to rounding, no tree nodes materialised).
//
// Layout: messages live in per-variable CSR order ("slot" p = position of (v,f) in v's adjacency), so everything a
// variable WRITES (its m2v, m2f and marginal) is contiguous; the only irregular access is the gather of the
// neighbour's previous m2f (one K*4-byte message = one 32-byte sector at K=8 fp32).  m2f is double buffered.
// Load balance by degree: variables with <= 4 pairwise factors are handled by a group of K lanes entirely in
// registers (lane a owns state a, contractions by group shuffles); larger ones get one CTA each (NG groups scan
// segments of the adjacency, partial products are combined through shared memory).
// HBM-bound: (d+1)*K*4 B read + (2d+1)*K*4 B written per variable of degree d.
#include <algorithm>
#include <cmath>
#include <type_traits>

#include "common.cuh"

namespace cxb {

constexpr int PW_SMALL_MAX = 4;  // pairwise factors per variable handled by the register path (unary + 4 = 5 = n<=5 path)

struct PwView {
    int n_tables;
    const uint32_t* adj_off;   // [n+1] slots of variable v
    const uint32_t* opp;       // [P] slot of the opposite directed edge (the neighbour's message towards the same factor)
    const uint8_t* tsel;       // [P] table id * 2 + (1 if this variable is the HIGHER endpoint of the factor)
    const void* tables;        // [n_tables][2][K][K]: [0] = psi[x_lo][x_hi], [1] = its transpose
    const void* unary;         // [n][K]
    const void* m2f_cur;       // [P][K]
    void* m2f_nxt;             // [P][K]
    void* m2v;                 // [P][K]
    void* marg;                // [n][K]
};

// Group helpers: the K lanes of a group name only themselves in the shuffle masks, so groups of one warp may
// run different trip counts (segments of different length) without deadlocking each other.
template <int K>
__device__ __forceinline__ unsigned group_mask() {
    return K == 32 ? 0xffffffffu : (((1u << (K & 31)) - 1u) << ((threadIdx.x & 31) / K * K));
}
template <class T, int K>
__device__ __forceinline__ T gsum(T v, unsigned gm) {
#pragma unroll
    for (int o = K / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gm, v, o, K);
    return v;
}
// ---- building blocks (K lanes per variable, lane a owns state a) -------------------------------------------------------
// The incoming message is loaded WHOLE by every lane of the group (K*sizeof(T) bytes, the same sector(s) for the
// K lanes: one DRAM/L2 access, no shuffles), the lane's table row [a][0..K) is read with 128-bit shared loads.
template <class T, int K>
__device__ __forceinline__ void load_msg(const T* p, T (&m)[K]) {
    constexpr int V = 16 / sizeof(T);
    using VT = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    if (K % V == 0) {
#pragma unroll
        for (int i = 0; i < K / V; ++i) {
            VT q = __ldg(reinterpret_cast<const VT*>(p) + i);
            const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int j = 0; j < V; ++j) m[i * V + j] = s[j];
        }
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i) m[i] = __ldg(p + i);
    }
}
// out[a] = sum_b psi(b -> a) in[b]; row = the lane's table row: row[b] = weight of in[b]
template <class T, int K>
__device__ __forceinline__ T contract_row(const T* row, const T (&in)[K]) {
    constexpr int V = 16 / sizeof(T);
    using VT = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    T acc0 = T(0), acc1 = T(0);
    if (K % V == 0) {
#pragma unroll
        for (int i = 0; i < K / V; ++i) {
            VT q = *(reinterpret_cast<const VT*>(row) + i);
            const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if ((j & 1) == 0)
                    acc0 = fma(s[j], in[i * V + j], acc0);
                else
                    acc1 = fma(s[j], in[i * V + j], acc1);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i) acc0 = fma(row[i], in[i], acc0);
    }
    return acc0 + acc1;
}
template <class T>
__device__ __forceinline__ T recip(T x);
template <>
__device__ __forceinline__ float recip<float>(float x) {
    return __frcp_rn(x);
}
template <>
__device__ __forceinline__ double recip<double>(double x) {
    return 1.0 / x;
}
// normalise the K-vector spread over the group's lanes to sum 1 (one reciprocal, one multiply per lane)
template <class T, int K>
__device__ __forceinline__ T gnorm(T v, unsigned gm) {
    return v * recip<T>(gsum<T, K>(v, gm));
}
// the lane's table row for slot selector `sel`: this variable higher endpoint -> out[a=x_hi] needs psi[b=x_lo][a] =
// transpose block row a; lower endpoint -> out[a=x_lo] needs psi[a][b] = plain block row a
template <class T, int K>
__device__ __forceinline__ const T* table_row(const T* sh_tables, int sel, int a) {
    return sh_tables + (((size_t)(sel >> 1) * 2 + ((sel & 1) ? 1 : 0)) * K + a) * K;
}

// ---- variables with <= DMAX pairwise factors: one group of K lanes per variable, all messages in registers ---------
// DMAX = 4  : the reference's n <= 5 path (src/dependencies.jl:60-88): products left to right over "all others";
// DMAX > 4  : the reference uses the segment tree; here exclusive products by a renormalised prefix/suffix scan.
// Groups name only their own lanes in the shuffle masks, so each group runs exactly its own degree.
template <class T, int K, int DMAX>
__global__ void __launch_bounds__(256) k_pw_reg(PwView g, const uint32_t* __restrict__ vars, uint32_t n_vars) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sh_tables = reinterpret_cast<T*>(smem_raw);
    for (int x = threadIdx.x; x < g.n_tables * 2 * K * K; x += blockDim.x) sh_tables[x] = ((const T*)g.tables)[x];
    __syncthreads();
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / K;
    const int a = threadIdx.x % K;
    const unsigned gm = group_mask<K>();
    if (gid >= n_vars) return;  // whole groups leave together
    const uint32_t v = vars[gid];
    const uint32_t p0 = g.adj_off[v], d = g.adj_off[v + 1] - p0;
    const T un = __ldg((const T*)g.unary + (size_t)v * K + a);
    uint32_t op[DMAX];
    int sel[DMAX];
    T x[DMAX];
#pragma unroll
    for (int k = 0; k < DMAX; ++k)
        if ((uint32_t)k < d) {
            op[k] = __ldg(g.opp + p0 + k);
            sel[k] = __ldg(g.tsel + p0 + k);
        }
#pragma unroll
    for (int k = 0; k < DMAX; ++k)
        if ((uint32_t)k < d) {
            T in[K];
            load_msg<T, K>((const T*)g.m2f_cur + (size_t)op[k] * K, in);
            T m = gnorm<T, K>(contract_row<T, K>(table_row<T, K>(sh_tables, sel[k], a), in), gm);
            x[k] = m;
            __stcs((T*)g.m2v + (size_t)(p0 + k) * K + a, m);
        }
    if (DMAX <= PW_SMALL_MAX) {
        T acc = un;
#pragma unroll
        for (int k = 0; k < DMAX; ++k)
            if ((uint32_t)k < d) acc = acc * x[k];
        __stcs((T*)g.marg + (size_t)v * K + a, gnorm<T, K>(acc, gm));
#pragma unroll
        for (int k = 0; k < DMAX; ++k)
            if ((uint32_t)k < d) {
                T o = un;
#pragma unroll
                for (int j = 0; j < DMAX; ++j)
                    if (j != k && (uint32_t)j < d) o = o * x[j];
                __stcs((T*)g.m2f_nxt + (size_t)(p0 + k) * K + a, gnorm<T, K>(o, gm));
            }
    } else {
        T pre[DMAX];  // pre[k] = normalise(unary * x_0 * ... * x_{k-1})
        pre[0] = un;
#pragma unroll
        for (int k = 1; k < DMAX; ++k)
            if ((uint32_t)k < d) pre[k] = gnorm<T, K>(pre[k - 1] * x[k - 1], gm);
        T suf = T(1);
#pragma unroll
        for (int k = DMAX - 1; k >= 0; --k)
            if ((uint32_t)k < d) {
                if ((uint32_t)k == d - 1) __stcs((T*)g.marg + (size_t)v * K + a, gnorm<T, K>(pre[k] * x[k], gm));
                __stcs((T*)g.m2f_nxt + (size_t)(p0 + k) * K + a, gnorm<T, K>(pre[k] * suf, gm));
                suf = gnorm<T, K>(suf * x[k], gm);
            }
    }
}

// ---- medium hubs (17 .. 16*NG pairwise factors): one CTA per variable, NG groups, each group keeps its segment of
// <= 16 messages in registers (same code shape as k_pw_reg) and the segments are combined through shared memory.
template <class T, int K, int NG>
__global__ void __launch_bounds__(NG * K) k_pw_hub16(PwView g, const uint32_t* __restrict__ vars, uint32_t n_vars) {
    constexpr int S = 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sh_tables = reinterpret_cast<T*>(smem_raw);
    T* sh_part = sh_tables + (size_t)g.n_tables * 2 * K * K;  // [NG][K] segment products
    for (int x = threadIdx.x; x < g.n_tables * 2 * K * K; x += blockDim.x) sh_tables[x] = ((const T*)g.tables)[x];
    __syncthreads();
    const int grp = threadIdx.x / K, a = threadIdx.x % K;
    const unsigned gm = group_mask<K>();
    for (uint32_t hv = blockIdx.x; hv < n_vars; hv += gridDim.x) {
        const uint32_t v = vars[hv];
        const uint32_t p0 = g.adj_off[v], d = g.adj_off[v + 1] - p0;
        const uint32_t seg = (d + NG - 1) / NG;  // <= S by construction of the bin
        const uint32_t lo = min(d, grp * seg), n_loc = min(d, lo + seg) - lo;
        const uint32_t pb = p0 + lo;
        const T un = __ldg((const T*)g.unary + (size_t)v * K + a);
        uint32_t op[S];
        int sel[S];
        T x[S];
#pragma unroll
        for (int k = 0; k < S; ++k)
            if ((uint32_t)k < n_loc) {
                op[k] = __ldg(g.opp + pb + k);
                sel[k] = __ldg(g.tsel + pb + k);
            }
        T prod = T(1);
#pragma unroll
        for (int k = 0; k < S; ++k)
            if ((uint32_t)k < n_loc) {
                T in[K];
                load_msg<T, K>((const T*)g.m2f_cur + (size_t)op[k] * K, in);
                T m = gnorm<T, K>(contract_row<T, K>(table_row<T, K>(sh_tables, sel[k], a), in), gm);
                x[k] = m;
                __stcs((T*)g.m2v + (size_t)(pb + k) * K + a, m);
                prod = gnorm<T, K>(prod * m, gm);
            }
        sh_part[grp * K + a] = prod;
        __syncthreads();
        T pre0 = un, suf = T(1);
        for (int h = 0; h < grp; ++h) pre0 = gnorm<T, K>(pre0 * sh_part[h * K + a], gm);
        for (int h = NG - 1; h > grp; --h) suf = gnorm<T, K>(suf * sh_part[h * K + a], gm);
        if (grp == NG - 1) __stcs((T*)g.marg + (size_t)v * K + a, gnorm<T, K>(pre0 * prod, gm));
        T pre[S];
        pre[0] = pre0;
#pragma unroll
        for (int k = 1; k < S; ++k)
            if ((uint32_t)k < n_loc) pre[k] = gnorm<T, K>(pre[k - 1] * x[k - 1], gm);
#pragma unroll
        for (int k = S - 1; k >= 0; --k)
            if ((uint32_t)k < n_loc) {
                __stcs((T*)g.m2f_nxt + (size_t)(pb + k) * K + a, gnorm<T, K>(pre[k] * suf, gm));
                suf = gnorm<T, K>(suf * x[k], gm);
            }
        __syncthreads();  // sh_part is reused by the next hub
    }
}

// ---- big hubs: one CTA per variable, NG groups of K lanes, segments of any length ----------------------------------------
// pass 1: every group walks its contiguous segment of the adjacency, 4 slots at a time (4 gathers in flight): m2v per
//         slot (stored) and the segment product;
// pass 2: exclusive prefix (unary * earlier segments) / suffix (later segments) per group through shared memory;
// pass 3: forward over the segment stores the running exclusive prefix in m2f_nxt (scratch), backward combines it with
//         the running suffix into the final m2f. Products are renormalised at every step (as every BP message is).
template <class T, int K, int NG>
__global__ void __launch_bounds__(NG * K) k_pw_hub(PwView g, const uint32_t* __restrict__ vars, uint32_t n_vars) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sh_tables = reinterpret_cast<T*>(smem_raw);
    T* sh_part = sh_tables + (size_t)g.n_tables * 2 * K * K;  // [NG][K] segment products
    for (int x = threadIdx.x; x < g.n_tables * 2 * K * K; x += blockDim.x) sh_tables[x] = ((const T*)g.tables)[x];
    __syncthreads();
    const int grp = threadIdx.x / K, a = threadIdx.x % K;
    const unsigned gm = group_mask<K>();
    constexpr int U = 4;
    for (uint32_t hv = blockIdx.x; hv < n_vars; hv += gridDim.x) {
        const uint32_t v = vars[hv];
        const uint32_t p0 = g.adj_off[v], d = g.adj_off[v + 1] - p0;
        const uint32_t seg = (d + NG - 1) / NG;
        const uint32_t lo = min(d, grp * seg), hi = min(d, lo + seg);
        const T un = __ldg((const T*)g.unary + (size_t)v * K + a);
        // pass 1
        T prod = T(1);
        for (uint32_t k0 = lo; k0 < hi; k0 += U) {
            uint32_t op[U];
            int sel[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (k0 + u < hi) {
                    op[u] = __ldg(g.opp + p0 + k0 + u);
                    sel[u] = __ldg(g.tsel + p0 + k0 + u);
                }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (k0 + u < hi) {
                    T in[K];
                    load_msg<T, K>((const T*)g.m2f_cur + (size_t)op[u] * K, in);
                    T m = gnorm<T, K>(contract_row<T, K>(table_row<T, K>(sh_tables, sel[u], a), in), gm);
                    ((T*)g.m2v)[(size_t)(p0 + k0 + u) * K + a] = m;
                    prod = gnorm<T, K>(prod * m, gm);
                }
        }
        sh_part[grp * K + a] = prod;
        __syncthreads();
        // pass 2
        T pre = un, suf = T(1);
        for (int h = 0; h < grp; ++h) pre = gnorm<T, K>(pre * sh_part[h * K + a], gm);
        for (int h = NG - 1; h > grp; --h) suf = gnorm<T, K>(suf * sh_part[h * K + a], gm);
        if (grp == NG - 1) ((T*)g.marg)[(size_t)v * K + a] = gnorm<T, K>(pre * prod, gm);
        // pass 3 forward: exclusive prefixes into the scratch
        T run = pre;
        for (uint32_t k = lo; k < hi; ++k) {
            ((T*)g.m2f_nxt)[(size_t)(p0 + k) * K + a] = run;
            run = gnorm<T, K>(run * ((const T*)g.m2v)[(size_t)(p0 + k) * K + a], gm);
        }
        // pass 3 backward: m2f = prefix * suffix
        run = suf;
        for (uint32_t k = hi; k > lo; --k) {
            const size_t o = (size_t)(p0 + k - 1) * K + a;
            ((T*)g.m2f_nxt)[o] = gnorm<T, K>(((const T*)g.m2f_nxt)[o] * run, gm);
            run = gnorm<T, K>(run * ((const T*)g.m2v)[o], gm);
        }
        __syncthreads();  // sh_part is reused by the next hub
    }
}

template <class T>
__global__ void k_pw_fill(T* p, size_t n, T v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// gather per-slot planes into factor order: out[2f + side] = plane[slot_of_edge[2f + side]]
template <class T>
__global__ void k_pw_gather(const T* plane, const uint32_t* slot_of_edge, T* out, size_t n_edges, int K) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_edges * K) out[i] = plane[(size_t)slot_of_edge[i / K] * K + (i % K)];
}

struct Pairwise {
    int device = 0, dtype = CXB_F32, K = 0, n_tables = 0, cur = 0;
    long long n = 0, m = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    DBuf<uint32_t> adj_off, opp, small_vars, r8_vars, r16_vars, mid_vars, mid2_vars, big_vars, slot_of_edge;
    DBuf<uint8_t> tsel;
    DBuf<unsigned char> tables, unary, m2f[2], m2v, marg, scratch;
    uint32_t n_small = 0, n_r8 = 0, n_r16 = 0, n_mid = 0, n_mid2 = 0, n_big = 0;
    long long n_products = 0;
    bool have_graph = false, have_tables = false, have_unary = false, have_msgs = false, ran = false;
    size_t esz() const { return dtype == CXB_F32 ? 4 : 8; }
    ~Pairwise() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
    int32_t init() {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            err = "no CUDA device available (cortex_b200 has no CPU fallback)";
            return CXB_ERR_CUDA;
        }
        bool okK = K == 2 || K == 4 || K == 8 || K == 16 || K == 32;
        if (device < 0 || device >= count || n <= 0 || m < 0 || !okK || n_tables < 1 || n_tables > 127 || 2 * m >= 4000000000LL) {
            err = "bad device / shape (states must be 2,4,8,16 or 32; 1..127 tables)";
            return CXB_ERR_BAD_ARG;
        }
        CXB_CUDA(cudaSetDevice(device));
        CXB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CXB_CUDA(cudaEventCreate(&ev0));
        CXB_CUDA(cudaEventCreate(&ev1));
        return CXB_OK;
    }
    int32_t set_graph(const int64_t* fu, const int64_t* fv, const int32_t* ft) {
        CXB_CUDA(cudaSetDevice(device));
        const size_t P = (size_t)2 * m;
        std::vector<uint32_t> off((size_t)n + 1, 0);
        for (long long f = 0; f < m; ++f) {
            if (fu[f] < 0 || fv[f] >= n || fu[f] >= fv[f] || ft[f] < 0 || ft[f] >= n_tables) {
                err = "pairwise factors must satisfy 0 <= u < v < n_variables and 0 <= table < n_tables";
                return CXB_ERR_BAD_ARG;
            }
            ++off[(size_t)fu[f] + 1];
            ++off[(size_t)fv[f] + 1];
        }
        for (long long i = 0; i < n; ++i) off[i + 1] += off[i];
        std::vector<uint32_t> cursor(off.begin(), off.end() - 1), slot(P), oppv(P);
        std::vector<uint8_t> sel(P);
        for (long long f = 0; f < m; ++f) {  // ascending f => each adjacency is in ascending factor id
            uint32_t pu = cursor[fu[f]]++, pv = cursor[fv[f]]++;
            slot[2 * f] = pu;
            slot[2 * f + 1] = pv;
            oppv[pu] = pv;
            oppv[pv] = pu;
            sel[pu] = (uint8_t)(ft[f] * 2 + 0);  // u is the lower endpoint
            sel[pv] = (uint8_t)(ft[f] * 2 + 1);
        }
        // degree bins: <= 4 register path (reference n<=5 order), 5..8 and 9..16 register prefix/suffix,
        // <= 16*(32/K) one warp-sized CTA, <= 16*(256/K) one 256-thread CTA (segments in registers), larger: streamed
        std::vector<uint32_t> sm, r8, r16, md, md2, bg;
        const uint32_t mid_max = 16u * (uint32_t)std::max(1, 32 / K), mid2_max = 16u * (uint32_t)(256 / K);
        n_products = 0;
        for (long long v = 0; v < n; ++v) {
            uint32_t d = off[v + 1] - off[v];
            if (d <= PW_SMALL_MAX) {
                sm.push_back((uint32_t)v);
                continue;
            }
            n_products += (long long)d - 1;  // (d+1) factors incl. the unary -> d-1 ProductOfMessages nodes in the reference
            if (d <= 8)
                r8.push_back((uint32_t)v);
            else if (d <= 16)
                r16.push_back((uint32_t)v);
            else
                (d <= mid_max ? md : d <= mid2_max ? md2 : bg).push_back((uint32_t)v);
        }
        auto by_degree_desc = [&](uint32_t x, uint32_t y) { return off[x + 1] - off[x] > off[y + 1] - off[y]; };
        std::stable_sort(md.begin(), md.end(), by_degree_desc);
        std::stable_sort(md2.begin(), md2.end(), by_degree_desc);
        std::stable_sort(bg.begin(), bg.end(), by_degree_desc);
        n_small = (uint32_t)sm.size();
        n_r8 = (uint32_t)r8.size();
        n_r16 = (uint32_t)r16.size();
        n_mid = (uint32_t)md.size();
        n_mid2 = (uint32_t)md2.size();
        n_big = (uint32_t)bg.size();
        auto up = [&](auto& dbuf, const auto& vec) -> cudaError_t {
            cudaError_t e = dbuf.reserve(vec.size());
            if (e != cudaSuccess) return e;
            return vec.empty() ? cudaSuccess
                               : cudaMemcpyAsync(dbuf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice, stream);
        };
        CXB_CUDA(up(adj_off, off));
        CXB_CUDA(up(opp, oppv));
        CXB_CUDA(up(tsel, sel));
        CXB_CUDA(up(slot_of_edge, slot));
        CXB_CUDA(up(small_vars, sm));
        CXB_CUDA(up(r8_vars, r8));
        CXB_CUDA(up(r16_vars, r16));
        CXB_CUDA(up(mid_vars, md));
        CXB_CUDA(up(mid2_vars, md2));
        CXB_CUDA(up(big_vars, bg));
        size_t pb = std::max<size_t>(P, 1) * K * esz(), nb = (size_t)n * K * esz();
        CXB_CUDA(m2f[0].reserve(pb));
        CXB_CUDA(m2f[1].reserve(pb));
        CXB_CUDA(m2v.reserve(pb));
        CXB_CUDA(scratch.reserve(pb));
        CXB_CUDA(marg.reserve(nb));
        CXB_CUDA(unary.reserve(nb));
        CXB_CUDA(tables.reserve((size_t)n_tables * 2 * K * K * esz()));
        CXB_CUDA(cudaMemsetAsync(m2v.p, 0, pb, stream));
        CXB_CUDA(cudaMemsetAsync(marg.p, 0, nb, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        have_graph = true;
        return CXB_OK;
    }
    int32_t set_tables(const double* tb) {
        if (!have_graph) {
            err = "set the graph first";
            return CXB_ERR_STATE;
        }
        CXB_CUDA(cudaSetDevice(device));
        size_t cnt = (size_t)n_tables * 2 * K * K;
        std::vector<unsigned char> raw(cnt * esz());
        for (int t = 0; t < n_tables; ++t)
            for (int i = 0; i < K; ++i)
                for (int j = 0; j < K; ++j) {
                    double v = tb[((size_t)t * K + i) * K + j];
                    size_t o0 = (((size_t)t * 2 + 0) * K + i) * K + j, o1 = (((size_t)t * 2 + 1) * K + j) * K + i;
                    if (dtype == CXB_F32) {
                        ((float*)raw.data())[o0] = (float)v;
                        ((float*)raw.data())[o1] = (float)v;
                    } else {
                        ((double*)raw.data())[o0] = v;
                        ((double*)raw.data())[o1] = v;
                    }
                }
        CXB_CUDA(cudaMemcpyAsync(tables.p, raw.data(), raw.size(), cudaMemcpyHostToDevice, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        have_tables = true;
        return CXB_OK;
    }
    int32_t reset() {
        if (!have_graph) {
            err = "set the graph first";
            return CXB_ERR_STATE;
        }
        CXB_CUDA(cudaSetDevice(device));
        size_t cnt = (size_t)2 * m * K;
        for (int b = 0; b < 2 && cnt; ++b) {
            if (dtype == CXB_F32)
                CXB_LAUNCH(k_pw_fill<float>, cdiv(cnt, 256), 256, 0, stream, (float*)m2f[b].p, cnt, 1.0f / K);
            else
                CXB_LAUNCH(k_pw_fill<double>, cdiv(cnt, 256), 256, 0, stream, (double*)m2f[b].p, cnt, 1.0 / K);
        }
        CXB_CUDA(cudaGetLastError());
        cur = 0;
        have_msgs = true;
        return CXB_OK;
    }
    template <class T, int KK>
    int32_t launch_k(const PwView& g) {
        size_t tb = (size_t)n_tables * 2 * KK * KK * sizeof(T);
        if (tb + 64 * KK * sizeof(T) > 200 * 1024) {
            err = "tables do not fit in shared memory";
            return CXB_ERR_BAD_ARG;
        }
        auto attr = [&](auto kern, size_t smem) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        };
        if (n_small) {
            attr(k_pw_reg<T, KK, PW_SMALL_MAX>, tb);
            CXB_LAUNCH((k_pw_reg<T, KK, PW_SMALL_MAX>), cdiv((size_t)n_small * KK, 256), 256, tb, stream, g, small_vars.p, n_small);
        }
        if (n_r8) {
            attr(k_pw_reg<T, KK, 8>, tb);
            CXB_LAUNCH((k_pw_reg<T, KK, 8>), cdiv((size_t)n_r8 * KK, 256), 256, tb, stream, g, r8_vars.p, n_r8);
        }
        if (n_r16) {
            attr(k_pw_reg<T, KK, 16>, tb);
            CXB_LAUNCH((k_pw_reg<T, KK, 16>), cdiv((size_t)n_r16 * KK, 256), 256, tb, stream, g, r16_vars.p, n_r16);
        }
        if (n_mid) {  // 17 .. 16*NG1 factors: one warp-sized CTA per variable, segments in registers
            constexpr int NG = 32 / KK > 0 ? 32 / KK : 1;
            size_t smem = tb + (size_t)NG * KK * sizeof(T);
            attr(k_pw_hub16<T, KK, NG>, smem);
            CXB_LAUNCH((k_pw_hub16<T, KK, NG>), std::min<uint32_t>(n_mid, 148u * 64u), NG * KK, smem, stream, g, mid_vars.p, n_mid);
        }
        if (n_mid2) {  // up to 16*(256/K) factors: 256-thread CTA, segments in registers
            constexpr int NG = 256 / KK;
            size_t smem = tb + (size_t)NG * KK * sizeof(T);
            attr(k_pw_hub16<T, KK, NG>, smem);
            CXB_LAUNCH((k_pw_hub16<T, KK, NG>), std::min<uint32_t>(n_mid2, 148u * 8u), NG * KK, smem, stream, g, mid2_vars.p, n_mid2);
        }
        if (n_big) {  // anything larger: 256-thread CTA, segments streamed
            constexpr int NG = 256 / KK;
            size_t smem = tb + (size_t)NG * KK * sizeof(T);
            attr(k_pw_hub<T, KK, NG>, smem);
            CXB_LAUNCH((k_pw_hub<T, KK, NG>), std::min<uint32_t>(n_big, 148u * 8u), NG * KK, smem, stream, g, big_vars.p, n_big);
        }
        return CXB_OK;
    }
    template <class T>
    int32_t launch_t(const PwView& g) {
        switch (K) {
            case 2: return launch_k<T, 2>(g);
            case 4: return launch_k<T, 4>(g);
            case 8: return launch_k<T, 8>(g);
            case 16: return launch_k<T, 16>(g);
            default: return launch_k<T, 32>(g);
        }
    }
    int32_t sweep(int64_t* n_updates) {
        if (!have_graph || !have_tables || !have_unary || !have_msgs) {
            err = "set the graph, the tables, the unary evidence and reset the messages first";
            return CXB_ERR_STATE;
        }
        CXB_CUDA(cudaSetDevice(device));
        PwView g;
        g.n_tables = n_tables;
        g.adj_off = adj_off.p;
        g.opp = opp.p;
        g.tsel = tsel.p;
        g.tables = tables.p;
        g.unary = unary.p;
        g.m2f_cur = m2f[cur].p;
        g.m2f_nxt = m2f[cur ^ 1].p;
        g.m2v = m2v.p;
        g.marg = marg.p;
        CXB_CUDA(cudaEventRecord(ev0, stream));
        int32_t st = dtype == CXB_F32 ? launch_t<float>(g) : launch_t<double>(g);
        if (st) return st;
        CXB_CUDA(cudaEventRecord(ev1, stream));
        CXB_CUDA(cudaGetLastError());
        cur ^= 1;
        ran = true;
        if (n_updates) *n_updates = 4 * m + n + n_products;  // m2v + m2f + marginals + ProductOfMessages nodes
        return CXB_OK;
    }
    // per-edge planes in factor order: out[(2f + side)][K], side 0 = the lower endpoint u, 1 = v
    int32_t get_edges(const unsigned char* plane, void* out_host) {
        CXB_CUDA(cudaSetDevice(device));
        size_t E = (size_t)2 * m;
        if (E) {
            if (dtype == CXB_F32)
                CXB_LAUNCH(k_pw_gather<float>, cdiv(E * K, 256), 256, 0, stream, (const float*)plane, slot_of_edge.p, (float*)scratch.p, E, K);
            else
                CXB_LAUNCH(k_pw_gather<double>, cdiv(E * K, 256), 256, 0, stream, (const double*)plane, slot_of_edge.p, (double*)scratch.p,
                           E, K);
            CXB_CUDA(cudaMemcpyAsync(out_host, scratch.p, E * K * esz(), cudaMemcpyDeviceToHost, stream));
        }
        CXB_CUDA(cudaStreamSynchronize(stream));
        return CXB_OK;
    }
};

}  // namespace cxb

using cxb::Pairwise;
static inline Pairwise* PW(cxb_pairwise* g) { return reinterpret_cast<Pairwise*>(g); }
#define PW_CUDA(g, expr)                                        \
    do {                                                        \
        cudaError_t e__ = (expr);                               \
        if (e__ != cudaSuccess) {                               \
            PW(g)->err = ::cxb::cuda_msg(e__, #expr);           \
            return CXB_ERR_CUDA;                                \
        }                                                       \
    } while (0)

extern "C" {

int32_t cxb_pairwise_create(int32_t device, int32_t dtype, int64_t n_variables, int64_t n_factors, int32_t n_states,
                            int32_t n_tables, cxb_pairwise** out) {
    if (!out || (dtype != CXB_F32 && dtype != CXB_F64)) return CXB_ERR_BAD_ARG;
    *out = nullptr;
    Pairwise* g = new Pairwise();
    g->device = device;
    g->dtype = dtype;
    g->n = n_variables;
    g->m = n_factors;
    g->K = n_states;
    g->n_tables = n_tables;
    int32_t st = g->init();
    if (st) {
        fprintf(stderr, "cxb_pairwise_create: %s\n", g->err.c_str());
        delete g;
        return st;
    }
    *out = reinterpret_cast<cxb_pairwise*>(g);
    return CXB_OK;
}
void cxb_pairwise_destroy(cxb_pairwise* g) {
    if (g) {
        cudaSetDevice(PW(g)->device);
        delete PW(g);
    }
}
const char* cxb_pairwise_last_error(cxb_pairwise* g) { return g ? PW(g)->err.c_str() : "null handle"; }
int32_t cxb_pairwise_set_graph(cxb_pairwise* g, const int64_t* fac_u, const int64_t* fac_v, const int32_t* fac_table) {
    return PW(g)->set_graph(fac_u, fac_v, fac_table);
}
int32_t cxb_pairwise_set_tables(cxb_pairwise* g, const double* tables) { return PW(g)->set_tables(tables); }
int32_t cxb_pairwise_set_unary(cxb_pairwise* g, const void* unary_host) {
    Pairwise* h = PW(g);
    if (!h->have_graph) {
        h->err = "set the graph first";
        return CXB_ERR_STATE;
    }
    PW_CUDA(g, cudaSetDevice(h->device));
    PW_CUDA(g, cudaMemcpyAsync(h->unary.p, unary_host, (size_t)h->n * h->K * h->esz(), cudaMemcpyHostToDevice, h->stream));
    PW_CUDA(g, cudaStreamSynchronize(h->stream));
    h->have_unary = true;
    return CXB_OK;
}
int32_t cxb_pairwise_reset_messages(cxb_pairwise* g) { return PW(g)->reset(); }
int32_t cxb_pairwise_sweep(cxb_pairwise* g, int64_t* n_updates_out) { return PW(g)->sweep(n_updates_out); }
int32_t cxb_pairwise_get_marginals(cxb_pairwise* g, void* out_host) {
    Pairwise* h = PW(g);
    PW_CUDA(g, cudaSetDevice(h->device));
    PW_CUDA(g, cudaMemcpyAsync(out_host, h->marg.p, (size_t)h->n * h->K * h->esz(), cudaMemcpyDeviceToHost, h->stream));
    PW_CUDA(g, cudaStreamSynchronize(h->stream));
    return CXB_OK;
}
int32_t cxb_pairwise_get_messages(cxb_pairwise* g, int32_t which, void* out_host) {
    Pairwise* h = PW(g);
    if (which != 0 && which != 1) {
        h->err = "which must be 0 (m2v) or 1 (m2f)";
        return CXB_ERR_BAD_ARG;
    }
    return h->get_edges(which == 0 ? h->m2v.p : h->m2f[h->cur].p, out_host);
}
int64_t cxb_pairwise_algorithmic_bytes(cxb_pairwise* g) {
    Pairwise* h = PW(g);
    // per variable of degree d: (d + 1) messages read, (2d + 1) written (no ProductOfMessages traffic)
    return (int64_t)((size_t)(2 * h->m + h->n) + (size_t)(4 * h->m + h->n)) * h->K * (int64_t)h->esz();
}
void* cxb_pairwise_stream(cxb_pairwise* g) { return (void*)PW(g)->stream; }
int32_t cxb_pairwise_last_kernel_ms(cxb_pairwise* g, float* ms_out) {
    Pairwise* h = PW(g);
    if (!h->ran) {
        h->err = "no sweep has run yet";
        return CXB_ERR_STATE;
    }
    PW_CUDA(g, cudaEventSynchronize(h->ev1));
    PW_CUDA(g, cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
    return CXB_OK;
}
int32_t cxb_pairwise_sync(cxb_pairwise* g) {
    PW_CUDA(g, cudaStreamSynchronize(PW(g)->stream));
    return CXB_OK;
}

}  // extern "C"
