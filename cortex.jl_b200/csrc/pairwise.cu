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
// Layout: messages live in per-variable CSR order ("slot" p = position of (v,f) in v's adjacency).  m2f is stored in
// the RECEIVER's order ("inbox": inbox[p] = the message the other endpoint of slot p's factor sends to that factor), so
// everything a variable READS (inbox, unary, opp, tsel) and its m2v / marginal writes are contiguous streams; the only
// irregular access is the scatter of the outgoing m2f into the neighbours' inboxes: one K*4-byte message = one full
// 32-byte sector at K=8 fp32, fire-and-forget, nothing waits on it (a gather would cost 64 B of DRAM fetch per 32 B
// and sit on the latency chain).  The inbox is double buffered.
// Load balance by degree: variables with <= 4 pairwise factors are handled by a group of K lanes entirely in
// registers (lane a owns state a, contractions by group shuffles); larger ones get one CTA each (NG groups scan
// segments of the adjacency, partial products are combined through shared memory).
// HBM-bound: (d+1)*K*4 B read + (2d+1)*K*4 B written per variable of degree d.
#include <algorithm>
#include <cmath>
#include <type_traits>

#include "common.cuh"

namespace cxb {

constexpr int PW_EXACT_MAX = 4;  // pairwise factors per variable on the reference's n <= 5 path (unary + 4)
constexpr int PW_SEG = 8;        // slots one lane group keeps in registers

struct PwView {
    int n_tables;
    const uint32_t* opp;       // [P] slot of the opposite directed edge (the neighbour's message towards the same factor)
    const uint8_t* tsel;       // [P] table id * 2 + (1 if this variable is the HIGHER endpoint of the factor)
    const void* tables;        // bank-conflict-free shared-memory image of the tables (see Pairwise::set_tables)
    int tab_elems;             // elements in the image
    int tab_block;             // elements between the blocks of consecutive lanes (lane % 8 selects the block)
    int tab_blocks;            // 8 (one block per lane of a 128-bit shared-load phase) or L (one block per lane of a group)
    int tab_sel;               // elements between consecutive (table, orientation) selections inside a block
    int prefetch;              // issue prefetch.global.L2 for the next iteration's streams
    const void* unary;         // [n][K]
    const void* m2f_cur;       // [P][K]
    void* m2f_nxt;             // [P][K]
    void* m2v;                 // [P][K]
    void* marg;                // [n][K]
    void* chunk_prod;          // [n_chunks][K] hub chunks: product of the chunk's m2v
    void* chunk_pre;           // [n_chunks][K] unary * product of the earlier chunks
    void* chunk_suf;           // [n_chunks][K] product of the later chunks
};
// work records in processing order (degree-sorted inside a bin so that the lanes of a warp run the same trip counts)
struct PwBin {
    const uint32_t* v;   // variable
    const uint32_t* p0;  // first slot (of the variable, or of the hub chunk)
    const uint32_t* d;   // number of slots
    uint32_t n;
    uint32_t base;       // global record index of record 0 (the unary evidence is stored in record order)
};

// Lane geometry: a message of K states is spread over L = K / S adjacent lanes, S = min(K, 4) states (one 16-byte
// vector at fp32) per lane.  All shuffles are full-mask: every lane of a warp runs every shuffle (absent slots carry
// the neutral message), the loops are skipped warp-uniformly.
template <int K>
struct PwGeo {
    static constexpr int S = K < 4 ? K : 4;
    static constexpr int L = K / S;
};
constexpr unsigned PW_FULL_MASK = 0xffffffffu;

template <class T, int N>
__device__ __forceinline__ void ld_vec(const T* p, T (&r)[N]) {  // read-only path, 16-byte (or 8-byte) vectors
    constexpr int BYTES = N * (int)sizeof(T);
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) {
            uint4 q = __ldg(reinterpret_cast<const uint4*>(p) + i);
            const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int j = 0; j < 16 / (int)sizeof(T); ++j) r[i * (16 / (int)sizeof(T)) + j] = s[j];
        }
    } else {
        static_assert(BYTES == 8, "vector of 8 or a multiple of 16 bytes");
        uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
        const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int j = 0; j < N; ++j) r[j] = s[j];
    }
}
template <class T, int N>
__device__ __forceinline__ void ld_vec_plain(const T* p, T (&r)[N]) {  // coherent load (data written by an earlier kernel)
    constexpr int BYTES = N * (int)sizeof(T);
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) {
            uint4 q = *(reinterpret_cast<const uint4*>(p) + i);
            const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int j = 0; j < 16 / (int)sizeof(T); ++j) r[i * (16 / (int)sizeof(T)) + j] = s[j];
        }
    } else {
        uint2 q = *reinterpret_cast<const uint2*>(p);
        const T* s = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int j = 0; j < N; ++j) r[j] = s[j];
    }
}
template <bool STREAM, class T, int N>
__device__ __forceinline__ void st_vec(T* p, const T (&r)[N]) {
    constexpr int BYTES = N * (int)sizeof(T);
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) {
            uint4 q;
            T* s = reinterpret_cast<T*>(&q);
#pragma unroll
            for (int j = 0; j < 16 / (int)sizeof(T); ++j) s[j] = r[i * (16 / (int)sizeof(T)) + j];
            if (STREAM)
                __stcs(reinterpret_cast<uint4*>(p) + i, q);
            else
                *(reinterpret_cast<uint4*>(p) + i) = q;
        }
    } else {
        uint2 q;
        T* s = reinterpret_cast<T*>(&q);
#pragma unroll
        for (int j = 0; j < N; ++j) s[j] = r[j];
        if (STREAM)
            __stcs(reinterpret_cast<uint2*>(p), q);
        else
            *reinterpret_cast<uint2*>(p) = q;
    }
}
template <class T, int L>
__device__ __forceinline__ T gsum(T v) {  // sum over the L lanes of a group (groups are aligned, L a power of two)
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(PW_FULL_MASK, v, o);
    return v;
}
template <class T>
__device__ __forceinline__ T recip(T x);
// MUFU.RCP + one Newton step (<= 1 ulp on the normal, positive sums it is used on). __frcp_rn is a ~30-instruction
// subroutine behind a CALL: it was 30 % of the stall samples of the small-degree kernel.
template <>
__device__ __forceinline__ float recip<float>(float x) {
    float q;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(x));
    return fmaf(q, fmaf(-x, q, 1.0f), q);
}
template <>
__device__ __forceinline__ double recip<double>(double x) {
    return 1.0 / x;
}
// 2^-floor(log2 s): multiplying by it is exact, so intermediate products keep every bit of the unscaled product
__device__ __forceinline__ float pow2_inv(float s) {
    unsigned e = (__float_as_uint(s) >> 23) & 0xffu;
    return __uint_as_float((254u - e) << 23);
}
__device__ __forceinline__ double pow2_inv(double s) {
    unsigned e = ((unsigned)__double2hiint(s) >> 20) & 0x7ffu;
    return __hiloint2double((int)((2046u - e) << 20), 0);
}
template <class T, int K>
__device__ __forceinline__ void norm1(T (&v)[PwGeo<K>::S]) {  // normalise the group's K-vector to sum 1
    T s = v[0];
#pragma unroll
    for (int i = 1; i < PwGeo<K>::S; ++i) s += v[i];
    T r = recip<T>(gsum<T, PwGeo<K>::L>(s));
#pragma unroll
    for (int i = 0; i < PwGeo<K>::S; ++i) v[i] *= r;
}
template <class T, int K>
__device__ __forceinline__ void rescale(T (&v)[PwGeo<K>::S]) {  // keep a running product in range (exact power-of-two scaling)
    T s = v[0];
#pragma unroll
    for (int i = 1; i < PwGeo<K>::S; ++i) s += v[i];
    T r = pow2_inv(gsum<T, PwGeo<K>::L>(s));
#pragma unroll
    for (int i = 0; i < PwGeo<K>::S; ++i) v[i] *= r;
}
template <class T, int S>
__device__ __forceinline__ void vmul(T (&a)[S], const T (&b)[S]) {
#pragma unroll
    for (int i = 0; i < S; ++i) a[i] *= b[i];
}
template <class T, int S>
__device__ __forceinline__ void vset(T (&a)[S], T x) {
#pragma unroll
    for (int i = 0; i < S; ++i) a[i] = x;
}
template <class T, int S>
__device__ __forceinline__ void vcopy(T (&a)[S], const T (&b)[S]) {
#pragma unroll
    for (int i = 0; i < S; ++i) a[i] = b[i];
}
// out[r] = sum_b psi(b -> a0 + r) in[b]; rows = the lane's S table rows in shared memory (row a0 of the selected block:
// this variable higher endpoint -> transpose block, lower endpoint -> plain block)
template <class T, int K>
__device__ __forceinline__ void contract(const T* rows, const T (&in)[K], T (&out)[PwGeo<K>::S]) {
    constexpr int S = PwGeo<K>::S;
    constexpr int V = (K * (int)sizeof(T)) % 16 == 0 ? 16 / (int)sizeof(T) : 1;
#pragma unroll
    for (int r = 0; r < S; ++r) {
        T acc0 = T(0), acc1 = T(0);
        if constexpr (V > 1) {
#pragma unroll
            for (int i = 0; i < K / V; ++i) {
                uint4 q = *(reinterpret_cast<const uint4*>(rows + r * K) + i);
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
            for (int i = 0; i < K; ++i) acc0 = fma(rows[r * K + i], in[i], acc0);
        }
        out[r] = acc0 + acc1;
    }
}
template <class T>
__device__ __forceinline__ void stage_tables(T* sh, const PwView& g, int K) {
    for (int x = threadIdx.x; x < g.tab_elems; x += blockDim.x) sh[x] = ((const T*)g.tables)[x];
    __syncthreads();
}
// The lane's private view of the table image: block (lane % 8) % tab_blocks holds, for every selection, the S rows this
// lane's states need; consecutive blocks start 16 bytes (4 banks) further round the 32 banks, so the 8 lanes of a
// 128-bit load phase always hit 8 different bank groups whatever tables their slots use.
template <class T>
__device__ __forceinline__ const T* lane_tables(const T* sh, const PwView& g) {
    return sh + (size_t)(((threadIdx.x & 7) % g.tab_blocks) * g.tab_block);
}

// ---- software pipeline of the persistent loops: the records of iteration i+2 are fetched into registers, the data of
// iteration i+1 (whose records are already here) is pulled into L2, iteration i computes on L2 hits ---------------------------
struct PwRec {
    uint32_t v, p0, d;
    bool live;
};
__device__ __forceinline__ PwRec load_rec(const PwBin& bin, uint32_t idx) {
    PwRec r{0u, 0u, 0u, idx < bin.n};
    if (r.live) {
        r.v = __ldg(bin.v + idx);
        r.p0 = __ldg(bin.p0 + idx);
        r.d = __ldg(bin.d + idx);
    }
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p, size_t bytes) {  // every 128-byte line of [p, p + bytes)
    if (bytes == 0) return;
    const char* c = reinterpret_cast<const char*>(p);
    for (size_t o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + o));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + bytes - 1));
}

// ---- variables with <= 4 pairwise factors: one lane group per variable, the reference's n <= 5 path -----------------------
// (src/dependencies.jl:60-88): marginal = unary * x_0 * x_1 ..., m2f(v, f_k) = unary * prod_{j != k} x_j, left to right.
template <class T, int K, int MINB>
__global__ void __launch_bounds__(256, MINB) k_pw_exact(PwView g, PwBin bin) {
    constexpr int S = PwGeo<K>::S, L = PwGeo<K>::L, D = PW_EXACT_MAX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sh_tables = reinterpret_cast<T*>(smem_raw);
    stage_tables<T>(sh_tables, g, K);
    const T* my_tab = lane_tables<T>(sh_tables, g);
    const int tab_sel = g.tab_sel;
    const int a0 = (threadIdx.x % L) * S;
    // persistent CTAs: the tables are staged once, then the CTA walks the bin with a grid stride (all CTAs work on
    // neighbouring records at the same time, so the streamed arrays stay hot in L2)
    const uint32_t per_cta = blockDim.x / L, stride = gridDim.x * per_cta, li = threadIdx.x % L;
    const uint32_t first = blockIdx.x * per_cta + threadIdx.x / L;
    PwRec cur = load_rec(bin, first), nxt = load_rec(bin, first + stride);
    for (uint32_t base = blockIdx.x * per_cta; base < bin.n; base += stride) {
    const uint32_t gid = first + (base - blockIdx.x * per_cta), gid_next = gid + stride;
    const PwRec nn = load_rec(bin, gid + 2 * stride);
    if (g.prefetch && nxt.live) {  // pull the next iteration's streams into L2 (lane 0 of the group: messages, last lane: indices)
        if (li == 0) {
            prefetch_l2((const T*)g.unary + (size_t)(bin.base + gid_next) * K, K * sizeof(T));
            prefetch_l2((const T*)g.m2f_cur + (size_t)nxt.p0 * K, (size_t)nxt.d * K * sizeof(T));
        }
        if (li == L - 1) {
            prefetch_l2(g.opp + nxt.p0, (size_t)nxt.d * sizeof(uint32_t));
            prefetch_l2(g.tsel + nxt.p0, (size_t)nxt.d);
        }
    }
    const bool live = cur.live;
    const uint32_t v = cur.v, p0 = cur.p0, d = cur.d;
    cur = nxt;
    nxt = nn;
    T un[S];
    vset<T, S>(un, T(1));
    if (live) ld_vec<T, S>((const T*)g.unary + (size_t)(bin.base + gid) * K + a0, un);
    uint32_t op[D];
    int sel[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        op[k] = 0;
        sel[k] = 0;
        if ((uint32_t)k < d) {
            op[k] = __ldg(g.opp + p0 + k);
            sel[k] = __ldg(g.tsel + p0 + k);
        }
    }
    T x[D][S];
    T in[D][K];  // every incoming message is requested before the first one is used (one round trip, not D)
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
        for (int i = 0; i < K; ++i) in[k][i] = T(1);
        if ((uint32_t)k < d) ld_vec<T, K>((const T*)g.m2f_cur + (size_t)(p0 + k) * K, in[k]);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        vset<T, S>(x[k], T(1));
        const bool have = (uint32_t)k < d;
        if (!__any_sync(PW_FULL_MASK, have)) continue;
        T m[S];
        contract<T, K>(my_tab + sel[k] * tab_sel, in[k], m);
        norm1<T, K>(m);
        if (have) {
            st_vec<true, T, S>((T*)g.m2v + (size_t)(p0 + k) * K + a0, m);
            vcopy<T, S>(x[k], m);
        }
    }
    {
        T acc[S];
        vcopy<T, S>(acc, un);
#pragma unroll
        for (int k = 0; k < D; ++k) vmul<T, S>(acc, x[k]);
        norm1<T, K>(acc);
        if (live) st_vec<true, T, S>((T*)g.marg + (size_t)v * K + a0, acc);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const bool have = (uint32_t)k < d;
        if (!__any_sync(PW_FULL_MASK, have)) continue;
        T o[S];
        vcopy<T, S>(o, un);
#pragma unroll
        for (int j = 0; j < D; ++j)
            if (j != k) vmul<T, S>(o, x[j]);
        norm1<T, K>(o);
        if (have) st_vec<true, T, S>((T*)g.m2f_nxt + (size_t)op[k] * K + a0, o);  // into the receiver's inbox
    }
    }  // grid-stride loop
}

// ---- the same path with its streams staged through shared memory by the bulk-copy engine ----------------------------------
// After the slot renumbering the 32 / L records a warp handles in one iteration own ONE contiguous block of slots, so
// everything the warp reads — the inbox messages, the unary evidence (record order), the opp and tsel entries — is four
// contiguous ranges. One elected lane asks the bulk-copy engine (cp.async.bulk, completion on a per-warp mbarrier) for the
// ranges of the iteration after next while the warp computes on the current stage: loads no longer live in registers, two
// iterations of reads are always in flight per warp, and the compute reads 128-bit vectors from shared memory.
namespace pwasync {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "PW_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PW_DONE;\n\t"
        "bra PW_WAIT_LOOP;\n\t"
        "PW_DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
}  // namespace pwasync

constexpr int PW_STAGES = 3;
template <class T, int K>
struct PwStage {
    static constexpr int S = PwGeo<K>::S, L = PwGeo<K>::L, D = PW_EXACT_MAX, GPW = 32 / L;  // GPW = records per warp
    static constexpr uint32_t INBOX_B = GPW * D * K * sizeof(T), UNARY_B = GPW * K * sizeof(T);
    // aligned copy windows: up to 3 + 3 extra opp entries, 15 + 15 extra tsel bytes
    static constexpr uint32_t OPP_B = (GPW * D + 8) * 4, TSEL_B = ((GPW * D + 32 + 15) / 16) * 16;
    static constexpr uint32_t OFF_UNARY = INBOX_B, OFF_OPP = OFF_UNARY + UNARY_B, OFF_TSEL = OFF_OPP + OPP_B;
    static constexpr uint32_t BYTES = ((OFF_TSEL + TSEL_B + 127) / 128) * 128;
};
template <class T, int K>
__global__ void __launch_bounds__(256, 2) k_pw_exact_staged(PwView g, PwBin bin) {
    using ST = PwStage<T, K>;
    constexpr int S = ST::S, L = ST::L, D = ST::D, GPW = ST::GPW, WARPS = 8, NS = PW_STAGES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sh_tables = reinterpret_cast<T*>(smem_raw);
    const size_t tab_bytes = ((size_t)g.tab_elems * sizeof(T) + 127) / 128 * 128;
    unsigned char* stages = smem_raw + tab_bytes;                                     // [WARPS][NS][ST::BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)WARPS * NS * ST::BYTES);  // [WARPS][NS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane / L;
    if (lane == 0) {
        for (int i = 0; i < NS; ++i) pwasync::mbar_init(pwasync::smem_u32(&bars[warp * NS + i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_tables<T>(sh_tables, g, K);  // ends with __syncthreads: barriers initialised, tables staged
    const T* my_tab = lane_tables<T>(sh_tables, g);
    const int tab_sel = g.tab_sel;
    const int a0 = (lane % L) * S;
    const uint32_t per_cta = WARPS * GPW, stride = gridDim.x * per_cta;
    const uint32_t first = blockIdx.x * per_cta + warp * GPW + grp;
    unsigned char* my_stage = stages + (size_t)warp * NS * ST::BYTES;

    // ask the bulk-copy engine for the four ranges of one block of records (all lanes call; lane 0 issues)
    auto issue = [&](const PwRec& r, uint32_t rec_first, int stage) {
        const unsigned live_mask = __ballot_sync(PW_FULL_MASK, r.live);
        if (!live_mask) return;
        const uint32_t blk_start = __shfl_sync(PW_FULL_MASK, r.p0, 0);
        const uint32_t blk_end = __reduce_max_sync(PW_FULL_MASK, r.live ? r.p0 + r.d : 0u);
        const uint32_t n_live = (uint32_t)__popc(live_mask) / L;
        if (lane == 0) {
            const uint32_t bar = pwasync::smem_u32(&bars[warp * NS + stage]);
            const uint32_t dst = pwasync::smem_u32(my_stage + (size_t)stage * ST::BYTES);
            const uint32_t nslots = blk_end - blk_start;
            const uint32_t o0 = blk_start & ~3u, o1 = (blk_end + 3u) & ~3u, t0 = blk_start & ~15u, t1 = (blk_end + 15u) & ~15u;
            const uint32_t b_in = nslots * K * (uint32_t)sizeof(T), b_un = n_live * K * (uint32_t)sizeof(T);
            const uint32_t b_op = nslots ? (o1 - o0) * 4u : 0u, b_ts = nslots ? (t1 - t0) : 0u;
            pwasync::mbar_expect_tx(bar, b_in + b_un + b_op + b_ts);
            if (b_in) pwasync::bulk_g2s(dst, (const T*)g.m2f_cur + (size_t)blk_start * K, b_in, bar);
            pwasync::bulk_g2s(dst + ST::OFF_UNARY, (const T*)g.unary + (size_t)(bin.base + rec_first) * K, b_un, bar);
            if (b_op) pwasync::bulk_g2s(dst + ST::OFF_OPP, g.opp + o0, b_op, bar);
            if (b_ts) pwasync::bulk_g2s(dst + ST::OFF_TSEL, g.tsel + t0, b_ts, bar);
        }
    };

    // records are fetched four iterations ahead (registers), the ranges of three iterations are in flight in the stages
    PwRec cur = load_rec(bin, first), r1 = load_rec(bin, first + stride), r2 = load_rec(bin, first + 2 * stride),
          r3 = load_rec(bin, first + 3 * stride);
    issue(cur, first - grp, 0);
    issue(r1, first - grp + stride, 1 % NS);
    issue(r2, first - grp + 2 * stride, 2 % NS);
    uint32_t it = 0;
    for (uint32_t base = blockIdx.x * per_cta; base < bin.n; base += stride, ++it) {
        const uint32_t gid = first + it * stride;
        const PwRec r4 = load_rec(bin, gid + 4 * stride);
        const bool live = cur.live;
        const uint32_t v = cur.v, p0 = cur.p0, d = cur.d;
        const int stage = (int)(it % NS);
        if (__any_sync(PW_FULL_MASK, live)) {
            const uint32_t blk_start = __shfl_sync(PW_FULL_MASK, p0, 0);
            pwasync::mbar_wait(pwasync::smem_u32(&bars[warp * NS + stage]), (it / NS) & 1);
            const unsigned char* st = my_stage + (size_t)stage * ST::BYTES;
            const T* sm_in = reinterpret_cast<const T*>(st);
            const T* sm_un = reinterpret_cast<const T*>(st + ST::OFF_UNARY);
            const uint32_t* sm_op = reinterpret_cast<const uint32_t*>(st + ST::OFF_OPP);
            const uint8_t* sm_ts = st + ST::OFF_TSEL;
            const uint32_t rel = p0 - blk_start, o_off = blk_start & 3u, t_off = blk_start & 15u;
            T un[S];
            vset<T, S>(un, T(1));
            if (live) ld_vec_plain<T, S>(sm_un + (size_t)grp * K + a0, un);
            uint32_t op[D];
            int sel[D];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                op[k] = 0;
                sel[k] = 0;
                if ((uint32_t)k < d) {
                    op[k] = sm_op[rel + k + o_off];
                    sel[k] = sm_ts[rel + k + t_off];
                }
            }
            T x[D][S];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                vset<T, S>(x[k], T(1));
                const bool have = (uint32_t)k < d;
                if (!__any_sync(PW_FULL_MASK, have)) continue;
                T in[K];
#pragma unroll
                for (int i = 0; i < K; ++i) in[i] = T(1);
                if (have) ld_vec_plain<T, K>(sm_in + (size_t)(rel + k) * K, in);
                T m[S];
                contract<T, K>(my_tab + sel[k] * tab_sel, in, m);
                norm1<T, K>(m);
                if (have) {
                    st_vec<true, T, S>((T*)g.m2v + (size_t)(p0 + k) * K + a0, m);
                    vcopy<T, S>(x[k], m);
                }
            }
            {
                T acc[S];
                vcopy<T, S>(acc, un);
#pragma unroll
                for (int k = 0; k < D; ++k) vmul<T, S>(acc, x[k]);
                norm1<T, K>(acc);
                if (live) st_vec<true, T, S>((T*)g.marg + (size_t)v * K + a0, acc);
            }
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const bool have = (uint32_t)k < d;
                if (!__any_sync(PW_FULL_MASK, have)) continue;
                T o[S];
                vcopy<T, S>(o, un);
#pragma unroll
                for (int j = 0; j < D; ++j)
                    if (j != k) vmul<T, S>(o, x[j]);
                norm1<T, K>(o);
                if (have) st_vec<true, T, S>((T*)g.m2f_nxt + (size_t)op[k] * K + a0, o);  // into the receiver's inbox
            }
        }
        __syncwarp();  // every lane has finished reading this stage before the engine overwrites it
        issue(r3, gid - grp + 3 * stride, stage);
        cur = r1;
        r1 = r2;
        r2 = r3;
        r3 = r4;
    }
}

// ---- teams: G lane groups cooperate on one variable (FULL) or on one 8G-slot chunk of a hub (H1 / H3) -------------------
// Each group keeps its segment of <= 8 messages in registers; segment products are combined by a shuffle scan over the
// groups of the team; exclusive products inside the segment by a prefix / suffix pass (the reference uses the segment
// tree of src/dependencies.jl:90-173 there: same values up to rounding, no tree nodes materialised).
//   FULL : m2v from the gathered m2f, marginal, m2f               (variables with 5 .. 8G factors)
//   H1   : m2v from the gathered m2f, chunk product -> chunk_prod  (hub chunks, pass 1)
//   H3   : m2v re-read, chunk_pre / chunk_suf from k_pw_hub_scan, m2f (hub chunks, pass 3)
enum { PW_MODE_FULL = 0, PW_MODE_H1 = 1, PW_MODE_H3 = 2 };
template <class T, int K, int G, int MODE>
__global__ void __launch_bounds__(256) k_pw_team(PwView g, PwBin bin) {
    constexpr int S = PwGeo<K>::S, L = PwGeo<K>::L, TL = G * L, SEG = PW_SEG;
    static_assert(TL <= 32, "a team lives inside one warp");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sh_tables = reinterpret_cast<T*>(smem_raw);
    if (MODE != PW_MODE_H3) stage_tables<T>(sh_tables, g, K);
    const T* my_tab = lane_tables<T>(sh_tables, g);
    const int tab_sel = g.tab_sel;
    const int tl = threadIdx.x % TL, grp = tl / L, a0 = (tl % L) * S;
    const uint32_t per_cta = blockDim.x / TL, stride = gridDim.x * per_cta, li = tl % L;
    const uint32_t first = blockIdx.x * per_cta + threadIdx.x / TL;
    PwRec cur = load_rec(bin, first), nxt = load_rec(bin, first + stride);
    for (uint32_t base = blockIdx.x * per_cta; base < bin.n; base += stride) {
    const uint32_t team = first + (base - blockIdx.x * per_cta);
    const PwRec nn = load_rec(bin, team + 2 * stride);
    if (g.prefetch && nxt.live) {  // pull the next iteration's streams into L2 (lane 0 of the group: messages, last lane: indices)
        const uint32_t nseg = (nxt.d + G - 1) / G, nlo = min(nxt.d, (uint32_t)grp * nseg), nn_loc = min(nxt.d, nlo + nseg) - nlo;
        const uint32_t npb = nxt.p0 + nlo;
        if (li == 0) {
            if (MODE == PW_MODE_FULL && grp == 0) prefetch_l2((const T*)g.unary + (size_t)(bin.base + team + stride) * K, K * sizeof(T));
            if (MODE != PW_MODE_H3)
                prefetch_l2((const T*)g.m2f_cur + (size_t)npb * K, (size_t)nn_loc * K * sizeof(T));
            else
                prefetch_l2((const T*)g.m2v + (size_t)npb * K, (size_t)nn_loc * K * sizeof(T));
        }
        if (li == L - 1) {
            if (MODE != PW_MODE_H1) prefetch_l2(g.opp + npb, (size_t)nn_loc * sizeof(uint32_t));
            if (MODE != PW_MODE_H3) prefetch_l2(g.tsel + npb, (size_t)nn_loc);
        }
    }
    const bool live = cur.live;
    const uint32_t v = cur.v, p0 = cur.p0, d = cur.d;
    cur = nxt;
    nxt = nn;
    const uint32_t seg = (d + G - 1) / G;  // <= SEG by construction of the bins
    const uint32_t lo = min(d, (uint32_t)grp * seg), n_loc = min(d, lo + seg) - lo;
    const uint32_t pb = p0 + lo;
    T x[SEG][S];
    uint32_t op[SEG];  // the receiver's slot of each outgoing m2f (not needed by H1)
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
        op[k] = 0;
        if (MODE != PW_MODE_H1 && (uint32_t)k < n_loc) op[k] = __ldg(g.opp + pb + k);
    }
    if (MODE != PW_MODE_H3) {
        int sel[SEG];
#pragma unroll
        for (int k = 0; k < SEG; ++k) {
            sel[k] = 0;
            if ((uint32_t)k < n_loc) sel[k] = __ldg(g.tsel + pb + k);
        }
#pragma unroll
        for (int k = 0; k < SEG; ++k) vset<T, S>(x[k], T(1));
        constexpr int HB = 4;  // incoming messages requested together (one round trip per batch)
#pragma unroll
        for (int h = 0; h < SEG; h += HB) {
            if (!__any_sync(PW_FULL_MASK, (uint32_t)h < n_loc)) continue;
            T in[HB][K];
#pragma unroll
            for (int kk = 0; kk < HB; ++kk) {
#pragma unroll
                for (int i = 0; i < K; ++i) in[kk][i] = T(1);
                if ((uint32_t)(h + kk) < n_loc) ld_vec<T, K>((const T*)g.m2f_cur + (size_t)(pb + h + kk) * K, in[kk]);
            }
#pragma unroll
            for (int kk = 0; kk < HB; ++kk) {
                const int k = h + kk;
                const bool have = (uint32_t)k < n_loc;
                T m[S];
                contract<T, K>(my_tab + sel[k] * tab_sel, in[kk], m);
                norm1<T, K>(m);
                if (have) {
                    st_vec<MODE == PW_MODE_FULL, T, S>((T*)g.m2v + (size_t)(pb + k) * K + a0, m);
                    vcopy<T, S>(x[k], m);
                }
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < SEG; ++k) {
            vset<T, S>(x[k], T(1));
            if ((uint32_t)k < n_loc) ld_vec_plain<T, S>((const T*)g.m2v + (size_t)(pb + k) * K + a0, x[k]);
        }
    }
    // segment product
    T P[S];
    vcopy<T, S>(P, x[0]);
#pragma unroll
    for (int k = 1; k < SEG; ++k) {
        vmul<T, S>(P, x[k]);
        if (k & 1) rescale<T, K>(P);
    }
    // inclusive prefix over the groups of the team, exclusive = the previous group's inclusive
    T inc[S], exc[S];
    vcopy<T, S>(inc, P);
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        T y[S];
#pragma unroll
        for (int i = 0; i < S; ++i) y[i] = __shfl_up_sync(PW_FULL_MASK, inc[i], off * L, TL);
        if (grp >= off) vmul<T, S>(inc, y);
        rescale<T, K>(inc);
    }
#pragma unroll
    for (int i = 0; i < S; ++i) {
        T y = __shfl_up_sync(PW_FULL_MASK, inc[i], L, TL);
        exc[i] = (G > 1 && grp > 0) ? y : T(1);
    }
    if (MODE == PW_MODE_H1) {
        if (live && grp == G - 1) st_vec<false, T, S>((T*)g.chunk_prod + (size_t)team * K + a0, inc);
        continue;
    }
    // exclusive suffix over the groups
    T sinc[S], sexc[S];
    vcopy<T, S>(sinc, P);
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        T y[S];
#pragma unroll
        for (int i = 0; i < S; ++i) y[i] = __shfl_down_sync(PW_FULL_MASK, sinc[i], off * L, TL);
        if (grp + off < G) vmul<T, S>(sinc, y);
        rescale<T, K>(sinc);
    }
#pragma unroll
    for (int i = 0; i < S; ++i) {
        T y = __shfl_down_sync(PW_FULL_MASK, sinc[i], L, TL);
        sexc[i] = (G > 1 && grp < G - 1) ? y : T(1);
    }
    T ext_pre[S], ext_suf[S];  // what lies outside the team: unary (FULL) or the other chunks of the hub (H3)
    vset<T, S>(ext_pre, T(1));
    vset<T, S>(ext_suf, T(1));
    if (live) {
        if (MODE == PW_MODE_FULL) {
            ld_vec<T, S>((const T*)g.unary + (size_t)(bin.base + team) * K + a0, ext_pre);
        } else {
            ld_vec_plain<T, S>((const T*)g.chunk_pre + (size_t)team * K + a0, ext_pre);
            ld_vec_plain<T, S>((const T*)g.chunk_suf + (size_t)team * K + a0, ext_suf);
        }
    }
    if (MODE == PW_MODE_FULL) {  // marginal = unary * every segment product (held by the last group)
        T mg[S];
        vcopy<T, S>(mg, ext_pre);
        vmul<T, S>(mg, inc);
        norm1<T, K>(mg);
        if (live && grp == G - 1) st_vec<true, T, S>((T*)g.marg + (size_t)v * K + a0, mg);
    }
    T run[S], suf[S];
    vcopy<T, S>(run, ext_pre);
    vmul<T, S>(run, exc);
    rescale<T, K>(run);
    vcopy<T, S>(suf, ext_suf);
    vmul<T, S>(suf, sexc);
    rescale<T, K>(suf);
    T pre[SEG][S];
#pragma unroll
    for (int k = 0; k < SEG; ++k) {
        vcopy<T, S>(pre[k], run);
        vmul<T, S>(run, x[k]);
        if (k & 1) rescale<T, K>(run);
    }
#pragma unroll
    for (int k = SEG - 1; k >= 0; --k) {
        const bool have = (uint32_t)k < n_loc;
        if (!__any_sync(PW_FULL_MASK, have)) continue;
        T o[S];
        vcopy<T, S>(o, pre[k]);
        vmul<T, S>(o, suf);
        norm1<T, K>(o);
        if (have) st_vec<true, T, S>((T*)g.m2f_nxt + (size_t)op[k] * K + a0, o);  // into the receiver's inbox
        vmul<T, S>(suf, x[k]);
        if (k & 1) rescale<T, K>(suf);
    }
    }  // grid-stride loop
}

// ---- hub pass 2: one lane group per hub scans the products of its chunks (prefix includes the unary), writes the marginal
template <class T, int K>
__global__ void __launch_bounds__(128) k_pw_hub_scan(PwView g, PwBin hubs /* v, first chunk, number of chunks */) {
    constexpr int S = PwGeo<K>::S, L = PwGeo<K>::L, U = 8;
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / L;
    const int a0 = (threadIdx.x % L) * S;
    const bool live = gid < hubs.n;
    uint32_t v = 0, c0 = 0, nc = 0;
    if (live) {
        v = __ldg(hubs.v + gid);
        c0 = __ldg(hubs.p0 + gid);
        nc = __ldg(hubs.d + gid);
    }
    uint32_t ncmax = nc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ncmax = max(ncmax, __shfl_xor_sync(PW_FULL_MASK, ncmax, o));
    T run[S];
    vset<T, S>(run, T(1));
    if (live) ld_vec<T, S>((const T*)g.unary + (size_t)(hubs.base + gid) * K + a0, run);
    for (uint32_t cb = 0; cb < ncmax; cb += U) {  // U independent loads in flight, then the dependent product chain
        T cp[U][S];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            vset<T, S>(cp[u], T(1));
            if (cb + u < nc) ld_vec_plain<T, S>((const T*)g.chunk_prod + (size_t)(c0 + cb + u) * K + a0, cp[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (cb + u < nc) st_vec<false, T, S>((T*)g.chunk_pre + (size_t)(c0 + cb + u) * K + a0, run);
            vmul<T, S>(run, cp[u]);
            rescale<T, K>(run);
        }
    }
    norm1<T, K>(run);
    if (live) st_vec<true, T, S>((T*)g.marg + (size_t)v * K + a0, run);
    T suf[S];
    vset<T, S>(suf, T(1));
    const uint32_t nb = (ncmax + U - 1) / U;
    for (uint32_t b = nb; b > 0; --b) {
        const uint32_t cb = (b - 1) * U;
        T cp[U][S];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            vset<T, S>(cp[u], T(1));
            if (cb + u < nc) ld_vec_plain<T, S>((const T*)g.chunk_prod + (size_t)(c0 + cb + u) * K + a0, cp[u]);
        }
#pragma unroll
        for (int u = U - 1; u >= 0; --u) {
            if (cb + u < nc) st_vec<false, T, S>((T*)g.chunk_suf + (size_t)(c0 + cb + u) * K + a0, suf);
            vmul<T, S>(suf, cp[u]);
            rescale<T, K>(suf);
        }
    }
}

// unary evidence from variable order into record (processing) order: dst[rec_of_var[v]] = src[v]
template <class T>
__global__ void k_pw_permute_rows(const T* src, const uint32_t* rec_of_var, T* dst, size_t n, int K) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * K) dst[(size_t)rec_of_var[i / K] * K + (i % K)] = src[i];
}
template <class T>
__global__ void k_pw_fill(T* p, size_t n, T v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// gather per-slot planes into factor order: out[2f + side] = plane[slot_of_edge[2f + side]]
template <class T>
__global__ void k_pw_gather(const T* plane, const uint32_t* slot_of_edge, T* out, size_t n_edges, int K, int flip) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_edges * K) out[i] = plane[(size_t)slot_of_edge[(i / K) ^ (size_t)flip] * K + (i % K)];
}

struct Pairwise {
    int device = 0, dtype = CXB_F32, K = 0, n_tables = 0, cur = 0;
    long long n = 0, m = 0;
    cudaStream_t stream = nullptr;
    // the bins of one sweep touch disjoint variables: they are launched on side streams (fork / join with events) so that
    // the light kernels and the three dependent hub launches fill the SMs the heavy ones leave idle
    static constexpr int N_AUX = 3;
    cudaStream_t aux[N_AUX] = {nullptr, nullptr, nullptr}, ls = nullptr;  // ls = stream of the launch helpers
    cudaEvent_t ev_fork = nullptr, ev_join[N_AUX] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    // work bins: 0 = exact path (<= 4 factors), 1 + log2(G) = teams of G groups, then hub chunks and hubs
    static constexpr int N_TEAM_BINS = 5;  // G = 1, 2, 4, 8, 16
    struct Bin {
        DBuf<uint32_t> v, p0, d;
        uint32_t n = 0, base = 0;  // base = global record index of the bin's first record (unary_rec is in record order)
        PwBin view() const { return PwBin{v.p, p0.p, d.p, n, base}; }
    };
    Bin exact_bin, team_bin[N_TEAM_BINS], chunk_bin, hub_bin;
    DBuf<uint32_t> opp, slot_of_edge, rec_of_var_d;
    DBuf<uint8_t> tsel;
    DBuf<unsigned char> tables, unary, unary_rec, m2f[2], m2v, marg, scratch, chunk_prod, chunk_pre, chunk_suf;
    int g_max = 16;  // largest team (groups) that fits one warp at this K
    int n_sm = 148;
    long long n_products = 0;
    bool have_graph = false, have_tables = false, have_unary = false, have_msgs = false, ran = false;
    size_t esz() const { return dtype == CXB_F32 ? 4 : 8; }
    ~Pairwise() {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (int i = 0; i < N_AUX; ++i) {
            if (ev_join[i]) cudaEventDestroy(ev_join[i]);
            if (aux[i]) cudaStreamDestroy(aux[i]);
        }
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
        if (const char* e = getenv("CXB_L2_FETCH")) {  // tuning knob: L2 -> DRAM fetch granularity hint (32 / 64 / 128 bytes)
            size_t before = 0, after = 0;
            cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
            cudaError_t le = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
            cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
            fprintf(stderr, "cxb_pairwise: L2 fetch granularity %zu -> %zu (%s)\n", before, after, cudaGetErrorString(le));
        }
        CXB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CXB_CUDA(cudaEventCreate(&ev0));
        CXB_CUDA(cudaEventCreate(&ev1));
        CXB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
        CXB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        for (int i = 0; i < N_AUX; ++i) {
            CXB_CUDA(cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking));
            CXB_CUDA(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
        }
        ls = stream;
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
        // degree bins:
        //   <= 4 factors: exact path; <= 8G: team of G groups, G = 1 .. g_max; larger: hub, cut into chunks of 8 g_max slots
        const int lanes = K / std::min(K, 4);
        g_max = std::min(16, 32 / lanes);
        const uint32_t chunk_slots = (uint32_t)PW_SEG * (uint32_t)g_max;
        struct Rec {
            uint32_t v, p0, d;
        };
        std::vector<Rec> ex, tm[N_TEAM_BINS], hubs, chunks;
        n_products = 0;
        for (long long v = 0; v < n; ++v) {
            uint32_t d = off[v + 1] - off[v];
            Rec r{(uint32_t)v, off[v], d};
            if (d <= (uint32_t)PW_EXACT_MAX) {
                ex.push_back(r);
                continue;
            }
            n_products += (long long)d - 1;  // (d+1) factors incl. the unary -> d-1 ProductOfMessages nodes in the reference
            if (d > chunk_slots) {
                hubs.push_back(r);
                continue;
            }
            int b = 0;
            while ((uint32_t)PW_SEG << b < d) ++b;
            tm[b].push_back(r);
        }
        auto by_degree_desc = [](const Rec& x, const Rec& y) { return x.d > y.d; };
        // Every bin is sorted by degree (ascending id inside a degree): the lanes of a warp run the same trip counts. The
        // locality this would cost is restored by RENUMBERING the slots in processing order: the slots of the records of a
        // bin are consecutive in the order the kernel visits them, so the inbox, opp, tsel and m2v of a bin are pure
        // streams (a warp touches one contiguous block), and the unary evidence is kept in a second copy in record order.
        std::stable_sort(ex.begin(), ex.end(), by_degree_desc);
        for (auto& t : tm) std::stable_sort(t.begin(), t.end(), by_degree_desc);
        std::stable_sort(hubs.begin(), hubs.end(), by_degree_desc);
        std::vector<uint32_t> newoff((size_t)n, 0), rec_of_var((size_t)n, 0);
        {
            uint32_t run = 0, rec = 0;
            auto number = [&](std::vector<Rec>& recs) {
                for (Rec& r : recs) {
                    newoff[r.v] = run;
                    r.p0 = run;
                    run += r.d;
                    rec_of_var[r.v] = rec++;
                }
            };
            exact_bin.base = rec;
            number(ex);
            for (int b2 = 0; b2 < N_TEAM_BINS; ++b2) {
                team_bin[b2].base = rec;
                number(tm[b2]);
            }
            hub_bin.base = rec;
            number(hubs);
        }
        std::vector<uint32_t> cursor(newoff), slot(P), oppv(P);
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
        std::vector<Rec> hub_recs;  // v, first chunk, number of chunks
        for (const Rec& h : hubs) {
            uint32_t nc = (h.d + chunk_slots - 1) / chunk_slots;
            hub_recs.push_back(Rec{h.v, (uint32_t)chunks.size(), nc});
            for (uint32_t c = 0; c < nc; ++c)
                chunks.push_back(Rec{h.v, h.p0 + c * chunk_slots, std::min(chunk_slots, h.d - c * chunk_slots)});
        }
        auto up = [&](auto& dbuf, const auto& vec) -> cudaError_t {
            cudaError_t e = dbuf.reserve(vec.size());
            if (e != cudaSuccess) return e;
            return vec.empty() ? cudaSuccess
                               : cudaMemcpyAsync(dbuf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice, stream);
        };
        std::vector<uint32_t> tmp_v, tmp_p, tmp_d;
        auto up_bin = [&](Bin& bin, const std::vector<Rec>& recs) -> cudaError_t {
            tmp_v.resize(recs.size());
            tmp_p.resize(recs.size());
            tmp_d.resize(recs.size());
            for (size_t i = 0; i < recs.size(); ++i) {
                tmp_v[i] = recs[i].v;
                tmp_p[i] = recs[i].p0;
                tmp_d[i] = recs[i].d;
            }
            bin.n = (uint32_t)recs.size();
            cudaError_t e = up(bin.v, tmp_v);
            if (e == cudaSuccess) e = up(bin.p0, tmp_p);
            if (e == cudaSuccess) e = up(bin.d, tmp_d);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // tmp_* are reused
            return e;
        };
        CXB_CUDA(up_bin(exact_bin, ex));
        for (int b = 0; b < N_TEAM_BINS; ++b) CXB_CUDA(up_bin(team_bin[b], tm[b]));
        CXB_CUDA(up_bin(chunk_bin, chunks));
        CXB_CUDA(up_bin(hub_bin, hub_recs));
        CXB_CUDA(up(rec_of_var_d, rec_of_var));
        oppv.resize(P + 16, 0);  // the bulk copies fetch 16-byte aligned windows: a few entries past the end may be read
        sel.resize(P + 32, 0);
        CXB_CUDA(up(opp, oppv));
        CXB_CUDA(up(tsel, sel));
        CXB_CUDA(up(slot_of_edge, slot));
        size_t cb = std::max<size_t>(chunks.size(), 1) * K * esz();
        CXB_CUDA(chunk_prod.reserve(cb));
        CXB_CUDA(chunk_pre.reserve(cb));
        CXB_CUDA(chunk_suf.reserve(cb));
        size_t pb = std::max<size_t>(P, 1) * K * esz(), nb = (size_t)n * K * esz();
        CXB_CUDA(m2f[0].reserve(pb));
        CXB_CUDA(m2f[1].reserve(pb));
        CXB_CUDA(m2v.reserve(pb));
        CXB_CUDA(scratch.reserve(pb));
        CXB_CUDA(marg.reserve(nb));
        CXB_CUDA(unary.reserve(nb));
        CXB_CUDA(unary_rec.reserve(nb));
        CXB_CUDA(cudaMemsetAsync(m2v.p, 0, pb, stream));
        CXB_CUDA(cudaMemsetAsync(marg.p, 0, nb, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        have_graph = true;
        return CXB_OK;
    }
    // Shared-memory image of the tables. Selection sel = 2 * table + side: side 0 (this variable is the lower endpoint of
    // the factor) needs rows of psi, side 1 rows of its transpose: row a of selection sel, column b = weight of in[b] for
    // output state a. A lane owns S = min(K, 4) consecutive states, the L = K / S lanes of a group own different rows.
    // Block j (j = lane % 8) holds rows [(j % L) S, (j % L) S + S) of every selection, selections tab_sel elements apart
    // (a multiple of 128 bytes); blocks are tab_block = plane + 16 bytes apart, so block j starts 4 j banks round.
    int tab_elems = 0, tab_block = 0, tab_blocks = 0, tab_sel = 0;
    int32_t set_tables(const double* tb) {
        if (!have_graph) {
            err = "set the graph first";
            return CXB_ERR_STATE;
        }
        CXB_CUDA(cudaSetDevice(device));
        const int S = std::min(K, 4), L = K / S, QE = 16 / (int)esz(), n_sel = 2 * n_tables;
        tab_sel = (S * K + 8 * QE - 1) / (8 * QE) * (8 * QE);
        tab_block = n_sel * tab_sel + QE;
        tab_blocks = (size_t)8 * tab_block * esz() <= 96 * 1024 ? 8 : L;
        tab_elems = tab_blocks * tab_block;
        if ((size_t)tab_elems * esz() > 200 * 1024) {
            err = "tables do not fit in shared memory";
            return CXB_ERR_BAD_ARG;
        }
        std::vector<unsigned char> raw((size_t)tab_elems * esz(), 0);
        for (int j = 0; j < tab_blocks; ++j)
            for (int sel = 0; sel < n_sel; ++sel)
                for (int r = 0; r < S; ++r)
                    for (int bcol = 0; bcol < K; ++bcol) {
                        const int t = sel >> 1, a = (j % L) * S + r;
                        const double v = (sel & 1) ? tb[((size_t)t * K + bcol) * K + a] : tb[((size_t)t * K + a) * K + bcol];
                        const size_t o = (size_t)j * tab_block + (size_t)sel * tab_sel + (size_t)r * K + bcol;
                        if (dtype == CXB_F32)
                            ((float*)raw.data())[o] = (float)v;
                        else
                            ((double*)raw.data())[o] = v;
                    }
        CXB_CUDA(tables.reserve(raw.size()));
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
    // persistent CTAs of 256 threads: exactly as many as are resident at once (a CTA stages the table image once, ~33 KB)
    template <class Kern>
    unsigned grid_for(Kern kern, size_t threads, size_t smem) const {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        return std::min<unsigned>(cdiv(threads, 256), (unsigned)n_sm * (unsigned)per_sm);
    }
    template <class T, int KK, int G>
    void launch_team(const PwView& g, size_t tb) {
        constexpr int TL = G * PwGeo<KK>::L;
        if constexpr (TL <= 32) {
            const Bin& bin = team_bin[G == 1 ? 0 : G == 2 ? 1 : G == 4 ? 2 : G == 8 ? 3 : 4];
            if (!bin.n) return;
            if (tb > 48 * 1024)
                cudaFuncSetAttribute(k_pw_team<T, KK, G, PW_MODE_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb);
            CXB_LAUNCH((k_pw_team<T, KK, G, PW_MODE_FULL>), grid_for(k_pw_team<T, KK, G, PW_MODE_FULL>, (size_t)bin.n * TL, tb), 256, tb, ls, g, bin.view());
        }
    }
    template <class T, int KK, int G>
    void launch_hubs(const PwView& g, size_t tb) {
        constexpr int L = PwGeo<KK>::L, TL = G * L;
        if constexpr (TL <= 32) {
            if (G != g_max || !chunk_bin.n) return;
            if (tb > 48 * 1024)
                cudaFuncSetAttribute(k_pw_team<T, KK, G, PW_MODE_H1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb);
            CXB_LAUNCH((k_pw_team<T, KK, G, PW_MODE_H1>), grid_for(k_pw_team<T, KK, G, PW_MODE_H1>, (size_t)chunk_bin.n * TL, tb), 256, tb, ls, g, chunk_bin.view());
            CXB_LAUNCH((k_pw_hub_scan<T, KK>), cdiv((size_t)hub_bin.n * L, 128), 128, 0, ls, g, hub_bin.view());
            CXB_LAUNCH((k_pw_team<T, KK, G, PW_MODE_H3>), grid_for(k_pw_team<T, KK, G, PW_MODE_H3>, (size_t)chunk_bin.n * TL, 0), 256, 0, ls, g, chunk_bin.view());
        }
    }
    template <class T, int KK>
    int32_t launch_k(const PwView& g) {
        size_t tb = (size_t)tab_elems * sizeof(T);
        const bool multi = !(getenv("CXB_PW_ONE_STREAM") && atoi(getenv("CXB_PW_ONE_STREAM")));
        if (multi) {
            CXB_CUDA(cudaEventRecord(ev_fork, stream));
            for (int i = 0; i < N_AUX; ++i) CXB_CUDA(cudaStreamWaitEvent(aux[i], ev_fork, 0));
        }
        // side stream 0: hubs (three dependent launches) and the large teams; 1: teams of 4 and 2; 2: teams of 1; main: exact
        ls = multi ? aux[0] : stream;
        launch_hubs<T, KK, 16>(g, tb);
        launch_hubs<T, KK, 8>(g, tb);
        launch_hubs<T, KK, 4>(g, tb);
        launch_team<T, KK, 16>(g, tb);
        launch_team<T, KK, 8>(g, tb);
        ls = multi ? aux[1] : stream;
        launch_team<T, KK, 4>(g, tb);
        launch_team<T, KK, 2>(g, tb);
        ls = multi ? aux[2] : stream;
        launch_team<T, KK, 1>(g, tb);
        ls = stream;
        const bool staged = KK * sizeof(T) >= 16 && !(getenv("CXB_PW_STAGED") && !atoi(getenv("CXB_PW_STAGED")));
        if (exact_bin.n && staged) {
            using ST = PwStage<T, KK>;
            const size_t smem = (tb + 127) / 128 * 128 + (size_t)8 * PW_STAGES * ST::BYTES + 8 * PW_STAGES * sizeof(uint64_t);
            cudaFuncSetAttribute(k_pw_exact_staged<T, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            CXB_LAUNCH((k_pw_exact_staged<T, KK>), grid_for(k_pw_exact_staged<T, KK>, (size_t)exact_bin.n * PwGeo<KK>::L, smem), 256, smem, ls, g, exact_bin.view());
        } else if (exact_bin.n) {
            int minb = 2;
            if (const char* e = getenv("CXB_PW_MINB")) minb = atoi(e);
            if (minb >= 3) {
                if (tb > 48 * 1024) cudaFuncSetAttribute(k_pw_exact<T, KK, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb);
                CXB_LAUNCH((k_pw_exact<T, KK, 3>), grid_for(k_pw_exact<T, KK, 3>, (size_t)exact_bin.n * PwGeo<KK>::L, tb), 256, tb, ls, g, exact_bin.view());
            } else {
                if (tb > 48 * 1024) cudaFuncSetAttribute(k_pw_exact<T, KK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb);
                CXB_LAUNCH((k_pw_exact<T, KK, 2>), grid_for(k_pw_exact<T, KK, 2>, (size_t)exact_bin.n * PwGeo<KK>::L, tb), 256, tb, ls, g, exact_bin.view());
            }
        }
        if (multi)
            for (int i = 0; i < N_AUX; ++i) {
                CXB_CUDA(cudaEventRecord(ev_join[i], aux[i]));
                CXB_CUDA(cudaStreamWaitEvent(stream, ev_join[i], 0));
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
        g.opp = opp.p;
        g.tsel = tsel.p;
        g.tables = tables.p;
        g.tab_elems = tab_elems;
        g.tab_block = tab_block;
        g.tab_blocks = tab_blocks;
        g.tab_sel = tab_sel;
        g.prefetch = 1;
        if (const char* e = getenv("CXB_PW_PREFETCH")) g.prefetch = atoi(e);
        g.unary = unary_rec.p;  // record order
        g.m2f_cur = m2f[cur].p;
        g.m2f_nxt = m2f[cur ^ 1].p;
        g.m2v = m2v.p;
        g.marg = marg.p;
        g.chunk_prod = chunk_prod.p;
        g.chunk_pre = chunk_pre.p;
        g.chunk_suf = chunk_suf.p;
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
    int32_t get_edges(const unsigned char* plane, void* out_host, bool inbox) {
        CXB_CUDA(cudaSetDevice(device));
        size_t E = (size_t)2 * m;
        if (E) {
            const int flip = inbox ? 1 : 0;  // m2f(v, f) is stored in the inbox slot of the OTHER endpoint of f
            if (dtype == CXB_F32)
                CXB_LAUNCH(k_pw_gather<float>, cdiv(E * K, 256), 256, 0, stream, (const float*)plane, slot_of_edge.p, (float*)scratch.p, E, K,
                           flip);
            else
                CXB_LAUNCH(k_pw_gather<double>, cdiv(E * K, 256), 256, 0, stream, (const double*)plane, slot_of_edge.p, (double*)scratch.p,
                           E, K, flip);
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
                            int32_t n_tables, cxb_pairwise** out) try {
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
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
void cxb_pairwise_destroy(cxb_pairwise* g) {
    if (g) {
        cudaSetDevice(PW(g)->device);
        delete PW(g);
    }
}
const char* cxb_pairwise_last_error(cxb_pairwise* g) { return g ? PW(g)->err.c_str() : "null handle"; }
int32_t cxb_pairwise_set_graph(cxb_pairwise* g, const int64_t* fac_u, const int64_t* fac_v, const int32_t* fac_table) try {
    return PW(g)->set_graph(fac_u, fac_v, fac_table);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_pairwise_set_tables(cxb_pairwise* g, const double* tables) try { return PW(g)->set_tables(tables); } CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_pairwise_set_unary(cxb_pairwise* g, const void* unary_host) try {
    Pairwise* h = PW(g);
    if (!h->have_graph) {
        h->err = "set the graph first";
        return CXB_ERR_STATE;
    }
    PW_CUDA(g, cudaSetDevice(h->device));
    PW_CUDA(g, cudaMemcpyAsync(h->unary.p, unary_host, (size_t)h->n * h->K * h->esz(), cudaMemcpyHostToDevice, h->stream));
    const size_t cnt = (size_t)h->n * h->K;
    if (h->dtype == CXB_F32)
        CXB_LAUNCH(cxb::k_pw_permute_rows<float>, cxb::cdiv(cnt, 256), 256, 0, h->stream, (const float*)h->unary.p, h->rec_of_var_d.p,
                   (float*)h->unary_rec.p, (size_t)h->n, h->K);
    else
        CXB_LAUNCH(cxb::k_pw_permute_rows<double>, cxb::cdiv(cnt, 256), 256, 0, h->stream, (const double*)h->unary.p, h->rec_of_var_d.p,
                   (double*)h->unary_rec.p, (size_t)h->n, h->K);
    PW_CUDA(g, cudaStreamSynchronize(h->stream));
    h->have_unary = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_pairwise_reset_messages(cxb_pairwise* g) try { return PW(g)->reset(); } CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_pairwise_sweep(cxb_pairwise* g, int64_t* n_updates_out) try { return PW(g)->sweep(n_updates_out); } CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_pairwise_get_marginals(cxb_pairwise* g, void* out_host) try {
    Pairwise* h = PW(g);
    PW_CUDA(g, cudaSetDevice(h->device));
    PW_CUDA(g, cudaMemcpyAsync(out_host, h->marg.p, (size_t)h->n * h->K * h->esz(), cudaMemcpyDeviceToHost, h->stream));
    PW_CUDA(g, cudaStreamSynchronize(h->stream));
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_pairwise_get_messages(cxb_pairwise* g, int32_t which, void* out_host) try {
    Pairwise* h = PW(g);
    if (which != 0 && which != 1) {
        h->err = "which must be 0 (m2v) or 1 (m2f)";
        return CXB_ERR_BAD_ARG;
    }
    return h->get_edges(which == 0 ? h->m2v.p : h->m2f[h->cur].p, out_host, which == 1);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int64_t cxb_pairwise_algorithmic_bytes(cxb_pairwise* g) try {
    Pairwise* h = PW(g);
    // per variable of degree d: (d + 1) messages read, (2d + 1) written (no ProductOfMessages traffic)
    return (int64_t)((size_t)(2 * h->m + h->n) + (size_t)(4 * h->m + h->n)) * h->K * (int64_t)h->esz();
} CXB_ABI_CATCH(-1)
void* cxb_pairwise_stream(cxb_pairwise* g) { return (void*)PW(g)->stream; }
int32_t cxb_pairwise_last_kernel_ms(cxb_pairwise* g, float* ms_out) try {
    Pairwise* h = PW(g);
    if (!h->ran) {
        h->err = "no sweep has run yet";
        return CXB_ERR_STATE;
    }
    PW_CUDA(g, cudaEventSynchronize(h->ev1));
    PW_CUDA(g, cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_pairwise_sync(cxb_pairwise* g) try {
    PW_CUDA(g, cudaStreamSynchronize(PW(g)->stream));
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)

}  // extern "C"
