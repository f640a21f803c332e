// hmm.cu — structured engine for a batch of discrete HMMs (BASELINE config 3).
//
// Graph per chain: hidden z_t, observed y_t, emission factor (z_t, y_t), transition factor
// (z_t, z_{t+1}), uniform prior leaf on z_1; data enter as set_value!(m2f(y_t, em_t), o_t).  One call =
// update_marginals!(engine, z[1:T]) of every chain = 6T-4 message updates per chain (categorical sum-product,
// every message normalised to sum 1).  Materialised in HBM: the forward message m2f(z_t, tr_t) and the marginal
// ("forward message + marginal" contract of SURVEY §8d: 12K+2 bytes per (chain, step)).
//
// Kernel shape: the recursion over t is strictly sequential, the batch is small (1,024 chains), so the unit of
// parallelism is ONE WARP PER CHAIN with lane l owning states l, l+32, ...:
//   out[j] = sum_i Tbl[i][j] * v[i]        (Tbl = A forward, A^T backward)
// For K <= 64 (fp32) the lane's columns of Tbl live in REGISTERS (2 x 64 values), v is staged in a per-warp
// shared-memory line and read back with broadcast 128-bit loads, so a step is K*K/32 FFMAs + K/4 LDS per warp
// and no block-level barrier.  Larger K stream Tbl through shared memory in row tiles shared by the 8 chains of
// the CTA.  (The tcgen05 path for K = 512 named by the north star is future work: see DESIGN.md.)
#include <algorithm>
#include <cmath>
#include <type_traits>

#include <cstring>

#include "common.cuh"
#include "hmm_tc.cuh"
#include "hmm64_tc.cuh"

namespace cxb {

constexpr int HMM_WARPS = 8;  // chains per CTA

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One pass over time. FWD: t ascending, writes fwd[t] = m2f(z_t, tr_t) = normalise(em_t * pred_t).
// BWD: t descending, reads fwd[t], writes marg[t] = normalise(fwd[t] * bwd_t), carries
// m2f(z_t, tr_{t-1}) = normalise(em_t * bwd_t).  tbl is A (FWD) or A^T (BWD), row-major [K][K];
// emis_n is the column-normalised emission table transposed to [M][K] (m2v(z_t, em_t) = emis_n[o_t]), staged in
// shared memory.  Nothing on the per-step critical path waits on global memory: observations are fetched 32 steps at a
// time (one per lane, broadcast by shuffle, double buffered), the forward message of the next step is already in
// registers and the one 8 steps ahead is being pulled into L2 (prefetch.global.L2).
template <class T, int K, bool REGA, bool FWD>
__global__ void __launch_bounds__(HMM_WARPS * 32)
k_hmm_pass(const T* __restrict__ tbl, const T* __restrict__ emis_n, const uint8_t* __restrict__ obs, T* __restrict__ fwd,
           T* __restrict__ marg, long long B, long long Tn, int Kdyn, int n_sym, int tile_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Kk = REGA ? K : Kdyn;
    constexpr int CPL_MAX = REGA ? K / 32 : 32;  // states per lane (<= 1024 states in the streamed path)
    const int cpl = (Kk + 31) / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * HMM_WARPS + warp;
    const bool live = b < B;
    const long long bb = live ? b : 0;
    T* sh_v = reinterpret_cast<T*>(smem_raw) + (size_t)warp * Kk;            // per-warp staging of v
    T* sh_em = reinterpret_cast<T*>(smem_raw) + (size_t)HMM_WARPS * Kk;        // [M][K] emission messages
    T* sh_tbl = sh_em + (size_t)n_sym * Kk;                                    // streamed path: row tile of tbl

    T areg[REGA ? K : 1][REGA ? K / 32 : 1];
    if (REGA) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int c = 0; c < K / 32; ++c) areg[i][c] = tbl[(size_t)i * K + lane + 32 * c];
    }
    for (int x = threadIdx.x; x < n_sym * Kk; x += blockDim.x) sh_em[x] = emis_n[x];
    const bool whole_table = !REGA && tile_rows >= Kk;
    if (whole_table)
        for (int x = threadIdx.x; x < Kk * Kk; x += blockDim.x) sh_tbl[x] = tbl[x];
    __syncthreads();

    auto time_of = [&](long long step) { return FWD ? step : Tn - 1 - step; };
    auto load_obs_block = [&](long long step0) -> int {  // lane l fetches the symbol of step0 + l
        long long st = step0 + lane;
        return (st < Tn) ? (int)obs[(size_t)time_of(st) * B + bb] : 0;
    };
    int obs_cur = 0, obs_next = load_obs_block(0);

    T v[CPL_MAX];  // carried message (lane's states)
#pragma unroll
    for (int c = 0; c < CPL_MAX; ++c) v[c] = T(0);
    T a_next[CPL_MAX];  // BWD: forward message of the upcoming step
#pragma unroll
    for (int c = 0; c < CPL_MAX; ++c) {
        int j = lane + 32 * c;
        a_next[c] = (!FWD && c < cpl && j < Kk) ? __ldcs(&fwd[((size_t)time_of(0) * B + bb) * Kk + j]) : T(0);
    }

    for (long long step = 0; step < Tn; ++step) {
        const long long t = time_of(step);
        if ((step & 31) == 0) {
            obs_cur = obs_next;
            obs_next = load_obs_block(step + 32);
        }
        int o = __shfl_sync(0xffffffffu, obs_cur, (int)(step & 31));
        if (o >= n_sym) o = n_sym - 1;
        T a[CPL_MAX];
        if (!FWD) {
#pragma unroll
            for (int c = 0; c < CPL_MAX; ++c) a[c] = a_next[c];
            if (step + 1 < Tn) {
                const T* nx = &fwd[((size_t)time_of(step + 1) * B + bb) * Kk];
#pragma unroll
                for (int c = 0; c < CPL_MAX; ++c) {
                    int j = lane + 32 * c;
                    if (c < cpl && j < Kk) a_next[c] = __ldcs(nx + j);
                }
            }
            if (step + 8 < Tn) {
                const T* far = &fwd[((size_t)time_of(step + 8) * B + bb) * Kk];
                if (lane * 32 < Kk * (int)sizeof(T)) asm volatile("prefetch.global.L2 [%0];" ::"l"(far + lane * (32 / sizeof(T))));
            }
        }
        T out[CPL_MAX];
#pragma unroll
        for (int c = 0; c < CPL_MAX; ++c) out[c] = T(0);
        const bool has_prev = step > 0;
        if (has_prev) {
            // stage v, then out[j] = sum_i tbl[i][j] v[i]
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CPL_MAX; ++c)
                if (c < cpl && lane + 32 * c < Kk) sh_v[lane + 32 * c] = v[c];
            __syncwarp();
            if (REGA) {
                T acc[REGA ? K / 32 : 1][4];  // 4 independent FMA chains per column
#pragma unroll
                for (int c = 0; c < K / 32; ++c)
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[c][u] = T(0);
#pragma unroll
                for (int i = 0; i < K; i += 4) {
                    T vi[4];
                    if (sizeof(T) == 4) {
                        float4 q = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sh_v) + i);
                        vi[0] = q.x; vi[1] = q.y; vi[2] = q.z; vi[3] = q.w;
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u) vi[u] = sh_v[i + u];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int c = 0; c < K / 32; ++c) acc[c][u] = fma(areg[i + u][c], vi[u], acc[c][u]);
                }
#pragma unroll
                for (int c = 0; c < K / 32; ++c) out[c] = (acc[c][0] + acc[c][1]) + (acc[c][2] + acc[c][3]);
            } else if (whole_table) {
                for (int i = 0; i < Kk; ++i) {
                    T vi = sh_v[i];
#pragma unroll 4
                    for (int c = 0; c < cpl; ++c) {
                        int j = lane + 32 * c;
                        if (j < Kk) out[c] = fma(sh_tbl[(size_t)i * Kk + j], vi, out[c]);
                    }
                }
            } else {
                for (int r0 = 0; r0 < Kk; r0 += tile_rows) {
                    int nr = min(tile_rows, Kk - r0);
                    __syncthreads();
                    for (int x = threadIdx.x; x < nr * Kk; x += blockDim.x) sh_tbl[x] = tbl[(size_t)r0 * Kk + x];
                    __syncthreads();
                    for (int i = 0; i < nr; ++i) {
                        T vi = sh_v[r0 + i];
#pragma unroll 4
                        for (int c = 0; c < cpl; ++c) {
                            int j = lane + 32 * c;
                            if (j < Kk) out[c] = fma(sh_tbl[(size_t)i * Kk + j], vi, out[c]);
                        }
                    }
                }
            }
        }
        // emission message of this step (shared memory)
        T em[CPL_MAX];
#pragma unroll
        for (int c = 0; c < CPL_MAX; ++c) {
            int j = lane + 32 * c;
            em[c] = (c < cpl && j < Kk) ? sh_em[(size_t)o * Kk + j] : T(0);
        }
        const size_t base = ((size_t)t * B + bb) * Kk;
        if (FWD) {
            // m2f(z_t, tr_t) = normalise(em * pred); at t = 0 the uniform prior leaves normalise(em)
            T part = T(0);
#pragma unroll
            for (int c = 0; c < CPL_MAX; ++c) {
                v[c] = has_prev ? em[c] * out[c] : em[c];
                part += v[c];
            }
            T tot = warp_sum(part);
#pragma unroll
            for (int c = 0; c < CPL_MAX; ++c) {
                v[c] = v[c] / tot;
                int j = lane + 32 * c;
                if (live && c < cpl && j < Kk) __stcs(&fwd[base + j], v[c]);
            }
        } else {
            T part = T(0), part2 = T(0);
#pragma unroll
            for (int c = 0; c < CPL_MAX; ++c) {
                T g = has_prev ? a[c] * out[c] : a[c];    // marginal = fwd * bwd
                T m = has_prev ? em[c] * out[c] : em[c];  // m2f(z_t, tr_{t-1}) = em * bwd
                a[c] = g;
                v[c] = m;
                part += g;
                part2 += m;
            }
            T tot = warp_sum(part), tot2 = warp_sum(part2);
#pragma unroll
            for (int c = 0; c < CPL_MAX; ++c) {
                int j = lane + 32 * c;
                v[c] = v[c] / tot2;
                if (live && c < cpl && j < Kk) __stcs(&marg[base + j], a[c] / tot);
            }
        }
    }
}

// ---- K = 64, fp32: one warp per chain, one warp per CTA (1,024 CTAs = 6.9 per SM for config 3) ------------------------------
// The 64 x 64 table lives in registers: lane (jg = lane / 2, ig = lane % 2) owns the 4 x 32 tile
// tbl[32 ig .. 32 ig + 31][4 jg .. 4 jg + 3] as 64 packed pairs and does 64 FFMA2 (fma.rn.f32x2) per step; the carried
// message is read back with eight conflict-free 128-bit shared loads; the two i-halves are combined by a 2-shuffle
// transpose-reduce that leaves output states (2 lane, 2 lane + 1) in the lane, so messages move as one float2 per lane
// (256 contiguous bytes per warp).  Nothing but the matvec is on the per-step critical path: the carried message is NOT
// normalised exactly, it is scaled by the power of two 2^-floor(log2 sum(u_{s-1})) (exact, keeps sum(u_s) within
// [n_s, 2 n_s), n_s = the true normaliser), and the exactly normalised message / marginal of step s-1
// (u_{s-1} / sum(u_{s-1})) is written one step late, while the matvec of step s is in flight.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
// branch-free reciprocal (MUFU.RCP + one Newton step, <= 1 ulp for the normal, positive sums it is used on): __frcp_rn
// carries a slow-path call that would split the step into basic blocks and serialise the shuffle chain with the matvec
__device__ __forceinline__ float hmm_rcp(float x) {
    float q;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(x));
    return fmaf(q, fmaf(-x, q, 1.0f), q);
}
__device__ __forceinline__ float hmm_pow2_inv(float s) {
    unsigned e = (__float_as_uint(s) >> 23) & 0xffu;
    return __uint_as_float((254u - e) << 23);
}
template <bool FWD, bool EM_SMEM>
__global__ void __launch_bounds__(256)
k_hmm64_pass(const float* __restrict__ tbl, const float* __restrict__ emis_n, const uint8_t* __restrict__ obs,
             float* __restrict__ fwd, float* __restrict__ marg, long long B, long long Tn, int n_sym) {
    constexpr int K = 64, LOOK = 4, VS = 72;  // VS: the second half of the message is shifted by 4 banks
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    float* sh_v = reinterpret_cast<float*>(smem_raw) + warp * 2 * VS;  // [2][VS] carried message, double buffered
    float* sh_em = reinterpret_cast<float*>(smem_raw) + n_warps * 2 * VS;  // [M][64] emission messages (if they fit)
    const int lane = threadIdx.x & 31, jg = lane >> 1, ig = lane & 1;
    const long long b = (long long)blockIdx.x * n_warps + warp;
    float2 a2[4][16];  // a2[jj][ip] = (tbl[32 ig + 2 ip][4 jg + jj], tbl[32 ig + 2 ip + 1][4 jg + jj])
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int ip = 0; ip < 16; ++ip) {
            const int j = 4 * jg + jj, i = 32 * ig + 2 * ip;
            a2[jj][ip] = make_float2(tbl[(size_t)i * K + j], tbl[(size_t)(i + 1) * K + j]);
        }
    if (EM_SMEM) {
        for (int x = threadIdx.x; x < n_sym * K; x += blockDim.x) sh_em[x] = emis_n[x];
        __syncthreads();
    }
    if (b >= B) return;  // whole warps leave; nothing below synchronises across warps
    const int my_slot = 2 * lane + 4 * (lane >> 4);  // where states (2 lane, 2 lane + 1) live in sh_v

    auto time_of = [&](long long step) { return FWD ? step : Tn - 1 - step; };
    auto load_obs_block = [&](long long step0) -> int {  // lane l fetches the symbol of step0 + l
        long long st = step0 + lane;
        return (st < Tn) ? (int)obs[(size_t)time_of(st) * B + b] : 0;
    };
    auto fwd_at = [&](long long step) -> float2 {  // BWD: forward message of `step` for this lane's two states
        return (step < Tn) ? __ldcs(reinterpret_cast<const float2*>(&fwd[((size_t)time_of(step) * B + b) * K]) + lane)
                           : make_float2(0.0f, 0.0f);
    };
    int obs_cur = 0, obs_next = load_obs_block(0);
    float2 a_cur[LOOK], a_nxt[LOOK];
#pragma unroll
    for (int u = 0; u < LOOK; ++u) {
        a_cur[u] = make_float2(0.0f, 0.0f);
        a_nxt[u] = FWD ? make_float2(0.0f, 0.0f) : fwd_at(u);
    }
    float2 u1 = make_float2(0.0f, 0.0f);  // carried message of the previous step (scaled, not normalised)
    float2 ring[LOOK];                      // what steps s-1 .. s-4 leave behind (FWD: the message, BWD: fwd * bwd)
#pragma unroll
    for (int u = 0; u < LOOK; ++u) ring[u] = make_float2(0.0f, 0.0f);
    float p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;  // butterfly partial sums in flight (one level per step)

    // One step. GEN = false is the steady state (4 <= s < Tn): straight-line code.
    //  * scale: r = 2^-floor(log2 max(u_{s-1})) from ONE warp-wide redux.sync.max.f32 (exact scaling, bounded range);
    //  * exact sums for the write-out are a software-pipelined butterfly: every step issues the five shuffles of five
    //    different ages (no dependent shuffle chain inside a step), so the value of step s-4 leaves, exactly
    //    normalised, during step s.
    auto step = [&](auto gen_tag, long long s, int slot, float2 a_now) {
        constexpr bool GEN = decltype(gen_tag)::value;
        int o = __shfl_sync(0xffffffffu, obs_cur, (int)(s & 31));
        if (o >= n_sym) o = n_sym - 1;
        const float2 em = EM_SMEM ? *reinterpret_cast<const float2*>(sh_em + o * K + 2 * lane)
                                  : __ldg(reinterpret_cast<const float2*>(emis_n + (size_t)o * K) + lane);
        float mx;
        {
            const float lm = fmaxf(u1.x, u1.y);
            asm("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(mx) : "f"(lm));
        }
        const float r = hmm_pow2_inv(mx);
        {   // pipelined butterfly: ring[(s-1)&3] entered at level 1, ring[s&3] (step s-4) completes now
            const float2 x1 = ring[(slot + LOOK - 1) % LOOK];
            const float q1 = x1.x + x1.y;
            const float n1 = q1 + __shfl_xor_sync(0xffffffffu, q1, 16);
            const float n2 = p1 + __shfl_xor_sync(0xffffffffu, p1, 8);
            const float n3 = p2 + __shfl_xor_sync(0xffffffffu, p2, 4);
            const float m4 = p3 + __shfl_xor_sync(0xffffffffu, p3, 2);
            const float tot = m4 + __shfl_xor_sync(0xffffffffu, m4, 1);
            p1 = n1;
            p2 = n2;
            p3 = n3;
            if (!GEN || (s >= LOOK && s - LOOK < Tn)) {
                const float q = hmm_rcp(tot);
                const float2 x4 = ring[slot];
                float* dst = FWD ? fwd : marg;
                __stcs(reinterpret_cast<float2*>(&dst[((size_t)time_of(s - LOOK) * B + b) * K]) + lane,
                       make_float2(x4.x * q, x4.y * q));
            }
        }
        if (!GEN || s < Tn) {
            float2 val = make_float2(1.0f, 1.0f);
            if (!GEN || s > 0) {
                const float4* vp = reinterpret_cast<const float4*>(sh_v + (s & 1) * VS + 36 * ig);
                float4 v4[8];  // the whole half-message first: 8 independent 128-bit shared loads in flight
#pragma unroll
                for (int q = 0; q < 8; ++q) v4[q] = vp[q];
                float2 c[4], e[4];  // 8 independent FFMA2 chains
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) c[jj] = e[jj] = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float2 lo = make_float2(v4[q].x, v4[q].y), hi = make_float2(v4[q].z, v4[q].w);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) c[jj] = ffma2(a2[jj][2 * q], lo, c[jj]);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) e[jj] = ffma2(a2[jj][2 * q + 1], hi, e[jj]);
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) c[jj] = make_float2(c[jj].x + e[jj].x, c[jj].y + e[jj].y);
                // transpose-reduce over the two i-halves: lane ig keeps states 2 ig, 2 ig + 1 of its j-group
                const float acc0 = c[0].x + c[0].y, acc1 = c[1].x + c[1].y, acc2 = c[2].x + c[2].y, acc3 = c[3].x + c[3].y;
                const float keep0 = ig ? acc2 : acc0, keep1 = ig ? acc3 : acc1;
                const float send0 = ig ? acc0 : acc2, send1 = ig ? acc1 : acc3;
                val.x = keep0 + __shfl_xor_sync(0xffffffffu, send0, 1);
                val.y = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);
            }
            const float2 unew = (!GEN || s > 0) ? make_float2(em.x * val.x * r, em.y * val.y * r) : em;
            ring[slot] = FWD ? unew : make_float2(a_now.x * val.x, a_now.y * val.y);
            *reinterpret_cast<float2*>(sh_v + ((s + 1) & 1) * VS + my_slot) = unew;
            u1 = unew;
            __syncwarp();
        }
    };

    for (long long s0 = 0; s0 < Tn + LOOK; s0 += LOOK) {
        if ((s0 & 31) == 0) {
            obs_cur = obs_next;
            obs_next = load_obs_block(s0 + 32);
        }
        if (!FWD) {
#pragma unroll
            for (int u = 0; u < LOOK; ++u) {
                a_cur[u] = a_nxt[u];
                a_nxt[u] = fwd_at(s0 + LOOK + u);
            }
            if (s0 + 16 < Tn && lane < 2)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&fwd[((size_t)time_of(s0 + 16) * B + b) * K + 32 * lane]));
        }
        if (s0 >= LOOK && s0 + LOOK <= Tn) {
#pragma unroll
            for (int uu = 0; uu < LOOK; ++uu) step(std::false_type{}, s0 + uu, uu, a_cur[uu]);
        } else {
#pragma unroll
            for (int uu = 0; uu < LOOK; ++uu) step(std::true_type{}, s0 + uu, uu, a_cur[uu]);
        }
    }
}

// ---- K = 64, fp32: TWO warps per chain (one CTA of 64 threads per chain: 2,048 warps = 3.5 per scheduler at B = 1,024) ----
// MEASURED ALTERNATIVE, off by default (see launch_k64): slower than one warp per chain at config-3 size.
// With one warp per chain the batch of config 3 leaves 1.75 warps per scheduler and every step pays its dependent chain
// (shared loads -> 64 FFMA2 -> shuffles) almost in full (630 cycles per step against a 221-cycle FMA-pipe floor). Here warp h
// of a chain owns the output states [32 h, 32 h + 32): lane (cg = lane / 4, rg = lane % 4) holds the 16 x 4 tile
// tbl[16 rg .. 16 rg + 15][32 h + 4 cg .. + 3] as 32 packed pairs (half the registers), does 32 FFMA2 per step on four
// 128-bit shared loads, and a 3-shuffle transpose-reduce over the four row groups leaves output state 32 h + lane in the
// lane (one float per lane: 128 contiguous bytes per warp in every global access). The two warps exchange the carried
// message, the maximum of their half (exact power-of-two scaling, as above) and the partial sums of the deferred exact
// normalisation through double-buffered shared memory with ONE 64-thread barrier per step. The exactly normalised
// message / marginal of step s-5 leaves during step s (per-warp pipelined butterfly, 4 steps; cross-warp total, 1 more).
template <bool FWD, bool EM_SMEM>
__global__ void __launch_bounds__(64, 7)  // 7 chains per SM: 1,024 chains in one wave on 148 SMs
k_hmm64_split(const float* __restrict__ tbl, const float* __restrict__ emis_n, const uint8_t* __restrict__ obs,
              float* __restrict__ fwd, float* __restrict__ marg, long long B, long long Tn, int n_sym) {
    constexpr int K = 64, LOOK = 8, LAG = 5, VS = 72;  // VS: states 32 .. 63 are shifted by 4 banks
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sh_v = reinterpret_cast<float*>(smem_raw);  // [2][VS] carried message, double buffered
    float* sh_mx = sh_v + 2 * VS;                      // [2][2] max of each warp's half of the carried message
    float* sh_tot = sh_mx + 4;                         // [2][2] per-warp sums of the value that leaves next
    float* sh_em = sh_tot + 4;                         // [M][64] emission messages (if they fit)
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31, cg = lane >> 2, rg = lane & 3;
    const long long b = blockIdx.x;
    float2 a2[4][8];  // a2[jj][ip] = (tbl[16 rg + 2 ip][j], tbl[16 rg + 2 ip + 1][j]), j = 32 h + 4 cg + jj
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int ip = 0; ip < 8; ++ip) {
            const int j = 32 * h + 4 * cg + jj, i = 16 * rg + 2 * ip;
            a2[jj][ip] = make_float2(tbl[(size_t)i * K + j], tbl[(size_t)(i + 1) * K + j]);
        }
    if (EM_SMEM)
        for (int x = threadIdx.x; x < n_sym * K; x += blockDim.x) sh_em[x] = emis_n[x];
    if (threadIdx.x < 4) {
        sh_mx[threadIdx.x] = 1.0f;
        sh_tot[threadIdx.x] = 1.0f;
    }
    __syncthreads();
    const int my_state = 32 * h + lane;
    const int my_slot = my_state + 4 * h;            // where this lane's state lives in sh_v
    const int rd_off = 16 * rg + 4 * (rg >> 1);      // this lane's 16 input states

    auto time_of = [&](long long step) { return FWD ? step : Tn - 1 - step; };
    auto load_obs_block = [&](long long step0) -> int {  // lane l fetches the symbol of step0 + l
        long long st = step0 + lane;
        return (st < Tn) ? (int)obs[(size_t)time_of(st) * B + b] : 0;
    };
    auto fwd_at = [&](long long step) -> float {  // BWD: forward message of `step` for this lane's state
        return (step < Tn) ? __ldcs(&fwd[((size_t)time_of(step) * B + b) * K + my_state]) : 0.0f;
    };
    int obs_cur = 0, obs_next = load_obs_block(0);
    float a_cur[LOOK], a_nxt[LOOK];
#pragma unroll
    for (int u = 0; u < LOOK; ++u) {
        a_cur[u] = 0.0f;
        a_nxt[u] = FWD ? 0.0f : fwd_at(u);
    }
    float ring[LOOK];  // what steps s-1 .. s-8 leave behind (FWD: the message, BWD: fwd * bwd)
#pragma unroll
    for (int u = 0; u < LOOK; ++u) ring[u] = 0.0f;
    float p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;  // butterfly partial sums in flight (one level per step)

    auto step = [&](auto gen_tag, long long s, int slot, float a_now) {
        constexpr bool GEN = decltype(gen_tag)::value;
        const int cur = (int)(s & 1), prv = cur ^ 1;
        int o = __shfl_sync(0xffffffffu, obs_cur, (int)(s & 31));
        if (o >= n_sym) o = n_sym - 1;
        const float em = EM_SMEM ? sh_em[o * K + my_state] : __ldg(emis_n + (size_t)o * K + my_state);
        // what the two warps left at the previous step: maxima of the carried message, sums of the value of step s-5
        const float2 mx2 = *reinterpret_cast<const float2*>(sh_mx + 2 * prv);
        const float2 tt2 = *reinterpret_cast<const float2*>(sh_tot + 2 * prv);
        const float r = hmm_pow2_inv(fmaxf(mx2.x, mx2.y));
        if (!GEN || (s >= LAG && s - LAG < Tn)) {
            const float q = hmm_rcp(tt2.x + tt2.y);
            float* dst = FWD ? fwd : marg;
            __stcs(&dst[((size_t)time_of(s - LAG) * B + b) * K + my_state], ring[(slot + LOOK - LAG) % LOOK] * q);
        }
        float tot;
        {   // pipelined butterfly over the warp's 32 states: step s-1 enters, step s-4 completes
            const float x1 = ring[(slot + LOOK - 1) % LOOK];
            const float n1 = x1 + __shfl_xor_sync(0xffffffffu, x1, 16);
            const float n2 = p1 + __shfl_xor_sync(0xffffffffu, p1, 8);
            const float n3 = p2 + __shfl_xor_sync(0xffffffffu, p2, 4);
            const float m4 = p3 + __shfl_xor_sync(0xffffffffu, p3, 2);
            tot = m4 + __shfl_xor_sync(0xffffffffu, m4, 1);
            p1 = n1;
            p2 = n2;
            p3 = n3;
        }
        float unew = 0.0f;
        if (!GEN || s < Tn) {
            float val = 1.0f;
            if (!GEN || s > 0) {
                const float4* vp = reinterpret_cast<const float4*>(sh_v + cur * VS + rd_off);
                float4 v4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v4[q] = vp[q];
                float2 c[4], e[4];  // 8 independent FFMA2 chains of length 4
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) c[jj] = e[jj] = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 lo = make_float2(v4[q].x, v4[q].y), hi = make_float2(v4[q].z, v4[q].w);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) c[jj] = ffma2(a2[jj][2 * q], lo, c[jj]);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) e[jj] = ffma2(a2[jj][2 * q + 1], hi, e[jj]);
                }
                float acc[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[jj] = (c[jj].x + e[jj].x) + (c[jj].y + e[jj].y);
                // transpose-reduce over the four row groups: lane rg ends with column jj = rg of its column group
                const bool b0 = rg & 1, b1 = rg & 2;
                const float kA = (b0 ? acc[1] : acc[0]) + __shfl_xor_sync(0xffffffffu, b0 ? acc[0] : acc[1], 1);
                const float kB = (b0 ? acc[3] : acc[2]) + __shfl_xor_sync(0xffffffffu, b0 ? acc[2] : acc[3], 1);
                val = (b1 ? kB : kA) + __shfl_xor_sync(0xffffffffu, b1 ? kA : kB, 2);
            }
            unew = (!GEN || s > 0) ? em * val * r : em;
            ring[slot] = FWD ? unew : a_now * val;
            sh_v[prv * VS + my_slot] = unew;
        }
        float mx;
        asm("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(mx) : "f"(unew));
        if (lane == 0) {
            sh_mx[2 * cur + h] = mx;
            sh_tot[2 * cur + h] = tot;
        }
        __syncthreads();  // the chain's two warps: message, maxima and sums of this step are visible
    };

    for (long long s0 = 0; s0 < Tn + LAG; s0 += LOOK) {
        if ((s0 & 31) == 0) {
            obs_cur = obs_next;
            obs_next = load_obs_block(s0 + 32);
        }
        if (!FWD) {
#pragma unroll
            for (int u = 0; u < LOOK; ++u) {
                a_cur[u] = a_nxt[u];
                a_nxt[u] = fwd_at(s0 + LOOK + u);
            }
            if (s0 + 32 < Tn && lane < 2)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&fwd[((size_t)time_of(s0 + 32) * B + b) * K + 32 * lane]));
        }
        if (s0 >= LOOK && s0 + LOOK <= Tn) {
#pragma unroll
            for (int uu = 0; uu < LOOK; ++uu) step(std::false_type{}, s0 + uu, uu, a_cur[uu]);
        } else {
#pragma unroll
            for (int uu = 0; uu < LOOK; ++uu) step(std::true_type{}, s0 + uu, uu, a_cur[uu]);
        }
    }
}

// ---- K = 64, fp32, tensor cores: one CTA of 4 warps per 8 chains (1,024 chains = 128 CTAs, one per SM) -----------------------
// The sequential recursion leaves one 64 x 64 x 8 product per SM and step: warp w computes output states 16 w .. 16 w + 15 of
// 8 chains with mma.sync.m16n8k16 (bf16 split operands, hi*hi + hi*lo + lo*hi, fp32 accumulate: 12 MMAs per warp and step
// instead of 64 FFMA2 per lane), the table fragments stay in registers, the 8 carried messages are exchanged between the
// 4 warps through a double-buffered bf16 hi/lo image in shared memory (one block barrier per step). Scaling and deferred
// exact normalisation as in k_hmm64_pass: per-chain max / sum partials travel with the message, the previous step's result
// is written out, exactly normalised, one step late.
__device__ __forceinline__ void mma_bf16_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
}
template <bool FWD>
__global__ void __launch_bounds__(128, 1)
k_hmm64_mma(const float* __restrict__ tbl, const float* __restrict__ emis_n, const uint8_t* __restrict__ obs,
            float* __restrict__ fwd, float* __restrict__ marg, long long B, long long Tn, int n_sym) {
    constexpr int K = 64, NB = 8, VS = 72;  // VS: bf16 elements per chain row (64 + 8: rows start 4 banks apart)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t* sV = reinterpret_cast<uint16_t*>(smem_raw);                  // [2 buffers][2 (hi, lo)][NB][VS]
    float* sPart = reinterpret_cast<float*>(sV + 2 * 2 * NB * VS);         // [2 buffers][2 (max, sum)][NB][4 warps]
    uint8_t* sObs = reinterpret_cast<uint8_t*>(sPart + 2 * 2 * NB * 4);    // [2 buffers][32 steps][NB]
    float* sEm = reinterpret_cast<float*>(sObs + 2 * 32 * NB);             // [n_sym][64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int j0 = 16 * warp;
    const long long b0 = (long long)blockIdx.x * NB;
    auto time_of = [&](long long step) { return FWD ? step : Tn - 1 - step; };

    // table fragments: M[j][i] = tbl[i][j] (out[j] = sum_i tbl[i][j] v[i]), rows j0 + g / j0 + g + 8, 4 K-steps of 16
    uint32_t ahi[4][4], alo[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = j0 + g + 8 * (r & 1), i = 16 * ks + 2 * t + 8 * (r >> 1);
            uint16_t h0, l0, h1, l1;
            split_bf16(tbl[(size_t)i * K + j], h0, l0);
            split_bf16(tbl[(size_t)(i + 1) * K + j], h1, l1);
            ahi[ks][r] = (uint32_t)h0 | ((uint32_t)h1 << 16);
            alo[ks][r] = (uint32_t)l0 | ((uint32_t)l1 << 16);
        }
    for (int x = threadIdx.x; x < n_sym * K; x += blockDim.x) sEm[x] = emis_n[x];
    auto load_obs_block = [&](long long step0, int buf) {  // 32 steps x 8 chains, two entries per thread
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = threadIdx.x * 2 + e, st = idx / NB, n = idx % NB;
            const long long step = step0 + st;
            uint8_t o = 0;
            if (step < Tn && b0 + n < B) o = obs[(size_t)time_of(step) * B + b0 + n];
            sObs[(buf * 32 + st) * NB + n] = o;
        }
    };
    load_obs_block(0, 0);
    // this thread's four cells: (state j0 + g + 8 h, chain 2 t + c), h, c in {0, 1}; cell index = 2 h + c (the D fragment order)
    auto cell_ptr = [&](float* plane, long long step, int h, int c) -> float* {
        return plane + ((size_t)time_of(step) * B + (size_t)(b0 + 2 * t + c)) * K + j0 + g + 8 * h;
    };
    const bool live_c[2] = {b0 + 2 * t < B, b0 + 2 * t + 1 < B};
    float a_cur[4] = {0, 0, 0, 0}, a_n1[4] = {0, 0, 0, 0}, a_n2[4] = {0, 0, 0, 0};  // BWD: forward message cells of steps s, s+1, s+2
    auto load_fwd_cells = [&](long long step, float (&dst)[4]) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                dst[2 * h + c] = (step < Tn && live_c[c]) ? __ldcs(cell_ptr(fwd, step, h, c)) : 0.0f;
    };
    if (!FWD) {
        load_fwd_cells(0, a_cur);
        load_fwd_cells(1, a_n1);
    }
    float uprev[4] = {0, 0, 0, 0}, gprev[4] = {0, 0, 0, 0};
    __syncthreads();

    for (long long s = 0; s <= Tn; ++s) {
        if ((s & 31) == 0) load_obs_block(s + 32, (int)(((s >> 5) + 1) & 1));  // read 32 steps from now
        if (!FWD) {
            load_fwd_cells(s + 2, a_n2);
            if (s + 16 < Tn && threadIdx.x < 2 * NB)  // pull the forward messages of step s + 16 into L2 (8 chains x 256 B)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(fwd + ((size_t)time_of(s + 16) * B + (size_t)(b0 + (threadIdx.x >> 1))) * K +
                                                                32 * (threadIdx.x & 1)));
        }
        float r[2] = {1.0f, 1.0f};
        if (s >= 1) {  // totals of step s - 1 (partials of the 4 warps), its exactly normalised write-out, this step's scale
            const int pb = (int)((s - 1) & 1);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int n = 2 * t + c;
                const float4 pm = *reinterpret_cast<const float4*>(sPart + ((pb * 2 + 0) * NB + n) * 4);
                const float4 ps = *reinterpret_cast<const float4*>(sPart + ((pb * 2 + 1) * NB + n) * 4);
                r[c] = hmm_pow2_inv(fmaxf(fmaxf(pm.x, pm.y), fmaxf(pm.z, pm.w)));
                const float q = hmm_rcp((ps.x + ps.y) + (ps.z + ps.w));
                if (live_c[c]) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        __stcs(cell_ptr(FWD ? fwd : marg, s - 1, h, c), (FWD ? uprev[2 * h + c] : gprev[2 * h + c]) * q);
                }
            }
        }
        if (s < Tn) {
            float pred[4] = {1.0f, 1.0f, 1.0f, 1.0f};
            if (s > 0) {
                const uint16_t* vhi = sV + ((size_t)((s & 1) * 2 + 0) * NB + g) * VS;
                const uint16_t* vlo = sV + ((size_t)((s & 1) * 2 + 1) * NB + g) * VS;
                // twelve INDEPENDENT accumulators: the MMAs of a step do not wait for one another (mma.sync latency, not
                // throughput, is what a step pays for), the partial products are added afterwards
                float d[4][3][4];
                uint32_t bh0[4], bh1[4], bl0[4], bl1[4];
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    bh0[ks] = *reinterpret_cast<const uint32_t*>(vhi + 16 * ks + 2 * t);
                    bh1[ks] = *reinterpret_cast<const uint32_t*>(vhi + 16 * ks + 2 * t + 8);
                    bl0[ks] = *reinterpret_cast<const uint32_t*>(vlo + 16 * ks + 2 * t);
                    bl1[ks] = *reinterpret_cast<const uint32_t*>(vlo + 16 * ks + 2 * t + 8);
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int i = 0; i < 4; ++i) d[ks][p][i] = 0.0f;
                    mma_bf16_m16n8k16(d[ks][0], ahi[ks], bh0[ks], bh1[ks]);
                    mma_bf16_m16n8k16(d[ks][1], ahi[ks], bl0[ks], bl1[ks]);
                    mma_bf16_m16n8k16(d[ks][2], alo[ks], bh0[ks], bh1[ks]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float hh = (d[0][0][i] + d[1][0][i]) + (d[2][0][i] + d[3][0][i]);
                    const float cr = ((d[0][1][i] + d[0][2][i]) + (d[1][1][i] + d[1][2][i])) + ((d[2][1][i] + d[2][2][i]) + (d[3][1][i] + d[3][2][i]));
                    pred[i] = hh + cr;
                }
            }
            const uint8_t* ob = sObs + ((size_t)((s >> 5) & 1) * 32 + (s & 31)) * NB;
            float unew[4], gnew[4];
            float mx[2] = {0.0f, 0.0f}, sm[2] = {0.0f, 0.0f};
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                int o = ob[2 * t + c];
                if (o >= n_sym) o = n_sym - 1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float em = sEm[o * K + j0 + g + 8 * h];
                    const int cell = 2 * h + c;
                    unew[cell] = (s > 0) ? em * pred[cell] * r[c] : em;
                    gnew[cell] = FWD ? unew[cell] : a_cur[cell] * pred[cell];
                    mx[c] = fmaxf(mx[c], unew[cell]);
                    sm[c] += gnew[cell];
                }
            }
            // per-chain max / sum over this warp's 16 states: lanes with the same t (xor 4, 8, 16)
#pragma unroll
            for (int d = 4; d < 32; d <<= 1)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
                    sm[c] += __shfl_xor_sync(0xffffffffu, sm[c], d);
                }
            const int wb = (int)(s & 1);
            if (g == 0) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    sPart[((wb * 2 + 0) * NB + 2 * t + c) * 4 + warp] = mx[c];
                    sPart[((wb * 2 + 1) * NB + 2 * t + c) * 4 + warp] = sm[c];
                }
            }
            // the carried message of the next step, as bf16 hi / lo rows per chain
            const int nb = (int)((s + 1) & 1);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint16_t hi, lo;
                    split_bf16(unew[2 * h + c], hi, lo);
                    sV[((size_t)(nb * 2 + 0) * NB + 2 * t + c) * VS + j0 + g + 8 * h] = hi;
                    sV[((size_t)(nb * 2 + 1) * NB + 2 * t + c) * VS + j0 + g + 8 * h] = lo;
                }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uprev[i] = unew[i];
                gprev[i] = gnew[i];
            }
        }
        if (!FWD) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a_cur[i] = a_n1[i];
                a_n1[i] = a_n2[i];
            }
        }
        __syncthreads();
    }
}

struct Hmm {
    int device = 0, dtype = CXB_F32, K = 0, M = 0;
    long long B = 0, T = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    DBuf<unsigned char> A, At, En, fwd, marg;
    DBuf<uint8_t> obs;
    // tensor-core path (hmm_tc.cuh): table images (forward: rows of A^T, backward: rows of A), message operands, row sums
    DBuf<uint16_t> tc_img_f, tc_img_b, tc_op[2], tc_op_b[2];  // _b: the backward pass's own ping-pong (paired launches)
    DBuf<float> tc_part[2], tc_part_b[2];
    bool tc_ready = false;
    bool tc_paired = true;
    int tc_nt = 64;  // output states per CTA of the tensor-core step kernel (32 or 64)
    int tc_np = 3;   // bf16 pieces per operand (3: fp32-level accuracy, the default; 2: ~2^-17 per term, CXB_HMM_TC_PIECES=2)
    bool tc_eligible() const {
        if (const char* e = getenv("CXB_HMM_NO_TC"))
            if (atoi(e)) return false;
        return dtype == CXB_F32 && K % 64 == 0 && K >= 128 && K <= tc::MAX_K && M <= 64;
    }
    long long tc_bpad() const { return (B + tc::M_TILE - 1) / tc::M_TILE * tc::M_TILE; }
    bool have_tables = false, have_obs = false, ran = false;
    size_t esz() const { return dtype == CXB_F32 ? 4 : 8; }
    ~Hmm() {
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
        if (device < 0 || device >= count || B <= 0 || T <= 0 || K < 2 || K > 1024 || M < 1 || M > 256) {
            err = "bad device / shape (2 <= states <= 1024, 1 <= symbols <= 256)";
            return CXB_ERR_BAD_ARG;
        }
        CXB_CUDA(cudaSetDevice(device));
        CXB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CXB_CUDA(cudaEventCreate(&ev0));
        CXB_CUDA(cudaEventCreate(&ev1));
        CXB_CUDA(A.reserve((size_t)K * K * esz()));
        CXB_CUDA(At.reserve((size_t)K * K * esz()));
        CXB_CUDA(En.reserve((size_t)M * K * esz()));
        CXB_CUDA(obs.reserve((size_t)T * B));
        CXB_CUDA(fwd.reserve((size_t)T * B * K * esz()));
        CXB_CUDA(marg.reserve((size_t)T * B * K * esz()));
        return CXB_OK;
    }
    void put(std::vector<unsigned char>& raw, size_t i, double v) {
        if (dtype == CXB_F32)
            ((float*)raw.data())[i] = (float)v;
        else
            ((double*)raw.data())[i] = v;
    }
    int32_t set_tables(const double* trans, const double* emis) {
        CXB_CUDA(cudaSetDevice(device));
        std::vector<unsigned char> a((size_t)K * K * esz()), at((size_t)K * K * esz()), en((size_t)M * K * esz());
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
                put(a, (size_t)i * K + j, trans[(size_t)i * K + j]);
                put(at, (size_t)j * K + i, trans[(size_t)i * K + j]);
            }
        for (int o = 0; o < M; ++o) {  // m2v(z, em) = normalise(E[:, o]) (HMM_EMIT rule)
            double s = 0;
            for (int j = 0; j < K; ++j) s += emis[(size_t)j * M + o];
            if (!(s > 0)) {
                err = "emission column sums must be positive";
                return CXB_ERR_BAD_ARG;
            }
            for (int j = 0; j < K; ++j) put(en, (size_t)o * K + j, emis[(size_t)j * M + o] / s);
        }
        CXB_CUDA(cudaMemcpyAsync(A.p, a.data(), a.size(), cudaMemcpyHostToDevice, stream));
        CXB_CUDA(cudaMemcpyAsync(At.p, at.data(), at.size(), cudaMemcpyHostToDevice, stream));
        CXB_CUDA(cudaMemcpyAsync(En.p, en.data(), en.size(), cudaMemcpyHostToDevice, stream));
        if (tc_eligible()) {
            // slices of 32 output states when 64-state slices would leave most SMs idle
            int n_sm = 148;
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
            tc_paired = true;  // both passes in one launch per step (k_hmm_tc_step_pair); CXB_HMM_TC_PAIRED=0: one pass after the other
            if (const char* e = getenv("CXB_HMM_TC_PAIRED")) tc_paired = atoi(e) != 0;
            // the widest grid that is still resident at once: (chain tiles) x (K / NT slices) x (passes per launch) CTAs, one per SM
            tc_nt = (tc_bpad() / tc::M_TILE) * (K / 32) * (tc_paired ? 2 : 1) <= n_sm ? 32 : 64;
            if (const char* e = getenv("CXB_HMM_TC_NT")) tc_nt = atoi(e) == 32 ? 32 : 64;
            tc_np = 3;
            if (const char* e = getenv("CXB_HMM_TC_PIECES")) tc_np = atoi(e) == 2 ? 2 : 3;
            std::vector<uint16_t> img;
            tc::build_table_image((const float*)at.data(), K, tc_nt, tc_np, img);  // forward: pred[j] = sum_i msg[i] A[i][j] -> rows of A^T
            CXB_CUDA(tc_img_f.reserve(img.size()));
            CXB_CUDA(cudaMemcpyAsync(tc_img_f.p, img.data(), img.size() * 2, cudaMemcpyHostToDevice, stream));
            CXB_CUDA(cudaStreamSynchronize(stream));
            tc::build_table_image((const float*)a.data(), K, tc_nt, tc_np, img);   // backward: pred[j] = sum_i A[j][i] msg[i] -> rows of A
            CXB_CUDA(tc_img_b.reserve(img.size()));
            CXB_CUDA(cudaMemcpyAsync(tc_img_b.p, img.data(), img.size() * 2, cudaMemcpyHostToDevice, stream));
            const long long bpad = tc_bpad();
            const size_t op_elems = (size_t)(bpad / tc::M_TILE) * tc_np * (K / tc::K_CHUNK) * (tc::A_CHUNK_BYTES / 2);
            for (int i = 0; i < 2; ++i) {
                CXB_CUDA(tc_op[i].reserve(op_elems));
                CXB_CUDA(tc_part[i].reserve((size_t)2 * (K / tc_nt) * bpad));
                CXB_CUDA(cudaMemsetAsync(tc_op[i].p, 0, op_elems * 2, stream));
                CXB_CUDA(cudaMemsetAsync(tc_part[i].p, 0, (size_t)2 * (K / tc_nt) * bpad * sizeof(float), stream));
                if (tc_paired) {
                    CXB_CUDA(tc_op_b[i].reserve(op_elems));
                    CXB_CUDA(tc_part_b[i].reserve((size_t)2 * (K / tc_nt) * bpad));
                    CXB_CUDA(cudaMemsetAsync(tc_op_b[i].p, 0, op_elems * 2, stream));
                    CXB_CUDA(cudaMemsetAsync(tc_part_b[i].p, 0, (size_t)2 * (K / tc_nt) * bpad * sizeof(float), stream));
                }
            }
            tc_ready = true;
        }
        CXB_CUDA(cudaStreamSynchronize(stream));
        have_tables = true;
        return CXB_OK;
    }
    template <class T, int KK, bool REGA>
    int32_t launch_pair(int tile_rows, size_t smem) {
        unsigned grid = cdiv((size_t)B, HMM_WARPS);
        if (smem > 48 * 1024) {
            CXB_CUDA(cudaFuncSetAttribute(k_hmm_pass<T, KK, REGA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CXB_CUDA(cudaFuncSetAttribute(k_hmm_pass<T, KK, REGA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        CXB_LAUNCH((k_hmm_pass<T, KK, REGA, true>), grid, HMM_WARPS * 32, smem, stream, (const T*)A.p, (const T*)En.p, obs.p,
                   (T*)fwd.p, (T*)marg.p, B, this->T, K, M, tile_rows);
        CXB_LAUNCH((k_hmm_pass<T, KK, REGA, false>), grid, HMM_WARPS * 32, smem, stream, (const T*)At.p, (const T*)En.p, obs.p,
                   (T*)fwd.p, (T*)marg.p, B, this->T, K, M, tile_rows);
        return CXB_OK;
    }
    // one launch per time step and pass (hmm_tc.cuh); 2 T + 4 launches
    int32_t launch_tc() {
        if (tc_paired) {
            if (tc_np == 2) return tc_nt == 32 ? launch_tc_pair_nt<32, 2>() : launch_tc_pair_nt<64, 2>();
            return tc_nt == 32 ? launch_tc_pair_nt<32, 3>() : launch_tc_pair_nt<64, 3>();
        }
        if (tc_np == 2) return tc_nt == 32 ? launch_tc_nt<32, 2>() : launch_tc_nt<64, 2>();
        return tc_nt == 32 ? launch_tc_nt<32, 3>() : launch_tc_nt<64, 3>();
    }
    // Paired schedule: launch s runs step s of BOTH passes (forward at time s, backward at time T-1-s): T + 5 launches.
    // The backward half writes the marginal directly once the forward message of its time step is final (normalised in
    // place by forward launch t+1, i.e. for T-1-s <= s-2); before that it stores the normalised backward prediction and
    // k_hmm_tc_combine multiplies the forward messages in at the end (times >= T - s_direct).
    template <int NT, int NP>
    int32_t launch_tc_pair_nt() {
        const int bpad = (int)tc_bpad(), tiles = bpad / tc::M_TILE, n_slices = K / NT;
        const size_t smem = tc::step_smem_bytes<NT, NP>(M), row = (size_t)B * K;
        CXB_CUDA(cudaFuncSetAttribute(tc::k_hmm_tc_step_pair<NT, NP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CXB_CUDA(cudaFuncSetAttribute(tc::k_hmm_tc_step_pair<NT, NP, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CXB_CUDA(cudaFuncSetAttribute(tc::k_hmm_tc_step_pair<NT, NP, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // clusters of 4 or 8 slices of a chain tile with a multicast message operand (CXB_HMM_TC_CLUSTER=4|8), when the slices
        // divide and every cluster of the grid is resident at once (a B200 holds 15 clusters of 8 CTAs of this size: config 3
        // needs 16, so it falls back)
        int cl = 1;
        if (const char* e = getenv("CXB_HMM_TC_CLUSTER")) cl = atoi(e) == 8 ? 8 : atoi(e) == 4 ? 4 : 1;
        if (cl > 1 && n_slices % cl) cl = 1;
        if (cl > 1) {
            cudaLaunchConfig_t q{};
            q.gridDim = dim3(tiles, n_slices, 2);
            q.blockDim = dim3(tc::THREADS);
            q.dynamicSmemBytes = smem;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 1;
            qa[0].val.clusterDim.y = cl;
            qa[0].val.clusterDim.z = 1;
            q.attrs = qa;
            q.numAttrs = 1;
            int n_clusters = 0;
            const int needed = tiles * (n_slices / cl) * 2, asked = cl;
            const cudaError_t qe = cl == 8 ? cudaOccupancyMaxActiveClusters(&n_clusters, tc::k_hmm_tc_step_pair<NT, NP, 8>, &q)
                                           : cudaOccupancyMaxActiveClusters(&n_clusters, tc::k_hmm_tc_step_pair<NT, NP, 4>, &q);
            if (qe != cudaSuccess || n_clusters < needed) cl = 1;
            cudaGetLastError();
            if (getenv("CXB_HMM_TC_VERBOSE"))
                fprintf(stderr, "cxb_hmm tc: clusters of %d requested: %d resident at once (%s), %d needed -> cluster size %d\n", asked, n_clusters,
                        cudaGetErrorString(qe), needed, cl);
        }
        float *fw = (float*)fwd.p, *mg = (float*)marg.p;
        tc::StepArgs2 p{};
        for (int z = 0; z < 2; ++z) {
            p.d[z].emis_n = (const float*)En.p;
            p.d[z].B = (int)B;
            p.d[z].Bpad = bpad;
            p.d[z].K = K;
            p.d[z].n_sym = M;
        }
        p.d[0].tbl_img = (const __nv_bfloat16*)tc_img_f.p;
        p.d[1].tbl_img = (const __nv_bfloat16*)tc_img_b.p;
        const dim3 grid(tiles, n_slices, 2), igrid(bpad / 128, n_slices);
        const bool pdl = !(getenv("CXB_HMM_TC_NO_PDL") && atoi(getenv("CXB_HMM_TC_NO_PDL")));
        const long long Tn = this->T, s_direct = (Tn + 2) / 2;  // first launch whose backward half sees a final forward message
        DBuf<long long> trace;
        const bool tracing = getenv("CXB_HMM_TC_TRACE") && atoi(getenv("CXB_HMM_TC_TRACE"));
        const size_t n_cta = (size_t)tiles * n_slices * 2;
        if (tracing) {
            CXB_CUDA(trace.reserve(n_cta * 16));
            CXB_CUDA(cudaMemsetAsync(trace.p, 0, n_cta * 16 * sizeof(long long), stream));
            p.d[0].trace = p.d[1].trace = trace.p;
        }
        for (long long s = 0; s < Tn; ++s) {
            const long long tf = s, tb = Tn - 1 - s;
            tc::StepArgs& f = p.d[0];
            tc::StepArgs& b = p.d[1];
            f.op_in = (const __nv_bfloat16*)tc_op[(s + 1) & 1].p;
            f.op_out = (__nv_bfloat16*)tc_op[s & 1].p;
            f.part_in = tc_part[(s + 1) & 1].p;
            f.part_out = tc_part[s & 1].p;
            f.obs_t = obs.p + (size_t)tf * B;
            f.raw_out = fw + (size_t)tf * row;
            f.raw_prev = s ? fw + (size_t)(tf - 1) * row : nullptr;
            f.fwd_t = nullptr;
            b.op_in = (const __nv_bfloat16*)tc_op_b[(s + 1) & 1].p;
            b.op_out = (__nv_bfloat16*)tc_op_b[s & 1].p;
            b.part_in = tc_part_b[(s + 1) & 1].p;
            b.part_out = tc_part_b[s & 1].p;
            b.obs_t = obs.p + (size_t)tb * B;
            b.raw_out = mg + (size_t)tb * row;
            b.raw_prev = s ? mg + (size_t)(tb + 1) * row : nullptr;
            b.fwd_t = s >= s_direct ? fw + (size_t)tb * row : nullptr;
            if (s == 0) {
                CXB_LAUNCH((tc::k_hmm_tc_init<true, NT, NP>), igrid, 128, 0, stream, f);
                CXB_LAUNCH((tc::k_hmm_tc_init<false, NT, NP>), igrid, 128, 0, stream, b);
            } else {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = grid;
                cfg.blockDim = dim3(tc::THREADS);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = stream;
                cudaLaunchAttribute attr[2];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = (pdl && s > 1) ? 1 : 0;  // launch 1 follows the two init kernels
                attr[1].id = cudaLaunchAttributeClusterDimension;
                attr[1].val.clusterDim.x = 1;
                attr[1].val.clusterDim.y = cl;
                attr[1].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = cl > 1 ? 2 : 1;
                if (cl == 8)
                    CXB_CUDA(cudaLaunchKernelEx(&cfg, tc::k_hmm_tc_step_pair<NT, NP, 8>, p));
                else if (cl == 4)
                    CXB_CUDA(cudaLaunchKernelEx(&cfg, tc::k_hmm_tc_step_pair<NT, NP, 4>, p));
                else
                    CXB_CUDA(cudaLaunchKernelEx(&cfg, tc::k_hmm_tc_step_pair<NT, NP, 1>, p));
                ++::cxb::g_kernel_launches;
            }
        }
        CXB_LAUNCH(tc::k_hmm_tc_finish, (unsigned)B, 128, 0, stream, fw + (size_t)(Tn - 1) * row, tc_part[(Tn - 1) & 1].p, (int)B, bpad, K,
                   n_slices);
        CXB_LAUNCH(tc::k_hmm_tc_finish, (unsigned)B, 128, 0, stream, mg, tc_part_b[(Tn - 1) & 1].p, (int)B, bpad, K, n_slices);
        const long long t0 = Tn - std::min(s_direct, Tn), n_rows = (Tn - t0) * B;  // times the backward half ran ahead of the forward pass
        if (n_rows > 0)
            CXB_LAUNCH(tc::k_hmm_tc_combine, (unsigned)((n_rows + 7) / 8), 256, 0, stream, fw + (size_t)t0 * row, mg + (size_t)t0 * row, n_rows, K);
        if (tracing) {  // stamps of the LAST paired launch, averaged over the CTAs of each half, relative to the earliest CTA start
            std::vector<long long> h(n_cta * 16);
            CXB_CUDA(cudaStreamSynchronize(stream));
            CXB_CUDA(cudaMemcpy(h.data(), trace.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            static const char* names[10] = {"start", "init+alloc done", "producer issued all", "mma: first chunk landed", "mma: last chunk landed",
                                            "mma: all issued", "epi: prev normalised", "epi: accumulator ready", "epi: done", "exit"};
            for (int z = 0; z < 2; ++z)
                for (int k = 0; k < 10; ++k) {
                    double acc = 0;
                    const size_t c0 = (size_t)z * tiles * n_slices, c1 = c0 + (size_t)tiles * n_slices;
                    for (size_t c = c0; c < c1; ++c) acc += (double)(h[c * 16 + k] - h[c * 16]);
                    fprintf(stderr, "cxb_hmm tc trace (%s half): %-26s %8.0f cycles after the CTA's own start\n", z ? "backward" : "forward", names[k],
                            acc / (tiles * n_slices));
                }
        }
        return CXB_OK;
    }
    template <int NT, int NP>
    int32_t launch_tc_nt() {
        const int bpad = (int)tc_bpad(), tiles = bpad / tc::M_TILE, n_slices = K / NT;
        const size_t smem = tc::step_smem_bytes<NT, NP>(M), row = (size_t)B * K;
        CXB_CUDA(cudaFuncSetAttribute(tc::k_hmm_tc_step<true, NT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CXB_CUDA(cudaFuncSetAttribute(tc::k_hmm_tc_step<false, NT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float *fw = (float*)fwd.p, *mg = (float*)marg.p;
        tc::StepArgs a{};
        a.emis_n = (const float*)En.p;
        a.B = (int)B;
        a.Bpad = bpad;
        a.K = K;
        a.n_sym = M;
        const dim3 grid(tiles, n_slices), igrid(bpad / 128, n_slices);
        DBuf<long long> trace;
        const bool tracing = getenv("CXB_HMM_TC_TRACE") && atoi(getenv("CXB_HMM_TC_TRACE"));
        const bool pdl = !(getenv("CXB_HMM_TC_NO_PDL") && atoi(getenv("CXB_HMM_TC_NO_PDL")));
        if (tracing) {
            CXB_CUDA(trace.reserve((size_t)tiles * n_slices * 16));
            CXB_CUDA(cudaMemsetAsync(trace.p, 0, (size_t)tiles * n_slices * 16 * sizeof(long long), stream));
            a.trace = trace.p;
        }
        for (int pass = 0; pass < 2; ++pass) {
            const bool f = pass == 0;
            a.tbl_img = (const __nv_bfloat16*)(f ? tc_img_f.p : tc_img_b.p);
            float* out = f ? fw : mg;
            for (long long s = 0; s < this->T; ++s) {
                const long long t = f ? s : this->T - 1 - s, tp = f ? t - 1 : t + 1;
                a.op_in = (const __nv_bfloat16*)tc_op[(s + 1) & 1].p;
                a.op_out = (__nv_bfloat16*)tc_op[s & 1].p;
                a.part_in = tc_part[(s + 1) & 1].p;
                a.part_out = tc_part[s & 1].p;
                a.obs_t = obs.p + (size_t)t * B;
                a.raw_out = out + (size_t)t * row;
                a.raw_prev = s ? out + (size_t)tp * row : nullptr;
                a.fwd_t = f ? nullptr : fw + (size_t)t * row;
                if (s == 0) {
                    if (f)
                        CXB_LAUNCH((tc::k_hmm_tc_init<true, NT, NP>), igrid, 128, 0, stream, a);
                    else
                        CXB_LAUNCH((tc::k_hmm_tc_init<false, NT, NP>), igrid, 128, 0, stream, a);
                } else {
                    cudaLaunchConfig_t cfg{};
                    cfg.gridDim = grid;
                    cfg.blockDim = dim3(tc::THREADS);
                    cfg.dynamicSmemBytes = smem;
                    cfg.stream = stream;
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
                    cfg.attrs = attr;
                    cfg.numAttrs = 1;
                    CXB_CUDA(cudaLaunchKernelEx(&cfg, f ? tc::k_hmm_tc_step<true, NT, NP> : tc::k_hmm_tc_step<false, NT, NP>, a));
                    ++::cxb::g_kernel_launches;
                }
            }
            const long long t_last = f ? this->T - 1 : 0;
            CXB_LAUNCH(tc::k_hmm_tc_finish, (unsigned)B, 128, 0, stream, out + (size_t)t_last * row, tc_part[(this->T - 1) & 1].p, (int)B,
                       bpad, K, n_slices);
        }
        if (tracing) {  // stamps of the LAST step kernel of the backward pass, averaged over the CTAs, relative to CTA start
            std::vector<long long> h((size_t)tiles * n_slices * 16);
            CXB_CUDA(cudaStreamSynchronize(stream));
            CXB_CUDA(cudaMemcpy(h.data(), trace.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            static const char* names[10] = {"start", "init+alloc done", "producer issued all", "mma: first chunk landed", "mma: last chunk landed",
                                            "mma: all issued", "epi: prev normalised", "epi: accumulator ready", "epi: done", "exit"};
            for (int k = 0; k < 10; ++k) {
                double acc = 0;
                for (int c = 0; c < tiles * n_slices; ++c) acc += (double)(h[(size_t)c * 16 + k] - h[(size_t)c * 16]);
                fprintf(stderr, "cxb_hmm tc trace: %-26s %8.0f cycles\n", names[k], acc / (tiles * n_slices));
            }
        }
        return CXB_OK;
    }
    int32_t launch_k64_mma() {
        const size_t smem = (size_t)2 * 2 * 8 * 72 * 2 + (size_t)2 * 2 * 8 * 4 * 4 + 2 * 32 * 8 + (size_t)M * 64 * sizeof(float);
        const unsigned grid = (unsigned)((B + 7) / 8);
        const float *a = (const float*)A.p, *at = (const float*)At.p, *en = (const float*)En.p;
        if (smem > 48 * 1024) {
            CXB_CUDA(cudaFuncSetAttribute(k_hmm64_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CXB_CUDA(cudaFuncSetAttribute(k_hmm64_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        CXB_LAUNCH((k_hmm64_mma<true>), grid, 128, smem, stream, a, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
        CXB_LAUNCH((k_hmm64_mma<false>), grid, 128, smem, stream, at, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
        return CXB_OK;
    }
    // Measured alternative for K = 64, fp32 (CXB_HMM64_TC=1): both recursions in one launch on mma.sync tensor cores with
    // three-piece bf16 operands, meeting in the middle (hmm64_tc.cuh). Parity-green, but slower than the FFMA2 register
    // kernel below: 216 instructions per warp and step on a mostly serial dependency chain run at 0.31 instructions per
    // cycle and scheduler with the two warps a scheduler gets (1,430 cycles per step of both recursions against
    // 2 x 605 for the two FFMA2 passes; profiles/r02_hmm64_tc_ncu.txt).
    int32_t launch_k64_tc() {
        const size_t smem = h64::smem_bytes(M, true);
        const unsigned grid = (unsigned)((B + h64::NB - 1) / h64::NB);
        CXB_CUDA(cudaFuncSetAttribute(h64::k_hmm64_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CXB_LAUNCH((h64::k_hmm64_tc<true>), grid, 256, smem, stream, (const float*)A.p, (const float*)At.p, (const float*)En.p, obs.p,
                   (float*)fwd.p, (float*)marg.p, B, this->T, M);
        return CXB_OK;
    }
    int32_t launch_k64() {
        {
            bool tc64 = false;
            if (const char* e = getenv("CXB_HMM64_TC")) tc64 = atoi(e) != 0;
            const bool other = (getenv("CXB_HMM64_MMA") && atoi(getenv("CXB_HMM64_MMA"))) || (getenv("CXB_HMM64_SPLIT") && atoi(getenv("CXB_HMM64_SPLIT")));
            if (tc64 && !other) return launch_k64_tc();
        }
        // The tensor-core variant (mma.sync, 8 chains per CTA, k_hmm64_mma) is kept as a measured alternative
        // (CXB_HMM64_MMA=1): with one 64 x 64 x 8 product per SM and step it has a single warp per scheduler and pays every
        // latency in full — 970 cycles per step against 630 for the FFMA2 kernel below (T = 4,000: 3.94 ms vs 2.55 ms).
        if (M <= 128 && getenv("CXB_HMM64_MMA") && atoi(getenv("CXB_HMM64_MMA"))) return launch_k64_mma();
        {   // Two warps per chain (k_hmm64_split), a measured alternative (CXB_HMM64_SPLIT=1): it halves the FFMA2 chain of a
            // warp but every warp repeats the per-step bookkeeping (emission, scaling, butterfly, stores) and the pair meets at
            // a barrier, so the step is issue-bound at a HIGHER instruction count: B = 1,024, T = 1e5: 80.6 ms against 63.9 ms
            // for one warp per chain (790 vs 630 cycles per step).
            bool split = false;
            if (const char* e = getenv("CXB_HMM64_SPLIT")) split = atoi(e) != 0;
            if (split) {
                const bool em_smem = M <= 128;
                const size_t smem = (size_t)(2 * 72 + 8) * sizeof(float) + (em_smem ? (size_t)M * 64 * sizeof(float) : 0);
                const float *a = (const float*)A.p, *at = (const float*)At.p, *en = (const float*)En.p;
                const unsigned grid = (unsigned)B;
                if (em_smem) {
                    CXB_LAUNCH((k_hmm64_split<true, true>), grid, 64, smem, stream, a, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
                    CXB_LAUNCH((k_hmm64_split<false, true>), grid, 64, smem, stream, at, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
                } else {
                    CXB_LAUNCH((k_hmm64_split<true, false>), grid, 64, smem, stream, a, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
                    CXB_LAUNCH((k_hmm64_split<false, false>), grid, 64, smem, stream, at, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
                }
                return CXB_OK;
            }
        }
        // one warp per chain; warps per CTA chosen so that one CTA per SM holds the whole batch when it can (its warps
        // then spread evenly over the 4 schedulers): 1,024 chains -> 147 CTAs of 7 warps
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
        int wpc = (int)std::min<long long>(8, std::max<long long>(1, (B + n_sm - 1) / n_sm));
        if (const char* e = getenv("CXB_HMM_WARPS")) wpc = std::max(1, std::min(8, atoi(e)));
        const unsigned grid = (unsigned)((B + wpc - 1) / wpc), threads = 32u * (unsigned)wpc;
        const bool em_smem = M <= 128;
        size_t smem = (size_t)wpc * 2 * 72 * sizeof(float) + (em_smem ? (size_t)M * 64 * sizeof(float) : 0);
        const float *a = (const float*)A.p, *at = (const float*)At.p, *en = (const float*)En.p;
        if (em_smem) {
            CXB_LAUNCH((k_hmm64_pass<true, true>), grid, threads, smem, stream, a, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
            CXB_LAUNCH((k_hmm64_pass<false, true>), grid, threads, smem, stream, at, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
        } else {
            CXB_LAUNCH((k_hmm64_pass<true, false>), grid, threads, smem, stream, a, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
            CXB_LAUNCH((k_hmm64_pass<false, false>), grid, threads, smem, stream, at, en, obs.p, (float*)fwd.p, (float*)marg.p, B, this->T, M);
        }
        return CXB_OK;
    }
    template <class T>
    int32_t launch_t() {
        size_t stage = ((size_t)HMM_WARPS * K + (size_t)M * K) * sizeof(T);  // per-warp staging + emission messages
        if (sizeof(T) == 4 && K == 64) return launch_k64();
        if (sizeof(T) == 4 && tc_ready && tc_eligible()) return launch_tc();
        if (K == 32) return launch_pair<T, 32, true>(0, stage);
        size_t budget = 200 * 1024 - stage;
        int tile_rows = (int)std::min<size_t>((size_t)K, budget / ((size_t)K * sizeof(T)));
        if (tile_rows < 1) {
            err = "state count too large";
            return CXB_ERR_BAD_ARG;
        }
        return launch_pair<T, 0, false>(tile_rows, stage + (size_t)tile_rows * K * sizeof(T));
    }
    int32_t launch() {
        if (!have_tables || !have_obs) {
            err = "set the tables and the observations first";
            return CXB_ERR_STATE;
        }
        CXB_CUDA(cudaSetDevice(device));
        CXB_CUDA(cudaEventRecord(ev0, stream));
        int32_t st = dtype == CXB_F32 ? launch_t<float>() : launch_t<double>();
        if (st) return st;
        CXB_CUDA(cudaEventRecord(ev1, stream));
        CXB_CUDA(cudaGetLastError());
        ran = true;
        return CXB_OK;
    }
};

}  // namespace cxb

using cxb::Hmm;
static inline Hmm* HM(cxb_hmm* m) { return reinterpret_cast<Hmm*>(m); }
#define HM_CUDA(m, expr)                                        \
    do {                                                        \
        cudaError_t e__ = (expr);                               \
        if (e__ != cudaSuccess) {                               \
            HM(m)->err = ::cxb::cuda_msg(e__, #expr);           \
            return CXB_ERR_CUDA;                                \
        }                                                       \
    } while (0)

extern "C" {

int32_t cxb_hmm_create(int32_t device, int32_t dtype, int64_t n_chains, int64_t n_steps, int32_t n_states, int32_t n_symbols,
                       cxb_hmm** out) try {
    if (!out || (dtype != CXB_F32 && dtype != CXB_F64)) return CXB_ERR_BAD_ARG;
    *out = nullptr;
    Hmm* m = new Hmm();
    m->device = device;
    m->dtype = dtype;
    m->B = n_chains;
    m->T = n_steps;
    m->K = n_states;
    m->M = n_symbols;
    int32_t st = m->init();
    if (st) {
        fprintf(stderr, "cxb_hmm_create: %s\n", m->err.c_str());
        delete m;
        return st;
    }
    *out = reinterpret_cast<cxb_hmm*>(m);
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
void cxb_hmm_destroy(cxb_hmm* m) {
    if (m) {
        cudaSetDevice(HM(m)->device);
        delete HM(m);
    }
}
const char* cxb_hmm_last_error(cxb_hmm* m) { return m ? HM(m)->err.c_str() : "null handle"; }
int32_t cxb_hmm_set_tables(cxb_hmm* m, const double* transition, const double* emission) try {
    return HM(m)->set_tables(transition, emission);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_hmm_set_observations(cxb_hmm* m, const uint8_t* obs_host) try {
    Hmm* h = HM(m);
    HM_CUDA(m, cudaSetDevice(h->device));
    HM_CUDA(m, cudaMemcpyAsync(h->obs.p, obs_host, (size_t)h->T * h->B, cudaMemcpyHostToDevice, h->stream));
    HM_CUDA(m, cudaStreamSynchronize(h->stream));
    h->have_obs = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_hmm_update_marginals(cxb_hmm* m, int64_t* n_updates_out) try {
    Hmm* h = HM(m);
    int32_t st = h->launch();
    if (st) return st;
    if (n_updates_out) *n_updates_out = h->B * (6 * h->T - 4);
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
static int32_t hmm_get(cxb_hmm* m, const unsigned char* src, int64_t t0, int64_t t1, void* out_host) {
    Hmm* h = HM(m);
    if (!h->ran || t0 < 0 || t1 > h->T || t0 >= t1) {
        h->err = "bad time slice or no update has run yet";
        return CXB_ERR_BAD_ARG;
    }
    size_t step = (size_t)h->B * h->K * h->esz();
    HM_CUDA(m, cudaSetDevice(h->device));
    HM_CUDA(m, cudaMemcpyAsync(out_host, src + (size_t)t0 * step, (size_t)(t1 - t0) * step, cudaMemcpyDeviceToHost, h->stream));
    HM_CUDA(m, cudaStreamSynchronize(h->stream));
    return CXB_OK;
}
int32_t cxb_hmm_get_marginals(cxb_hmm* m, int64_t t0, int64_t t1, void* out_host) try { return hmm_get(m, HM(m)->marg.p, t0, t1, out_host); } CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_hmm_get_forward(cxb_hmm* m, int64_t t0, int64_t t1, void* out_host) try { return hmm_get(m, HM(m)->fwd.p, t0, t1, out_host); } CXB_ABI_CATCH(CXB_ERR_INTERNAL)
void* cxb_hmm_stream(cxb_hmm* m) { return (void*)HM(m)->stream; }
int32_t cxb_hmm_last_kernel_ms(cxb_hmm* m, float* ms_out) try {
    Hmm* h = HM(m);
    if (!h->ran) {
        h->err = "no update has run yet";
        return CXB_ERR_STATE;
    }
    HM_CUDA(m, cudaEventSynchronize(h->ev1));
    HM_CUDA(m, cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_hmm_sync(cxb_hmm* m) try {
    HM_CUDA(m, cudaStreamSynchronize(HM(m)->stream));
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)

}  // extern "C"
