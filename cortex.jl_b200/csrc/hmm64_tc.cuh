// hmm64_tc.cuh — K = 64, fp32 HMM batch (BASELINE config 3, K = 64) on the warp-level tensor-core path, both recursions at once.
//
// Why this shape. The recursion over t is strictly sequential and the batch is small (1,024 chains), so a time step is
// latency-, not throughput-bound: the FFMA2 kernel (k_hmm64_pass) holds a 64 x 64 table in 128 registers per lane, which
// leaves ONE 7-warp CTA per SM, two passes one after the other and 600+ cycles per step. Here
//   * a CTA owns 8 chains and runs BOTH recursions at the same time: warps 0-3 the forward recursion (time 0 upwards),
//     warps 4-7 the backward recursion (time T-1 downwards) — two independent dependency chains per SM sub-partition
//     instead of one; the table costs 48 registers per thread (its fragments as three bf16 pieces);
//   * one step of one recursion is pred[64 states][8 chains] = M[64][64] . msg[64][8] on mma.sync.m16n8k16 (measured on a
//     B200: 8 cycles per instruction and sub-partition, 21 cycles latency — profiles/r02_mma_rate.log): warp w owns output
//     states 16 w .. 16 w + 15; both operands are split into THREE bf16 pieces (x = p0 + p1 + p2, products p_a q_b with
//     a + b < 3: six MMAs per 16 input states, summed small to large) — fp32-level accuracy, the scheme of hmm_tc.cuh;
//   * the carried message travels between the four warps of a recursion through a double-buffered shared-memory image that
//     is ALREADY the B-fragment layout (six 128-bit loads per lane and step, one named barrier per step and recursion);
//   * the two recursions meet in the middle: until then the forward half stores forward messages and the backward half
//     stores its (scaled) backward predictions into the marginal plane; after ONE block barrier the forward half multiplies
//     the stored predictions in as it goes (times >= T/2) and the backward half the stored forward messages (times < T/2).
//     Traffic: 14 K bytes per (chain, step) instead of 12 K — the kernel is latency-bound, not bandwidth-bound;
//   * scaling and exact normalisation are off the critical path as in k_hmm64_pass: the carried message is scaled by the
//     power of two of the previous step's sum, the exactly normalised rows leave one step late.
// Status: parity-green, measured SLOWER than k_hmm64_pass at config-3 size (1,430 cycles per step of both recursions against
// 2 x 605; profiles/r02_hmm64_tc_ncu.txt), so it runs only with CXB_HMM64_TC=1. Compile-time switches used for that analysis
// (make EXTRA=-D...): CXB_H64_TRACE (clock stamps inside a step, printed by CTA 3), H64_EXP=1|2|3 (ablations: 4 MMAs per
// step instead of 24 / no deferred sums and write-out / both - wrong values, timing only).
// Rules (HMM): m2f(z_t, tr_t) = normalise(em_t * pred_t), marginal = normalise(fwd_t * bwd_t),
// m2f(z_t, tr_{t-1}) = normalise(em_t * bwd_t) — the values k_hmm_pass / k_hmm64_pass produce.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace cxb {
namespace h64 {

constexpr int NB = 8;                   // chains per CTA (MMA N)
constexpr int MSG_BYTES = 2 * 6 * 32 * 16;  // [2 buffers][6 vectors][32 lanes] x 16 bytes: the message as B fragments
constexpr int PART_BYTES = 2 * 2 * NB * 4 * 4;  // [2 buffers][sum of the message, sum of the marginal][chain][warp]
constexpr int OBS_BYTES = 2 * 32 * NB;          // [2 buffers][32 steps][chain]
constexpr int HALF_BYTES = MSG_BYTES + PART_BYTES + OBS_BYTES;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two floats -> one bf16x2 word (first element in the low half, as the MMA fragments want it)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t w;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
    return w;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// (x0, x1) = piece 0 + piece 1 + piece 2 (each piece the bf16 of what the earlier ones left over)
__device__ __forceinline__ void split3(float x0, float x1, uint32_t (&w)[3]) {
    w[0] = pack_bf16(x0, x1);
    x0 -= bf16_lo(w[0]);
    x1 -= bf16_hi(w[0]);
    w[1] = pack_bf16(x0, x1);
    x0 -= bf16_lo(w[1]);
    x1 -= bf16_hi(w[1]);
    w[2] = pack_bf16(x0, x1);
}
__device__ __forceinline__ float pow2_inv(float s) {
    unsigned e = (__float_as_uint(s) >> 23) & 0xffu;
    return __uint_as_float((254u - e) << 23);
}
__device__ __forceinline__ float rcp_nr(float x) {  // MUFU.RCP + one Newton step
    float q;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(x));
    return fmaf(q, fmaf(-x, q, 1.0f), q);
}
#ifdef CXB_H64_TRACE
#define H64_STAMP(k, dep)                                                                   \
    do {                                                                                    \
        long long c__;                                                                      \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(c__) : "f"(dep) : "memory");         \
        tr_acc[k] += c__ - tr_last;                                                         \
        tr_last = c__;                                                                      \
    } while (0)
#else
#define H64_STAMP(k, dep) do { } while (0)
#endif
__device__ __forceinline__ void half_barrier(int half) { asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory"); }

// One recursion (FWD: time ascending, tbl = A; backward: time descending, tbl = A^T; out[j] = sum_i tbl[i][j] msg[i]).
template <bool FWD, bool EM_SMEM>
__device__ __forceinline__ void recursion(const float* __restrict__ tbl, const float* __restrict__ emis_n, const float* sEm, unsigned char* hs,
                                          const uint8_t* __restrict__ obs, float* fwd, float* marg, long long B, long long Tn, int n_sym) {
    constexpr int half = FWD ? 0 : 1;
    const int tid = threadIdx.x & 127, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    uint4* sMsg = reinterpret_cast<uint4*>(hs);
    float* sPart = reinterpret_cast<float*>(hs + MSG_BYTES);
    uint8_t* sObs = hs + MSG_BYTES + PART_BYTES;
    const long long b0 = (long long)blockIdx.x * NB;
    // After the exchange of a step a lane owns FOUR states of ONE chain: chain n, states kb, kb + 1, kb + 8, kb + 9 — two
    // packed pairs = the (b0, b1) words of B-fragment lane (g' = n, t' = g / 2) of K-step `warp`.
    const int n = 2 * t + (g & 1), kb = 16 * warp + (g & ~1);
    const int owner = n * 4 + (g >> 1);
    const bool odd = g & 1;
    const bool live = b0 + n < B;
    const long long Tm = Tn / 2, s_switch = FWD ? Tm : Tn - Tm;
    auto time_of = [&](long long s) -> long long { return FWD ? s : Tn - 1 - s; };

    // table fragments (A operand): rows = output states j0 + g (+ 8), columns = input states 16 kt + 2 t (+ 1, + 8, + 9)
    uint32_t ta[3][4][4];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = 16 * warp + g + 8 * (r & 1), i = 16 * kt + 2 * t + 8 * (r >> 1);
            uint32_t w[3];
            split3(tbl[(size_t)i * 64 + j], tbl[(size_t)(i + 1) * 64 + j], w);
            ta[0][kt][r] = w[0];
            ta[1][kt][r] = w[1];
            ta[2][kt][r] = w[2];
        }

    // 32 steps x 8 chains of symbols, two entries per thread, clamped here; requested and stored 16 steps apart so that
    // nothing waits on the load
    auto obs_request = [&](long long step0, int (&o)[2]) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid * 2 + e, st = idx / NB, nn = idx % NB;
            const long long step = step0 + st;
            o[e] = 0;
            if (step < Tn && b0 + nn < B) o[e] = obs[(size_t)time_of(step) * B + b0 + nn];
        }
    };
    auto obs_store = [&](const int (&o)[2], int buf) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid * 2 + e, st = idx / NB, nn = idx % NB;
            sObs[(buf * 32 + st) * NB + nn] = (uint8_t)min(o[e], n_sym - 1);
        }
    };
    auto em_of = [&](long long s, float (&e4)[4]) {  // emission message of step s for this lane's chain and states
        const int o = sObs[((int)((s >> 5) & 1) * 32 + (int)(s & 31)) * NB + n];
        float2 x, y;
        if (EM_SMEM) {
            x = *reinterpret_cast<const float2*>(sEm + o * 64 + kb);
            y = *reinterpret_cast<const float2*>(sEm + o * 64 + kb + 8);
        } else {
            x = __ldg(reinterpret_cast<const float2*>(emis_n + (size_t)o * 64 + kb));
            y = __ldg(reinterpret_cast<const float2*>(emis_n + (size_t)o * 64 + kb + 8));
        }
        e4[0] = x.x, e4[1] = x.y, e4[2] = y.x, e4[3] = y.y;
    };
    auto cell = [&](float* plane, long long s) -> float* { return plane + ((size_t)time_of(s) * B + (size_t)(b0 + n)) * 64 + kb; };
    auto load_cells = [&](const float* plane, long long s, float (&x)[4]) {  // written by the other half of this CTA: bypass L1
        x[0] = x[1] = x[2] = x[3] = 0.0f;
        if (s < Tn && live) {
            const float* p = cell(const_cast<float*>(plane), s);
            const float2 lo = __ldcg(reinterpret_cast<const float2*>(p)), hi = __ldcg(reinterpret_cast<const float2*>(p + 8));
            x[0] = lo.x, x[1] = lo.y, x[2] = hi.x, x[3] = hi.y;
        }
    };
    auto store_cells = [&](float* plane, long long s, const float (&x)[4], float q) {
        if (!live) return;
        float* p = cell(plane, s);
        __stcs(reinterpret_cast<float2*>(p), make_float2(x[0] * q, x[1] * q));
        __stcs(reinterpret_cast<float2*>(p + 8), make_float2(x[2] * q, x[3] * q));
    };

    // Pipeline of a step s (everything but the matrix product, the exchange and the split is off the critical path):
    //   * u1 / g1 = message / marginal product of step s-1, u2 / g2 of step s-2 (registers);
    //   * while the MMAs of step s run: the sums of step s-1 are reduced (two shuffles) and stored for the other warps,
    //     the totals of step s-2 are read, its rows leave exactly normalised, and the scale of this step is derived from
    //     them: r = 2^-(e / 2), e = exponent of sum(u_{s-2}) — a two-step-old sum with half gain (the plain 2^-e of the
    //     one-step-old sum would put the reduction on the critical path; with a two-step delay full gain oscillates,
    //     half gain is damped: e_s = e_{s-1} - e_{s-2} / 2 has both roots at |z| = 0.71).
    float u1[4] = {0, 0, 0, 0}, g1[4] = {0, 0, 0, 0}, u2[4] = {0, 0, 0, 0}, g2[4] = {0, 0, 0, 0};
    // the other half's rows of the steps ahead: a ring of RING steps, indexed statically inside the unrolled loop (a row is
    // consumed RING steps after its load was issued; rotating the registers instead would wait for every load at once)
    constexpr int RING = 4;
    float em_cur[4], x_ring[RING][4];
    int obs_hold[2] = {0, 0};  // symbols of the next block of 32 steps between their load and their store to shared memory
    // the other recursion's stored rows of the times this half visits after the switch: FWD reads the backward predictions
    // (marginal plane), the backward half reads the forward messages
    const float* xplane = FWD ? marg : fwd;
    const int T32 = (int)Tn, sw = (int)s_switch;

    // partial sums of the rows (u, g) of step sp over this warp's 16 states -> buffer sp & 1
    auto reduce_store = [&](int sp, bool p2, const float (&u)[4], const float (&gg)[4]) {
        float su = (u[0] + u[1]) + (u[2] + u[3]), sg = (gg[0] + gg[1]) + (gg[2] + gg[3]);
        su += __shfl_xor_sync(0xffffffffu, su, 8);
        if (p2) sg += __shfl_xor_sync(0xffffffffu, sg, 8);
        su += __shfl_xor_sync(0xffffffffu, su, 16);
        if (p2) sg += __shfl_xor_sync(0xffffffffu, sg, 16);
        if (lane < 8) {
            sPart[(size_t)(((sp & 1) * 2 + 0) * NB + n) * 4 + warp] = su;
            if (p2) sPart[(size_t)(((sp & 1) * 2 + 1) * NB + n) * 4 + warp] = sg;
        }
    };
    // exactly normalised rows of step sp (its partial sums are in buffer sp & 1 since the barrier after they were stored)
    auto write_out = [&](int sp, bool p2, bool enabled, const float (&u)[4], const float (&gg)[4]) {
        const float* pp = sPart + (size_t)((sp & 1) * 2) * NB * 4;
        if (FWD) {
            const float4 pu = *reinterpret_cast<const float4*>(pp + n * 4);
            const float q = rcp_nr((pu.x + pu.y) + (pu.z + pu.w));
            if (enabled) store_cells(fwd, sp, u, q);
        }
        if (p2) {
            const float4 pg = *reinterpret_cast<const float4*>(pp + (NB + n) * 4);
            const float q = rcp_nr((pg.x + pg.y) + (pg.z + pg.w));
            if (enabled) store_cells(marg, sp, gg, q);
        }
    };
    // end of a phase [s_begin, s_end): the rows of its last two steps leave
    auto finish_phase = [&](int s_begin, int s_end, bool p2) {
        if (s_end <= s_begin) return;
        reduce_store(s_end - 1, p2, u1, g1);
        write_out(s_end - 2, p2, s_end - 2 >= s_begin, u2, g2);
        half_barrier(half);
        write_out(s_end - 1, p2, true, u1, g1);
    };

#ifdef CXB_H64_TRACE
    long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tr_last = clock64();
#endif
    // FIRST: step 0 of the recursion (no prediction yet). phase_begin: first step of the current phase (rows of earlier
    // steps have left already).
    auto step = [&](auto p2_tag, auto slot_tag, auto first_tag, int s, int phase_begin) {
        constexpr bool P2 = decltype(p2_tag)::value, FIRST = decltype(first_tag)::value;
        constexpr int SLOT = decltype(slot_tag)::value;
        H64_STAMP(0, 0.0f);
        const int buf = s & 1;
        if ((s & 31) == 0) obs_request(s + 32, obs_hold);
        if ((s & 31) == 16) obs_store(obs_hold, ((s >> 5) + 1) & 1);
        float em_nxt[4];
        em_of(s + 1, em_nxt);
        float x_cur[4] = {0, 0, 0, 0};
        if (P2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) x_cur[i] = x_ring[SLOT][i];
            load_cells(xplane, s + RING, x_ring[SLOT]);
            if (s + 24 < T32 && tid < 2 * NB && b0 + (tid >> 1) < B)  // 8 chains x 256 bytes of step s + 24 into L2
                asm volatile("prefetch.global.L2 [%0];" ::"l"(xplane + ((size_t)time_of(s + 24) * B + (size_t)(b0 + (tid >> 1))) * 64 + 32 * (tid & 1)));
        }
        float d[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        if (!FIRST) {
            // message of step s-1 as B fragments: vector q = 2 piece + kt / 2 holds (b0, b1) of K-steps kt = 2 (q & 1), + 1
            uint4 bq[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) bq[q] = sMsg[(buf * 6 + q) * 32 + lane];
            float hh[4] = {0, 0, 0, 0}, c1a[4] = {0, 0, 0, 0}, c1b[4] = {0, 0, 0, 0}, c2a[4] = {0, 0, 0, 0}, c2b[4] = {0, 0, 0, 0},
                  c2c[4] = {0, 0, 0, 0};
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) {
                uint32_t v0[2], v1[2], v2[2];
                const uint4 q0 = bq[0 + (kt >> 1)], q1 = bq[2 + (kt >> 1)], q2 = bq[4 + (kt >> 1)];
                v0[0] = (kt & 1) ? q0.z : q0.x, v0[1] = (kt & 1) ? q0.w : q0.y;
                v1[0] = (kt & 1) ? q1.z : q1.x, v1[1] = (kt & 1) ? q1.w : q1.y;
                v2[0] = (kt & 1) ? q2.z : q2.x, v2[1] = (kt & 1) ? q2.w : q2.y;
                mma16816(hh, ta[0][kt], v0[0], v0[1]);
#if defined(H64_EXP) && (H64_EXP & 1)
                continue;
#endif
                mma16816(c1a, ta[0][kt], v1[0], v1[1]);
                mma16816(c1b, ta[1][kt], v0[0], v0[1]);
                mma16816(c2a, ta[0][kt], v2[0], v2[1]);
                mma16816(c2b, ta[1][kt], v1[0], v1[1]);
                mma16816(c2c, ta[2][kt], v0[0], v0[1]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = (((c2a[i] + c2b[i]) + c2c[i]) + (c1a[i] + c1b[i])) + hh[i];
        }
        // ---- in the shadow of the MMAs: sums of step s-1, rows of step s-2, this step's scale ----
        float r = 1.0f;
#if defined(H64_EXP) && (H64_EXP & 2)
        if (false) {
#else
        if (!FIRST) {
#endif
            reduce_store(s - 1, P2, u1, g1);
            const float4 pu = *reinterpret_cast<const float4*>(sPart + (size_t)((buf * 2) * NB + n) * 4);  // totals of step s-2
            const float tot = (pu.x + pu.y) + (pu.z + pu.w);
            const int e = (int)((__float_as_uint(tot) >> 23) & 0xffu) - 127;
            if (s >= 2) r = __uint_as_float((uint32_t)(127 - (e >> 1)) << 23);
            write_out(s - 2, P2, s - 2 >= phase_begin, u2, g2);
        }
        H64_STAMP(1, r);
        H64_STAMP(2, d[0] + d[1] + d[2] + d[3]);
        float p4[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        if (!FIRST) {
            // D fragment: d0 (state g, chain 2t), d1 (g, 2t+1), d2 (g+8, 2t), d3 (g+8, 2t+1). Lanes g and g ^ 1 swap the
            // chain they do not keep: even g keeps chain 2t, odd g chain 2t+1, each with states (g & ~1) + {0, 1, 8, 9}.
            const float r0 = __shfl_xor_sync(0xffffffffu, odd ? d[0] : d[1], 4);
            const float r1 = __shfl_xor_sync(0xffffffffu, odd ? d[2] : d[3], 4);
            p4[0] = (odd ? r0 : d[0]) * r;
            p4[1] = (odd ? d[1] : r0) * r;
            p4[2] = (odd ? r1 : d[2]) * r;
            p4[3] = (odd ? d[3] : r1) * r;
        }
        float u4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) u4[i] = em_cur[i] * p4[i];
        H64_STAMP(3, u4[0] + u4[1] + u4[2] + u4[3]);
        {   // the carried message of the next step, as the (b0, b1) words of its three pieces
            uint32_t wlo[3], whi[3];
            split3(u4[0], u4[1], wlo);
            split3(u4[2], u4[3], whi);
            uint2* dst = reinterpret_cast<uint2*>(sMsg);
#pragma unroll
            for (int p = 0; p < 3; ++p)
                dst[((size_t)((buf ^ 1) * 6 + 2 * p + (warp >> 1)) * 32 + owner) * 2 + (warp & 1)] = make_uint2(wlo[p], whi[p]);
        }
        H64_STAMP(4, 0.0f);
        if (!FWD && !P2) store_cells(marg, s, p4, 1.0f);  // the backward prediction of this time, for the forward half to pick up
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            u2[i] = u1[i];
            g2[i] = g1[i];
            u1[i] = u4[i];
            g1[i] = P2 ? (FWD ? u4[i] * x_cur[i] : x_cur[i] * p4[i]) : 0.0f;
            em_cur[i] = em_nxt[i];
        }
        H64_STAMP(6, 0.0f);
        half_barrier(half);
        H64_STAMP(7, 0.0f);
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    using I3 = std::integral_constant<int, 3>;

    obs_request(0, obs_hold);
    obs_store(obs_hold, 0);
    half_barrier(half);
    em_of(0, em_cur);
    if (sw > 0) step(std::false_type{}, I0{}, std::true_type{}, 0, 0);
    for (int s = 1; s < sw; ++s) step(std::false_type{}, I0{}, std::false_type{}, s, 0);
    finish_phase(0, sw, false);
    // every row the other half needs is stored: one block barrier, then the second halves
    __threadfence_block();
    asm volatile("bar.sync 0, 256;" ::: "memory");
#pragma unroll
    for (int u = 0; u < RING; ++u) load_cells(xplane, sw + u, x_ring[u]);
    if (sw == 0) {  // (T = 1: the forward half starts here)
        step(std::true_type{}, I0{}, std::true_type{}, 0, 0);
        if (1 < T32) step(std::true_type{}, I1{}, std::false_type{}, 1, 0);
        if (2 < T32) step(std::true_type{}, I2{}, std::false_type{}, 2, 0);
        if (3 < T32) step(std::true_type{}, I3{}, std::false_type{}, 3, 0);
    }
    for (int s0 = sw == 0 ? RING : sw; s0 < T32; s0 += RING) {
        if (s0 + 0 < T32) step(std::true_type{}, I0{}, std::false_type{}, s0 + 0, sw);
        if (s0 + 1 < T32) step(std::true_type{}, I1{}, std::false_type{}, s0 + 1, sw);
        if (s0 + 2 < T32) step(std::true_type{}, I2{}, std::false_type{}, s0 + 2, sw);
        if (s0 + 3 < T32) step(std::true_type{}, I3{}, std::false_type{}, s0 + 3, sw);
    }
    finish_phase(sw, T32, true);
#ifdef CXB_H64_TRACE
    if (blockIdx.x == 3 && (tid & 31) == 0)
        printf("h64 trace %s warp %d: top %lld | shadow issued %lld | mma done %lld | exchange+em %lld | split+sts %lld | - %lld | tail %lld | barrier %lld (cycles per step)\n",
               FWD ? "fwd" : "bwd", warp, tr_acc[0] / Tn, tr_acc[1] / Tn, tr_acc[2] / Tn, tr_acc[3] / Tn, tr_acc[4] / Tn, tr_acc[5] / Tn, tr_acc[6] / Tn, tr_acc[7] / Tn);
#endif
}

template <bool EM_SMEM>
__global__ void __launch_bounds__(256, 1)
k_hmm64_tc(const float* __restrict__ A, const float* __restrict__ At, const float* __restrict__ emis_n, const uint8_t* __restrict__ obs,
           float* fwd, float* marg, long long B, long long Tn, int n_sym) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sEm = reinterpret_cast<float*>(smem_raw + 2 * HALF_BYTES);
    if (EM_SMEM) {
        for (int x = threadIdx.x; x < n_sym * 64; x += blockDim.x) sEm[x] = emis_n[x];
        __syncthreads();
    }
    if (threadIdx.x < 128)
        recursion<true, EM_SMEM>(A, emis_n, sEm, smem_raw, obs, fwd, marg, B, Tn, n_sym);
    else
        recursion<false, EM_SMEM>(At, emis_n, sEm, smem_raw + HALF_BYTES, obs, fwd, marg, B, Tn, n_sym);
}

inline size_t smem_bytes(int n_sym, bool em_smem) { return (size_t)2 * HALF_BYTES + (em_smem ? (size_t)n_sym * 64 * sizeof(float) : 0); }

}  // namespace h64
}  // namespace cxb
