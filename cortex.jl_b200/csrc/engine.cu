// engine.cu — the generic device engine: Cortex.jl's Signal DAG + update_marginals! as a
// device-resident CSR with a level-synchronous frontier.
//
// What replaces what (reference paths relative to Cortex.jl v0.3.0):
//   Signal.props / SignalDependenciesProps (src/signal.jl:36-51)   -> props[N] bytes + packed 64-bit nibble chunks
//   is_pending / is_meeting_pending_criteria (:141-154, :668-730)  -> pending_eval() (same lazy two-flag protocol)
//   set_value! + notify_listener! (:232-253, :339-356)             -> k_apply (listener CSR with precomputed first slot)
//   process_dependencies! DFS (:466-490)                           -> k_bfs level-synchronous reachability
//   process! rule dispatch (src/inference_engine.jl:479-509)       -> frontier multisplit by rule key + one batched
//                                                                     kernel per (signal kind, factor type)
//   update_marginals! (:559-632)                                   -> DeviceEngine::update()
// No CPU fallback: every compute entry point needs the CUDA device.
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>

#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "host_graph.hpp"

namespace cxb {

unsigned long long g_kernel_launches = 0;

constexpr uint64_t ALL_W = 0x2222222222222222ull, ALL_C = 0x4444444444444444ull, ALL_F = 0x8888888888888888ull,
                   PASS = 0x1111111111111111ull;
constexpr int ERR_INDEPENDENCE = 1, ERR_LISTENER_DONE = 2, ERR_RULE_ARG = 4, ERR_WEAK_BENEATH_DONE = 16, ERR_LEFTOVER_FRESH = 32;  // 8 = ERR_NO_RULE_KEY
// strict rules of the level schedule (the oracle's update_lvl states them in the same words)
constexpr int ERR_STALE_BENEATH = 64;          // A: a visited, non-pending signal holds leftover freshness (first level)
constexpr int ERR_WORK_BENEATH_PENDING = 128;  // B: a marginal already pending at request time has pending work beneath it
constexpr int ERR_REVISITED = 256;             // D: a member found pending more than once, once through an intermediate slot, with a pending dependency
constexpr int ERR_NONLISTEN = 512;             // E: non-listening notifications (order inside a level / a request matters)
constexpr int ERR_FINAL_ORDER = 1024;          // F: final-phase candidates that depend on each other across the per-variable order
constexpr int ERR_SEQ_OVERFLOW = 2048;         // sequential executor: traversal deeper than the stack (a dependency cycle)
constexpr int ERR_SEQ_DIVERGED = 4096;         // sequential executor: the reference loop does not terminate on this request
constexpr uint32_t PROBE_BIT = 0x80000000u;  // breadth-first list entry: the signal is only probed (see bfs_visit)
constexpr uint32_t MULTI_BIT = 0x40000000u;  // breadth-first list entry: the reference's depth-first traversal reaches the signal more than once
constexpr uint32_t ENTRY_ID = 0x3FFFFFFFu;
constexpr uint32_t MK_VIS2 = 1, MK_VIA_I = 2, MK_MULTI = 4;  // View::mark2 flags (low 3 bits, level epoch above)
constexpr int KEY_COMBINE = 0;  // dense rule keys: 0 = family reduce, 1..n_types = m2v of a factor type, n_types+1 = no rule


// device view of the engine state (plain pointers, passed by value to kernels)
struct View {
    uint32_t n_sig;
    int dim;
    const uint32_t *dep_off, *dep_ids, *nib_off;
    uint64_t* nib;
    const uint32_t *lis_off, *lis_ids, *lis_slot;
    const uint8_t* lis_listen;
    uint8_t* props;
    const uint8_t *kind, *rkey;
    const int32_t *svar, *sfac;
    const uint8_t* sfam;              // per-signal value family (cxb_set_variable_families), nullptr = the engine family
    uint32_t *done_epoch, *visit_epoch;
    uint32_t* probe_epoch;            // graphs with weak dependencies only (else nullptr): probe visits of this level
    uint32_t* mark2;                  // (lvl_epoch << 3) | MK_*: visit multiplicities of the current level (strict rule D)
    uint32_t* nl_epoch;               // graphs with non-listening dependencies only (else nullptr): req_epoch of the last such notification
    uint32_t* nl_list;                // ... signals that received one during the current level, count in *nl_cnt
    uint32_t* nl_cnt;
    int strict;                       // strict rules A / B / D / E armed
    uint32_t* front_epoch;            // == lvl_epoch: member of the current level's frontier
    uint32_t* front;                  // frontier buffer, partitioned by rule key
    const uint32_t* key_base;         // [n_keys] start of each key's partition
    uint32_t* key_cnt;                // [n_keys] cursors
    int use_keys;
    int* err_flag;
    unsigned long long* kind_count;  // [6]
};

// ---- is_meeting_pending_criteria, src/signal.jl:668-730 (same chunk arithmetic) -------------------------
__device__ __forceinline__ bool criteria(const View& e, uint32_t s) {
    uint32_t nd = e.dep_off[s + 1] - e.dep_off[s];
    if (nd == 0) return false;
    uint32_t noff = e.nib_off[s], nch = e.nib_off[s + 1] - noff;
    for (uint32_t i = 0; i + 1 < nch; ++i) {
        uint64_t c = e.nib[noff + i];
        uint64_t W = (c & ALL_W) >> 1, C = (c & ALL_C) >> 2, F = (c & ALL_F) >> 3;
        if ((C & (W | F)) != PASS) return false;
    }
    uint32_t last = nd - 1;
    int shift = (int)((last & 15) << 2) + 4;
    uint64_t pad = shift >= 64 ? 0ull : (~0ull << shift);
    uint64_t c = e.nib[noff + (last >> 4)] | pad;
    uint64_t W = (c & ALL_W) >> 1, C = (c & ALL_C) >> 2, F = (c & ALL_F) >> 3;
    return (C & (W | F)) == PASS;
}
// is_pending, src/signal.jl:141-154. Concurrent evaluations of one signal are idempotent inside a level
// (no set_value! happens between them), so racing threads write the same byte.
__device__ __forceinline__ bool pending_eval(const View& e, uint32_t s) {
    uint8_t p = e.props[s];
    if (p & P_P) return true;
    if (p & P_PP) {
        bool r = criteria(e, s);
        e.props[s] = (uint8_t)((p & P_COMPUTED) | (r ? P_P : 0));
        return r;
    }
    return false;
}

// request_inference_for, src/inference_engine.jl:305-318: flag the direct dependencies of every requested
// marginal and its linked signals (is_potentially_pending = true, is_pending = false). One thread per requested
// marginal, then one thread per linked-signal entry (flat: hub variables link thousands of signals).
__global__ void k_request(View e, const uint32_t* req_marg, const uint32_t* link_ids, uint32_t n, uint32_t n_links) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint32_t m = req_marg[i];
        for (uint32_t k = e.dep_off[m]; k < e.dep_off[m + 1]; ++k) {
            uint32_t d = e.dep_ids[k];
            e.props[d] = (uint8_t)((e.props[d] & P_COMPUTED) | P_PP);
        }
    } else if (i < n + n_links) {
        uint32_t d = link_ids[i - n];
        e.props[d] = (uint8_t)((e.props[d] & P_COMPUTED) | P_PP);
    }
}

// Contract check at request time (after k_request): a requested marginal that is not pending must not hold a FRESH bit on a
// computed, non-input dependency - leftover freshness of an earlier request that could not complete makes the reference's
// answer depend on the order in which it visits the variables (see the oracle's update_lvl). Read-only.
__device__ __forceinline__ bool pending_now(const View& e, uint32_t s) {  // is_pending without the caching side effect
    const uint8_t p = e.props[s];
    return (p & P_P) || ((p & P_PP) && criteria(e, s));
}
// pend_at_req[i] = the marginal was already pending when the request arrived (strict rule B); *n_pend counts them
__global__ void k_request_check(View e, const uint32_t* req_marg, uint32_t n, uint8_t* pend_at_req, int* n_pend) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = req_marg[i];
    const uint8_t p = e.props[m];
    if ((p & P_P) || ((p & P_PP) && criteria(e, m))) {
        pend_at_req[i] = 1;
        atomicAdd(n_pend, 1);
        return;
    }
    pend_at_req[i] = 0;
    const uint32_t off = e.dep_off[m], nd = e.dep_off[m + 1] - off, noff = e.nib_off[m];
    for (uint32_t k = 0; k < nd; ++k) {
        const uint32_t nibble = (uint32_t)(e.nib[noff + (k >> 4)] >> ((k & 15) << 2)) & 0xF;
        const uint32_t d = e.dep_ids[off + k];
        if ((nibble & CXB_NIB_FRESH) && e.dep_off[d + 1] > e.dep_off[d]) atomicOr(e.err_flag, ERR_LEFTOVER_FRESH);
    }
}

// seeds of one level: marginals of the requested variables that are not ready yet (:585)
__global__ void k_seeds(const uint32_t* req_marg, const uint8_t* sel, uint8_t want, uint32_t n, uint32_t* out, uint32_t* n_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool take = i < n && sel[i] == want;
    unsigned m = __ballot_sync(0xffffffffu, take);
    if (!m) return;
    int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(n_out, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (take) out[base + __popc(m & ((1u << lane) - 1))] = req_marg[i];
}

// Append a pending signal to the frontier of this level, grouped by rule key: the frontier buffer is partitioned
// statically by key (key_base[k] = number of signals with a smaller key), the per-key cursors are bumped with
// warp-aggregated atomics (one atomic per distinct key per warp: __match_any over the active lanes).
__device__ __forceinline__ uint32_t mark_or(uint32_t* m, uint32_t lvl, uint32_t f) {  // returns the flags the level had before
    uint32_t old = *m, assumed;
    do {
        assumed = old;
        const uint32_t cur = (assumed >> 3) == lvl ? assumed : (lvl << 3);
        if ((cur & f) == f && (assumed >> 3) == lvl) return cur & 7u;
        old = atomicCAS(m, assumed, cur | f);
    } while (old != assumed);
    return (assumed >> 3) == lvl ? (assumed & 7u) : 0u;
}
__device__ __forceinline__ uint32_t mark_get(const uint32_t* m, uint32_t lvl) {
    const uint32_t v = *m;
    return (v >> 3) == lvl ? (v & 7u) : 0u;
}
// leftover freshness (strict rule A; the oracle's holds_leftover_freshness): FRESH on a strong dependency that is not an input
__device__ __forceinline__ bool holds_leftover(const View& e, uint32_t s) {
    const uint32_t off = e.dep_off[s], nd = e.dep_off[s + 1] - off, noff = e.nib_off[s];
    for (uint32_t k = 0; k < nd; ++k) {
        const uint32_t nibble = (uint32_t)(e.nib[noff + (k >> 4)] >> ((k & 15) << 2)) & 0xF;
        if ((nibble & CXB_NIB_FRESH) && !(nibble & CXB_NIB_WEAK)) {
            const uint32_t d = e.dep_ids[off + k];
            if (e.dep_off[d + 1] > e.dep_off[d]) return true;
        }
    }
    return false;
}
__device__ __forceinline__ bool frontier_push(const View& e, uint32_t d, uint32_t lvl_epoch) {
    if (atomicExch(&e.front_epoch[d], lvl_epoch) == lvl_epoch) return false;  // already in this level's frontier
    int key = e.use_keys ? e.rkey[d] : 0;
    unsigned act = __activemask();
    unsigned same = __match_any_sync(act, key);
    int lane = threadIdx.x & 31, leader = __ffs(same) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&e.key_cnt[key], (uint32_t)__popc(same));
    base = __shfl_sync(same, base, leader);
    e.front[e.key_base[key] + base + __popc(same & ((1u << lane) - 1))] = d;
    return true;
}

// One breadth-first step of process_dependencies! (src/signal.jl:466-490) over a list of signals:
// visit every dependency (f(dep) = is_pending), push the pending ones to the frontier, descend through non-pending
// *intermediate* ones. `done` signals are neither reported nor descended through (A.5). The list length lives on
// the device (written by the previous step), so consecutive steps need no host round trip.
// The dependencies of one list entry. A level never descends through a signal already computed in this request (`done`,
// A.5) — the reference does when a later round reaches it again, and anything pending it finds there is computed and
// the done signal recomputed on the retry. With strong dependencies that cannot find anything (a done signal consumed
// fresh dependencies; a recomputed dependency is caught by ERR_LISTENER_DONE). A WEAK dependency never blocks, so it
// can be pending underneath a done signal: graphs that have weak dependencies are therefore PROBED through their done
// signals (same descent rule, nothing is pushed), and a pending weak dependency found there refuses the request as
// order-dependent instead of silently leaving it for the final phase.
// Strict rules evaluated here (use_done bit 1 = first level of a request): A - a visited signal that is not pending but holds
// leftover freshness; D - visit multiplicities. The reference's traversal is depth-first WITHOUT a visited set: a signal
// reached along two paths is traversed twice and everything beneath it is visited twice. A breadth-first traversal visits
// each signal once, so the multiplicity ("more than once") is propagated instead: the second arrival at a traversed signal
// re-queues it once more with MULTI_BIT, and a traversal carrying MULTI_BIT counts double for everything it visits.
template <class Push>
__device__ __forceinline__ void bfs_visit(const View& e, uint32_t entry, uint32_t lvl_epoch, uint32_t req_epoch, int use_done, Push&& push_out) {
    const uint32_t s = entry & ENTRY_ID;
    const bool probing = (entry & PROBE_BIT) != 0, multi = (entry & MULTI_BIT) != 0;
    const bool first_level = (use_done & 2) != 0;
    const bool skip_done = (use_done & 1) != 0;
    const uint32_t off = e.dep_off[s], nd = e.dep_off[s + 1] - off, noff = e.nib_off[s];
    for (uint32_t k = 0; k < nd; ++k) {
        const uint32_t d = e.dep_ids[off + k];
        const bool done = skip_done && e.done_epoch[d] == req_epoch;
        if (done && !e.probe_epoch) continue;
        const uint32_t nibble = (uint32_t)(e.nib[noff + (k >> 4)] >> ((k & 15) << 2)) & 0xF;
        if (done || probing) {
            if (!done && pending_eval(e, d)) {
                if (nibble & CXB_NIB_WEAK) atomicOr(e.err_flag, ERR_WEAK_BENEATH_DONE);
            } else if ((nibble & CXB_NIB_INTERMEDIATE) && atomicExch(&e.probe_epoch[d], lvl_epoch) != lvl_epoch) {
                push_out(d | PROBE_BIT);
            }
        } else if (pending_eval(e, d)) {
            const bool first = frontier_push(e, d, lvl_epoch);
            if (e.strict) {
                const uint32_t f = ((nibble & CXB_NIB_INTERMEDIATE) ? MK_VIA_I : 0u) | ((!first || multi) ? MK_VIS2 : 0u);
                if (f) mark_or(&e.mark2[d], lvl_epoch, f);
            }
        } else {
            if (e.strict && first_level && holds_leftover(e, d)) atomicOr(e.err_flag, ERR_STALE_BENEATH);
            if (nibble & CXB_NIB_INTERMEDIATE) {
                if (atomicExch(&e.visit_epoch[d], lvl_epoch) != lvl_epoch) {
                    if (multi && e.strict) mark_or(&e.mark2[d], lvl_epoch, MK_MULTI);
                    push_out(d | (multi && e.strict ? MULTI_BIT : 0u));
                } else if (e.strict && !(mark_or(&e.mark2[d], lvl_epoch, MK_MULTI) & MK_MULTI)) {
                    push_out(d | MULTI_BIT);  // second arrival: everything beneath is visited (at least) twice
                }
            }
        }
    }
}
__global__ void k_bfs(View e, const uint32_t* in, const uint32_t* n_in_ptr, uint32_t* out, uint32_t* n_out, uint32_t lvl_epoch,
                      uint32_t req_epoch, int use_done) {
    const uint32_t n_in = *n_in_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x)
        bfs_visit(e, in[i], lvl_epoch, req_epoch, use_done, [&](uint32_t x) { out[atomicAdd(n_out, 1u)] = x; });
}

// final phase candidates (:610-628): the requested marginals (i < n), their linked signals (flat entries after)
__global__ void k_final_flags(View e, const uint32_t* req_marg, const uint32_t* link_ids, uint32_t n, uint32_t n_links, int mode,
                              uint32_t lvl_epoch) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode == 0) {
        if (i < n && pending_eval(e, req_marg[i])) frontier_push(e, req_marg[i], lvl_epoch);
    } else {
        if (i < n_links && pending_eval(e, link_ids[i])) frontier_push(e, link_ids[i], lvl_epoch);
    }
}

// readiness, :593-595
__global__ void k_ready(View e, const uint32_t* req_marg, uint8_t* ready, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || ready[i]) return;
    if (pending_eval(e, req_marg[i])) ready[i] = 1;
}

// the frontier of a level = up to 32 per-key segments of the frontier buffer
struct Segs {
    int n;
    uint32_t total;
    uint32_t base[32], cnt[32];
};
__device__ __forceinline__ uint32_t seg_member(const Segs& sg, const uint32_t* front, uint32_t i) {
    int k = 0;
    while (k < sg.n - 1 && i >= sg.cnt[k]) i -= sg.cnt[k++];
    return front[sg.base[k] + i];
}

// ---- set_value! side effects for the members of a level, src/signal.jl:232-253 + 339-356 -------------------
// (values are already written). check_mode: 0 none (user set_value!), 1 loop phase (done-listener contract),
// 2 final phase.  Listener lists of hubs are long (a top ProductOfMessages node of a degree-d variable is a
// dependency of d/2 m2f signals): members with more than 8 listeners are handled by the whole warp.
__device__ __forceinline__ void notify(const View& e, uint32_t k, uint32_t req_epoch, int check_mode) {
    uint32_t L = e.lis_ids[k], slot = e.lis_slot[k];
    if (e.lis_listen[k]) e.props[L] = (uint8_t)((e.props[L] & P_COMPUTED) | P_PP);
    atomicOr((unsigned long long*)&e.nib[e.nib_off[L] + (slot >> 4)],
             (unsigned long long)(CXB_NIB_COMPUTED | CXB_NIB_FRESH) << ((slot & 15) << 2));
    if (check_mode == 1 && e.done_epoch[L] == req_epoch) atomicOr(e.err_flag, ERR_LISTENER_DONE);
    if (check_mode == 1 && e.nl_epoch && !e.lis_listen[k]) {  // strict rule E: remember non-listening notifications
        e.nl_epoch[L] = req_epoch;
        e.nl_list[atomicAdd(e.nl_cnt, 1u)] = L;
    }
}
__global__ void k_apply(View e, Segs sg, uint32_t req_epoch, int check_mode) {
    __shared__ unsigned int kc[6];
    if (*(volatile int*)e.err_flag) return;  // the level was refused by its checks: nothing is applied
    if (threadIdx.x < 6) kc[threadIdx.x] = 0;
    __syncthreads();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < sg.total;
    uint32_t lo = 0, hi = 0;
    if (active) {
        uint32_t s = seg_member(sg, e.front, i);
        for (uint32_t c = e.nib_off[s]; c < e.nib_off[s + 1]; ++c) e.nib[c] &= ~ALL_F;  // unset_all_dependencies_fresh!
        e.props[s] = P_COMPUTED;                                                         // (pp, p) = (false, false)
        if (check_mode) e.done_epoch[s] = req_epoch;
        atomicAdd(&kc[e.kind[s]], 1u);
        lo = e.lis_off[s];
        hi = e.lis_off[s + 1];
    }
    const bool heavy_me = hi - lo > 8;
    if (!heavy_me)
        for (uint32_t k = lo; k < hi; ++k) notify(e, k, req_epoch, check_mode);
    unsigned heavy = __ballot_sync(0xffffffffu, heavy_me);
    const int lane = threadIdx.x & 31;
    while (heavy) {
        int src = __ffs(heavy) - 1;
        heavy &= heavy - 1;
        uint32_t hlo = __shfl_sync(0xffffffffu, lo, src), hhi = __shfl_sync(0xffffffffu, hi, src);
        for (uint32_t k = hlo + lane; k < hhi; k += 32) notify(e, k, req_epoch, check_mode);
    }
    __syncthreads();
    if (threadIdx.x < 6 && kc[threadIdx.x]) atomicAdd(&e.kind_count[threadIdx.x], (unsigned long long)kc[threadIdx.x]);
}
// set_value! side effects for an explicit list (user set_value! / compute!)
__global__ void k_apply_list(View e, const uint32_t* list, uint32_t n, uint32_t req_epoch) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = list[i];
    for (uint32_t c = e.nib_off[s]; c < e.nib_off[s + 1]; ++c) e.nib[c] &= ~ALL_F;
    e.props[s] = P_COMPUTED;
    for (uint32_t k = e.lis_off[s]; k < e.lis_off[s + 1]; ++k) notify(e, k, req_epoch, 0);
}
// independence of a level (A.5): no member may depend on another member. Members are recognisable by
// front_epoch == lvl_epoch (set when they were pushed); runs BEFORE the rules.
// The other contract checks of a level live here too, so that a refused level has not been applied (the rule and
// set_value! kernels of the level return at once when the error flag is up): the listener-done rule (check_mode 1: no
// member may have a listener already computed in this request), strict rule D and the first half of strict rule E.
__device__ __forceinline__ void check_member(const View& e, uint32_t s, uint32_t lvl_epoch, uint32_t req_epoch, int check_mode) {
    int err = 0;
    const uint32_t off = e.dep_off[s], end = e.dep_off[s + 1];
    for (uint32_t k = off; k < end; ++k)
        if (e.front_epoch[e.dep_ids[k]] == lvl_epoch) err |= ERR_INDEPENDENCE;
    if (check_mode == 1) {
        for (uint32_t k = e.lis_off[s]; k < e.lis_off[s + 1]; ++k)
            if (e.done_epoch[e.lis_ids[k]] == req_epoch) err |= ERR_LISTENER_DONE;
        if (e.strict) {
            const uint32_t mk = mark_get(&e.mark2[s], lvl_epoch);
            if ((mk & MK_VIS2) && (mk & MK_VIA_I))
                for (uint32_t k = off; k < end; ++k)
                    if (pending_now(e, e.dep_ids[k])) err |= ERR_REVISITED;
            if (e.nl_epoch && e.nl_epoch[s] == req_epoch) err |= ERR_NONLISTEN;
        }
    }
    if (err) atomicOr(e.err_flag, err);
}
__global__ void k_check_independent(View e, Segs sg, uint32_t lvl_epoch, uint32_t req_epoch, int check_mode) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sg.total) return;
    check_member(e, seg_member(sg, e.front, i), lvl_epoch, req_epoch, check_mode);
}
// strict rule E, second half (after the set_value! effects of a level): a non-listening notification of this level left a
// signal that is neither done nor a member with complete criteria
__global__ void k_check_nl(View e, uint32_t req_epoch) {
    const uint32_t n = *e.nl_cnt;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t L = e.nl_list[i];
        if (e.done_epoch[L] != req_epoch && criteria(e, L)) atomicOr(e.err_flag, ERR_NONLISTEN);
    }
}

// ---- values ------------------------------------------------------------------------------------------------
template <class T>
__global__ void k_write_values(T* val, int dim, const uint32_t* list, const T* src, uint32_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * dim) return;
    val[(size_t)list[i / dim] * dim + (i % dim)] = src[i];
}
template <class T>
__global__ void k_gather_values(const T* val, int dim, const uint32_t* list, T* dst, uint32_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * dim) return;
    dst[i] = val[(size_t)list[i / dim] * dim + (i % dim)];
}
__global__ void k_pending_single(View e, uint32_t s, int* out) { *out = pending_eval(e, s) ? 1 : 0; }

// ---- rule kernels ----------------------------------------------------------------------------------------------
// Small fixed-size values (dim <= 4): one thread per signal. Families GAUSS_CANON / GAUSS_MV / BETA / SUM and
// the scalar m2v rules. rule < 0 => family reduce (m2f / ProductOfMessages / marginal), left-to-right
// (test/inference_engine_tests.jl:385-413).
// moments of a 2-parameter value by family (test/runtests.jl:36-38, 57-59, 68-69); POINT: the value itself, variance 0
template <class T>
__device__ __forceinline__ T vmp_mean(int fam, const T* v) {
    return fam == CXB_FAMILY_GAMMA ? v[0] * v[1] : v[0];
}
template <class T>
__device__ __forceinline__ T vmp_var(int fam, const T* v) {
    if (fam == CXB_FAMILY_GAMMA) return v[0] * (v[1] * v[1]);
    if (fam == CXB_FAMILY_GAUSS_MP) return T(1) / v[1];
    if (fam == CXB_FAMILY_GAUSS_MV) return v[1];
    return T(0);
}
// CXB_RULE_PROGRAM: the user's stack program (include/cortex_b200.h, CXB_OP_*; the oracle's run_program is the same machine).
// prog = { n_consts, consts..., code... } in the engine dtype, n_prog elements.
template <class T>
__device__ bool run_program(const T* __restrict__ prog, int n_prog, const View& e, const T* __restrict__ val, uint32_t off, uint32_t nd, T param,
                            T* __restrict__ out) {
    const int dim = e.dim;
    if (n_prog < 1) return false;
    const int nc = (int)prog[0];
    if (nc < 0 || nc + 1 > n_prog) return false;
    const T* consts = prog + 1;
    const T* code = consts + nc;
    const int n_code = n_prog - 1 - nc;
    T st[16], tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int sp = 0, pc = 0;
    while (pc < n_code) {
        const int op = (int)code[pc++];
        int a = 0, b = 0;
        if (op == CXB_OP_DEP || op == CXB_OP_CONST || op == CXB_OP_STORE || op == CXB_OP_TSET || op == CXB_OP_TGET) {
            if (pc >= n_code) return false;
            a = (int)code[pc++];
        }
        if (op == CXB_OP_DEP) {
            if (pc >= n_code) return false;
            b = (int)code[pc++];
        }
        switch (op) {
            case CXB_OP_DEP:
                if (a < 0 || a >= (int)nd || b < 0 || b >= dim || sp >= 16) return false;
                st[sp++] = val[(size_t)e.dep_ids[off + a] * dim + b];
                break;
            case CXB_OP_CONST:
                if (a < 0 || a >= nc || sp >= 16) return false;
                st[sp++] = consts[a];
                break;
            case CXB_OP_PARAM:
                if (sp >= 16) return false;
                st[sp++] = param;
                break;
            case CXB_OP_NDEPS:
                if (sp >= 16) return false;
                st[sp++] = (T)nd;
                break;
            case CXB_OP_ADD: case CXB_OP_SUB: case CXB_OP_MUL: case CXB_OP_DIV: {
                if (sp < 2) return false;
                const T y = st[--sp], x = st[--sp];
                st[sp++] = op == CXB_OP_ADD ? x + y : op == CXB_OP_SUB ? x - y : op == CXB_OP_MUL ? x * y : x / y;
                break;
            }
            case CXB_OP_NEG: case CXB_OP_EXP: case CXB_OP_LOG: case CXB_OP_SQRT:
                if (sp < 1) return false;
                st[sp - 1] = op == CXB_OP_NEG ? -st[sp - 1] : op == CXB_OP_EXP ? exp(st[sp - 1]) : op == CXB_OP_LOG ? log(st[sp - 1]) : sqrt(st[sp - 1]);
                break;
            case CXB_OP_STORE:
                if (a < 0 || a >= dim || a >= 8 || sp < 1) return false;
                out[a] = st[--sp];
                break;
            case CXB_OP_TSET:
                if (a < 0 || a >= 8 || sp < 1) return false;
                tmp[a] = st[--sp];
                break;
            case CXB_OP_TGET:
                if (a < 0 || a >= 8 || sp >= 16) return false;
                st[sp++] = tmp[a];
                break;
            default:
                return false;
        }
    }
    return true;
}
template <class T>
__device__ __forceinline__ void rule_small_one(const View& e, T* __restrict__ val, uint32_t s, int family, int rule,
                                               const T* __restrict__ fparam, T default_param, const T* __restrict__ prog = nullptr, int n_prog = 0) {
    const int dim = e.dim;
    uint32_t off = e.dep_off[s], nd = e.dep_off[s + 1] - off;
    T acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (nd == 0) {
        atomicOr(e.err_flag, ERR_RULE_ARG);
        return;
    }
    {
        const T* a = val + (size_t)e.dep_ids[off] * dim;
        for (int k = 0; k < dim; ++k) acc[k] = a[k];
    }
    if (e.sfam) family = e.sfam[s];
    if (rule < 0) {
        if (family == CXB_FAMILY_POINT) {  // an observed value has no product
            atomicOr(e.err_flag, ERR_RULE_ARG);
            return;
        }
        for (uint32_t j = 1; j < nd; ++j) {
            const T* b = val + (size_t)e.dep_ids[off + j] * dim;
            if (family == CXB_FAMILY_GAUSS_MP) {  // test/runtests.jl:89-95
                T xi = acc[0] * acc[1] + b[0] * b[1];
                T w = acc[1] + b[1];
                acc[0] = (T(1) / w) * xi;
                acc[1] = w;
            } else if (family == CXB_FAMILY_GAMMA) {  // test/runtests.jl:97-99
                acc[0] = acc[0] + b[0] - T(1);
                acc[1] = (acc[1] * b[1]) / (acc[1] + b[1]);
            } else if (family == CXB_FAMILY_GAUSS_CANON || family == CXB_FAMILY_SUM) {
                for (int k = 0; k < dim; ++k) acc[k] = acc[k] + b[k];
            } else if (family == CXB_FAMILY_GAUSS_MV) {  // test/runtests.jl:40-46
                T xi = acc[0] / acc[1] + b[0] / b[1];
                T w = T(1) / acc[1] + T(1) / b[1];
                T variance = T(1) / w;
                acc[0] = variance * xi;
                acc[1] = variance;
            } else if (family == CXB_FAMILY_BETA) {  // test/inference_engine_tests.jl:273-278
                acc[0] = acc[0] + b[0] - T(1);
                acc[1] = acc[1] + b[1] - T(1);
            }
        }
    } else {
        T p = default_param;
        if (fparam) {
            T fp = fparam[e.sfac[s]];
            if (fp == fp) p = fp;  // NaN = "not set": use the rule default
        }
        switch (rule) {
            case CXB_RULE_GAUSS_OBS: {
                T y = acc[0];
                acc[0] = T(1) / p;
                acc[1] = y / p;
                break;
            }
            case CXB_RULE_GAUSS_RW: {
                T den = T(1) + p * acc[0];
                acc[0] = acc[0] / den;
                acc[1] = acc[1] / den;
                break;
            }
            case CXB_RULE_GAUSS_MV_OBS:
                acc[1] = p;
                break;
            case CXB_RULE_GAUSS_MV_RW:
                acc[1] = acc[1] + p;
                break;
            case CXB_RULE_BETA_BERNOULLI: {
                T r = acc[0];
                acc[0] = T(1) + r;
                acc[1] = T(2) - r;
                break;
            }
            case CXB_RULE_SCALE2:
                for (int k = 0; k < dim; ++k) acc[k] = T(2) * acc[k];
                break;
            case CXB_RULE_PROGRAM: {
                for (int k = 0; k < 8; ++k) acc[k] = T(0);
                if (!prog || !run_program<T>(prog, n_prog, e, val, off, nd, p, acc)) {
                    atomicOr(e.err_flag, ERR_RULE_ARG);
                    return;
                }
                break;
            }
            case CXB_RULE_NORMAL_MEAN_FIELD: {  // test/inference_engine_tests.jl:652-695
                if (nd != 2 || dim < 2 || !e.sfam) {
                    atomicOr(e.err_flag, ERR_RULE_ARG);
                    return;
                }
                const uint32_t da = e.dep_ids[off], db = e.dep_ids[off + 1];
                const int fa = e.sfam[da], fb = e.sfam[db];
                const T* a = val + (size_t)da * dim;
                const T* b = val + (size_t)db * dim;
                if ((fa == CXB_FAMILY_GAMMA) != (fb == CXB_FAMILY_GAMMA)) {  // towards out / mean (:666-676)
                    const T* w = fa == CXB_FAMILY_GAMMA ? a : b;
                    const T* o = fa == CXB_FAMILY_GAMMA ? b : a;
                    acc[0] = vmp_mean<T>(fa == CXB_FAMILY_GAMMA ? fb : fa, o);
                    acc[1] = w[0] * w[1];
                } else if (fa != CXB_FAMILY_GAMMA) {  // towards the precision (:678-692)
                    const T dm = vmp_mean<T>(fa, a) - vmp_mean<T>(fb, b);
                    acc[0] = T(1.5);
                    acc[1] = T(2) / (vmp_var<T>(fa, a) + vmp_var<T>(fb, b) + dm * dm);
                } else {
                    atomicOr(e.err_flag, ERR_RULE_ARG);
                    return;
                }
                break;
            }
            case CXB_RULE_NORMAL_STRUCTURED: {  // test/inference_engine_tests.jl:942-973, 1008-1028
                const uint32_t d0 = e.dep_ids[off];
                const T* v0 = val + (size_t)d0 * dim;
                if (e.kind[s] == CXB_KIND_JOINT) {  // (m2f a, m2f b, marginal of the precision) -> MvNormalMeanPrecision
                    if (nd != 3 || dim < 6) {
                        atomicOr(e.err_flag, ERR_RULE_ARG);
                        return;
                    }
                    const T* v1 = val + (size_t)e.dep_ids[off + 1] * dim;
                    const T* g = val + (size_t)e.dep_ids[off + 2] * dim;
                    const T xi_a = v0[1] * v0[0], w_a = v0[1], xi_b = v1[1] * v1[0], w_b = v1[1], w = g[0] * g[1];
                    const T W11 = w_a + w, W12 = -w, W21 = -w, W22 = w_b + w;
                    const T det = W11 * W22 - W12 * W21;
                    acc[0] = (W22 * xi_a - W12 * xi_b) / det;
                    acc[1] = (W11 * xi_b - W21 * xi_a) / det;
                    acc[2] = W11;
                    acc[3] = W12;
                    acc[4] = W21;
                    acc[5] = W22;
                    for (int k = 6; k < dim; ++k) acc[k] = T(0);
                } else if (nd == 2) {  // (m2f of the other state, marginal of the precision), in either order
                    const uint32_t d1 = e.dep_ids[off + 1];
                    const T* v1 = val + (size_t)d1 * dim;
                    const bool first_is_msg = e.kind[d0] == CXB_KIND_M2F;
                    if (first_is_msg == (e.kind[d1] == CXB_KIND_M2F)) {
                        atomicOr(e.err_flag, ERR_RULE_ARG);
                        return;
                    }
                    const T* m = first_is_msg ? v0 : v1;
                    const T* g = first_is_msg ? v1 : v0;
                    acc[0] = m[0];
                    acc[1] = T(1) / (T(1) / m[1] + T(1) / (g[0] * g[1]));
                    for (int k = 2; k < dim; ++k) acc[k] = T(0);
                } else if (nd == 1 && e.kind[d0] == CXB_KIND_JOINT && dim >= 6) {  // (JointMarginal) -> Gamma
                    const T det = v0[2] * v0[5] - v0[3] * v0[4];
                    const T V11 = v0[5] / det, V12 = -v0[3] / det, V21 = -v0[4] / det, V22 = v0[2] / det;
                    const T dm = v0[0] - v0[1];
                    acc[0] = T(1.5);
                    acc[1] = T(2) / (V11 - V12 - V21 + V22 + dm * dm);
                    for (int k = 2; k < dim; ++k) acc[k] = T(0);
                } else {
                    atomicOr(e.err_flag, ERR_RULE_ARG);
                    return;
                }
                break;
            }
            default:
                atomicOr(e.err_flag, ERR_RULE_ARG);
                return;
        }
    }
    T* o = val + (size_t)s * dim;
    for (int k = 0; k < dim; ++k) o[k] = acc[k];
}
template <class T>
__global__ void k_rule_small(View e, T* __restrict__ val, const uint32_t* list, uint32_t n, int family, int rule,
                             const T* __restrict__ fparam, T default_param, const T* __restrict__ prog, int n_prog) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || *(volatile int*)e.err_flag) return;  // refused level: no value is written
    rule_small_one<T>(e, val, list[i], family, rule, fparam, default_param, prog, n_prog);
}

// ---- resident level loop --------------------------------------------------------------------------------------------------
// update_marginals! (src/inference_engine.jl:559-632) of a SMALL graph entirely on the device: one CTA runs request
// levels, breadth-first discovery, independence check, rules, set_value! side effects, readiness and the final phase in
// a loop with block barriers only — no kernel launch and no host round trip per level (a T = 1000 chain has 2T-1
// levels of 2-3 signals each). Same device functions, same state machine and the same frontier sets as the
// per-level kernels above; used for the small-value families when no trace is requested.
constexpr int ERR_NO_RULE_KEY = 8;
struct ResidentArgs {
    const uint32_t* req_marg;
    const uint32_t* link_ids;
    uint8_t* ready;
    uint32_t n_req, n_links;
    uint32_t *list_a, *list_b;
    uint32_t req_epoch, lvl_epoch0;
    int n_keys, family;
    const int* key_rule;       // [n_keys] rule kind of a key: -1 = family reduce, -2 = no rule registered
    const double* key_param;   // [n_keys] default parameter of the key's rule
    const void* fparam;        // per factor id, NaN = unset
    const void* tables;        // categorical family: every CAT_TABLE / HMM_EMIT table (engine dtype)
    const long long* key_table;  // [n_keys] element offset of the key's table in `tables` (CAT_TABLE: [2][K][K], HMM_EMIT: [K][n_sym])
    const int* key_nsym;       // [n_keys] HMM_EMIT: number of symbols
    long long* out;            // [0] levels [1] updates [2] final marginals [3] final linked [4] last lvl_epoch [5] key without rule
    uint32_t *rec_list, *rec_desc;  // schedule recording (memoised replay): members level by level, descriptors (level, key, count, offset)
    uint32_t rec_list_cap, rec_desc_cap;  // out[6] = descriptors written, out[7] = members written (0xFFFFFFFF.. = overflow)
    const uint8_t* pend_at_req;  // [n_req] strict rule B (k_request_check)
    const uint32_t *hz_m, *hz_l;  // rule F: signals that must not be pending when the marginal / linked level of the final phase starts
    uint32_t n_hz_m, n_hz_l;
};
// rule F (final phase): see DeviceEngine::request for how the lists are made
__global__ void k_check_not_pending(View e, const uint32_t* list, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && pending_now(e, list[i])) atomicOr(e.err_flag, ERR_FINAL_ORDER);
}
__device__ __forceinline__ void apply_one(const View& e, uint32_t s, uint32_t req_epoch, int check_mode) {
    for (uint32_t c = e.nib_off[s]; c < e.nib_off[s + 1]; ++c) e.nib[c] &= ~ALL_F;  // unset_all_dependencies_fresh!
    e.props[s] = P_COMPUTED;
    if (check_mode) e.done_epoch[s] = req_epoch;
    atomicAdd(&e.kind_count[e.kind[s]], 1ull);
    for (uint32_t k = e.lis_off[s]; k < e.lis_off[s + 1]; ++k) notify(e, k, req_epoch, check_mode);
}
// categorical rule of ONE signal by ONE warp (lane l owns components l, l + 32; K <= 64): the body of k_rule_cat with the
// tables read from global memory (small graphs: they sit in L1 / L2) and the incoming message staged per warp
template <class T>
__device__ __forceinline__ void rule_cat_warp(const View& e, T* __restrict__ val, uint32_t s, int rule, const T* __restrict__ table, int n_sym,
                                              T potts_w, T* __restrict__ sh_in) {
    const int K = e.dim, lane = threadIdx.x & 31;
    const uint32_t off = e.dep_off[s], nd = e.dep_off[s + 1] - off;
    if (nd == 0) {
        if (lane == 0) atomicOr(e.err_flag, ERR_RULE_ARG);
        return;
    }
    T acc[2] = {T(0), T(0)};
    {
        const T* a0 = val + (size_t)e.dep_ids[off] * K;
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (lane + 32 * c < K) acc[c] = a0[lane + 32 * c];
    }
    if (rule < 0) {
        for (uint32_t j = 1; j < nd; ++j) {
            const T* b = val + (size_t)e.dep_ids[off + j] * K;
#pragma unroll
            for (int c = 0; c < 2; ++c)
                if (lane + 32 * c < K) acc[c] = acc[c] * b[lane + 32 * c];
        }
    } else if (rule == CXB_RULE_POTTS) {
        T part = acc[0] + acc[1];
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[c] = (lane + 32 * c < K) ? part + potts_w * acc[c] : T(0);
    } else if (rule == CXB_RULE_CAT_TABLE) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (lane + 32 * c < K) sh_in[lane + 32 * c] = acc[c];
        __syncwarp();
        const bool u_is_low = e.svar[e.dep_ids[off]] < e.svar[s];
        const T* tb = table + (u_is_low ? 0 : (size_t)K * K);  // [x_lo][x_hi] / its transpose
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int aidx = lane + 32 * c;
            T sum = T(0);
            if (aidx < K)
                for (int b = 0; b < K; ++b) sum += tb[b * K + aidx] * sh_in[b];
            acc[c] = sum;
        }
        __syncwarp();
    } else if (rule == CXB_RULE_HMM_EMIT) {
        int o = (int)val[(size_t)e.dep_ids[off] * K];
        if (o < 0 || o >= n_sym) {
            if (lane == 0) atomicOr(e.err_flag, ERR_RULE_ARG);
            o = 0;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[c] = (lane + 32 * c < K) ? table[(size_t)(lane + 32 * c) * n_sym + o] : T(0);
    } else {
        if (lane == 0) atomicOr(e.err_flag, ERR_RULE_ARG);
        return;
    }
    T part = acc[0] + acc[1];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    T* o = val + (size_t)s * K;
#pragma unroll
    for (int c = 0; c < 2; ++c)
        if (lane + 32 * c < K) o[lane + 32 * c] = acc[c] / part;
}

template <class T>
__global__ void __launch_bounds__(1024) k_update_resident(View e, T* __restrict__ val, ResidentArgs a) {
    __shared__ uint32_t s_cnt[2];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_key_cnt[256], s_key_base[256];  // per-key frontier cursors / partition starts (<= 252 keys)
    __shared__ int s_key_rule[256];
    __shared__ uint32_t s_rec_off[256], s_rec_n[3];  // recording: per-key offsets of the level, [0] members [1] descriptors [2] level index
    __shared__ T s_cat_in[32][64];  // categorical family: one incoming message per warp (K <= 64)
    const uint32_t tid = threadIdx.x, NT = blockDim.x;
    if (tid < 3) s_rec_n[tid] = 0;
    for (int k = tid; k < a.n_keys; k += NT) {
        s_key_base[k] = e.key_base[k];
        s_key_rule[k] = a.key_rule[k];
    }
    e.key_cnt = s_key_cnt;    // frontier_push bumps the cursors in shared memory
    e.key_base = s_key_base;
    __syncthreads();
    uint32_t lvl = a.lvl_epoch0;
    long long levels = 0, updates = 0, fin[2] = {0, 0};
    if (*(volatile int*)e.err_flag) {  // refused at request time (k_request_check): nothing is computed
        if (tid == 0) a.out[4] = lvl;
        return;
    }

    // the members of the current level (per-key segments of the frontier buffer): independence, rules, side effects
    auto run_level = [&](int check_mode) {
        for (int k = 0; k < a.n_keys; ++k) {
            const uint32_t cnt = e.key_cnt[k], base = e.key_base[k];
            for (uint32_t i = tid; i < cnt; i += NT) check_member(e, e.front[base + i], lvl, a.req_epoch, check_mode);
        }
        __syncthreads();
        if (*(volatile int*)e.err_flag) return;  // refused before anything of this level is computed or applied (uniform)
        if (a.rec_list) {  // record the level for the memoised schedule
            if (tid == 0) {
                uint32_t off = s_rec_n[0], nd = s_rec_n[1];
                for (int k = 0; k < a.n_keys; ++k) {
                    const uint32_t cnt = e.key_cnt[k];
                    s_rec_off[k] = off;
                    if (!cnt) continue;
                    if (nd < a.rec_desc_cap && off + cnt <= a.rec_list_cap) {
                        a.rec_desc[4 * nd] = s_rec_n[2];
                        a.rec_desc[4 * nd + 1] = (uint32_t)k;
                        a.rec_desc[4 * nd + 2] = cnt;
                        a.rec_desc[4 * nd + 3] = off;
                        ++nd;
                    } else {
                        nd = 0xFFFFFFFFu;  // overflow: the recording is dropped
                    }
                    off += cnt;
                }
                s_rec_n[0] = off;
                s_rec_n[1] = nd;
                ++s_rec_n[2];
            }
            __syncthreads();
            if (s_rec_n[1] != 0xFFFFFFFFu)
                for (int k = 0; k < a.n_keys; ++k) {
                    const uint32_t cnt = e.key_cnt[k], base = e.key_base[k];
                    for (uint32_t i = tid; i < cnt; i += NT) a.rec_list[s_rec_off[k] + i] = e.front[base + i];
                }
        }
        for (int k = 0; k < a.n_keys; ++k) {
            const uint32_t cnt = e.key_cnt[k], base = e.key_base[k];
            if (!cnt) continue;
            const int rule = s_key_rule[k];
            if (rule == -2) {
                if (tid == 0) {
                    atomicOr(e.err_flag, ERR_NO_RULE_KEY);
                    a.out[5] = k;
                }
                continue;
            }
            const T defp = (T)a.key_param[k];
            if (a.family == CXB_FAMILY_CATEGORICAL) {  // one warp per signal; key_param = Potts weight e^beta - 1
                const T* tb = a.key_table[k] >= 0 ? (const T*)a.tables + a.key_table[k] : nullptr;
                for (uint32_t i = tid >> 5; i < cnt; i += NT >> 5)
                    rule_cat_warp<T>(e, val, e.front[base + i], rule, tb, a.key_nsym[k], defp, s_cat_in[tid >> 5]);
            } else {
                for (uint32_t i = tid; i < cnt; i += NT)
                    rule_small_one<T>(e, val, e.front[base + i], a.family, rule, (const T*)a.fparam, defp,
                                      a.key_table[k] >= 0 ? (const T*)a.tables + a.key_table[k] : nullptr, a.key_nsym[k]);
            }
        }
        __syncthreads();
        if (*(volatile int*)e.err_flag) return;  // a rule rejected its arguments
        for (int k = 0; k < a.n_keys; ++k) {
            const uint32_t cnt = e.key_cnt[k], base = e.key_base[k];
            for (uint32_t i = tid; i < cnt; i += NT) apply_one(e, e.front[base + i], a.req_epoch, check_mode);
        }
        __syncthreads();
        if (check_mode == 1 && e.nl_epoch) {  // strict rule E, second half
            const uint32_t n_nl = *e.nl_cnt;
            for (uint32_t i = tid; i < n_nl; i += NT) {
                const uint32_t L = e.nl_list[i];
                if (e.done_epoch[L] != a.req_epoch && criteria(e, L)) atomicOr(e.err_flag, ERR_NONLISTEN);
            }
            __syncthreads();
            if (tid == 0) *e.nl_cnt = 0;
            __syncthreads();
        }
    };
    auto begin_level = [&]() {
        ++lvl;
        if (tid == 0) s_cnt[0] = s_cnt[1] = 0;
        for (int k = tid; k < a.n_keys; k += NT) e.key_cnt[k] = 0;
        __syncthreads();
    };
    auto frontier_total = [&]() -> uint32_t {
        if (tid == 0) {
            uint32_t t = 0;
            for (int k = 0; k < a.n_keys; ++k) t += e.key_cnt[k];
            s_total = t;
        }
        __syncthreads();
        const uint32_t t = s_total;
        __syncthreads();
        return t;
    };

    bool first_level = true;
    auto bfs_all = [&](int flags) {  // process_dependencies! from the seeds in list_a, one breadth-first step per iteration
        int which = 0;
        for (;;) {
            const uint32_t n_in = s_cnt[which];
            __syncthreads();
            if (n_in == 0) break;
            if (tid == 0) s_cnt[which ^ 1] = 0;
            __syncthreads();
            const uint32_t* in = which ? a.list_b : a.list_a;
            uint32_t* out = which ? a.list_a : a.list_b;
            for (uint32_t i = tid; i < n_in; i += NT)
                bfs_visit(e, in[i], lvl, a.req_epoch, flags, [&](uint32_t x) { out[atomicAdd(&s_cnt[which ^ 1], 1u)] = x; });
            __syncthreads();
            which ^= 1;
        }
    };
    while (a.n_req) {
        begin_level();
        if (first_level && e.strict) {
            // strict rule B: the marginals that were pending when the request arrived are traversed first, alone - the
            // reference gives such a variable one traversal and takes its marginal, so nothing may be pending beneath it
            for (uint32_t i = tid; i < a.n_req; i += NT)
                if (a.pend_at_req[i]) a.list_a[atomicAdd(&s_cnt[0], 1u)] = a.req_marg[i];
            __syncthreads();
            bfs_all(3);
            if (frontier_total()) {
                if (tid == 0) atomicOr(e.err_flag, ERR_WORK_BENEATH_PENDING);
                break;
            }
            if (tid == 0) s_cnt[0] = s_cnt[1] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < a.n_req; i += NT)
                if (!a.pend_at_req[i]) a.list_a[atomicAdd(&s_cnt[0], 1u)] = a.req_marg[i];
        } else {
            for (uint32_t i = tid; i < a.n_req; i += NT)
                if (!a.ready[i]) a.list_a[atomicAdd(&s_cnt[0], 1u)] = a.req_marg[i];  // seeds (:585)
        }
        __syncthreads();
        bfs_all(first_level ? 3 : 1);
        first_level = false;
        const uint32_t total = frontier_total();
        if (*(volatile int*)e.err_flag) break;  // refused by a traversal rule (A, weak beneath done): nothing of this level runs
        if (!total) break;
        run_level(1);
        if (*(volatile int*)e.err_flag) break;  // uniform: every thread reads it after the barrier that ends run_level
        for (uint32_t i = tid; i < a.n_req; i += NT)
            if (!a.ready[i] && pending_eval(e, a.req_marg[i])) a.ready[i] = 1;  // readiness (:593-595)
        ++levels;
        updates += total;
        __syncthreads();
    }
    for (int mode = 0; mode < 2 && a.n_req; ++mode) {  // final phase: marginals, then linked signals (:610-628)
        __syncthreads();
        if (*(volatile int*)e.err_flag) break;
        {  // rule F: candidates that depend on each other across the reference's per-variable order
            const uint32_t* hz = mode == 0 ? a.hz_m : a.hz_l;
            const uint32_t n_hz = mode == 0 ? a.n_hz_m : a.n_hz_l;
            for (uint32_t i = tid; i < n_hz; i += NT)
                if (pending_now(e, hz[i])) atomicOr(e.err_flag, ERR_FINAL_ORDER);
            if (n_hz) {
                __syncthreads();
                if (*(volatile int*)e.err_flag) break;
            }
        }
        begin_level();
        const uint32_t cnt = mode == 0 ? a.n_req : a.n_links;
        for (uint32_t i = tid; i < cnt; i += NT) {
            const uint32_t s = mode == 0 ? a.req_marg[i] : a.link_ids[i];
            if (pending_eval(e, s)) frontier_push(e, s, lvl);
        }
        __syncthreads();
        const uint32_t total = frontier_total();
        fin[mode] = total;
        if (total) run_level(2);
        updates += total;
    }
    if (tid == 0) {
        a.out[0] = levels;
        a.out[1] = updates;
        a.out[2] = fin[0];
        a.out[3] = fin[1];
        a.out[4] = lvl;
        a.out[6] = s_rec_n[1] == 0xFFFFFFFFu ? -1 : (long long)s_rec_n[1];
        a.out[7] = s_rec_n[0];
    }
}

// ---- memoised schedule: replay ----------------------------------------------------------------------------------------
// The level schedule is a function of the request, the structure and the FLAG state (props, nibbles) alone - values never
// steer it. A request that arrives with the same ids and a flag state bit-identical to a recorded run therefore executes
// the same signals level by level and ends in the same flag state: the engine replays the recorded levels (rules only: no
// traversal, no checks, no set_value! bookkeeping) and copies the recorded final flags. Small graphs: all levels in one
// launch (block barrier between levels); large graphs: one batched rule kernel per (level, rule key), no host round trip.
__global__ void k_state_differs(const uint8_t* props, const uint8_t* props0, size_t n, const uint64_t* nib, const uint64_t* nib0, size_t n_chunks,
                                int* flag) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool diff = false;
    for (size_t i = t; i < n && !diff; i += stride) diff = props[i] != props0[i];
    for (size_t i = t; i < n_chunks && !diff; i += stride) diff = nib[i] != nib0[i];
    if (diff) *flag = 1;
}
// Observable equality of two engine states over the same structure (what the tests compare): computed flag, is_pending
// (evaluated without caching, each state with its own props and nibbles: the lazily cached (pp, p) pair may differ between two
// schedules that left the same pending state), nibble chunks bit for bit, values bit for bit.
__global__ void k_state_equiv(View a, View b, const uint32_t* val_a, const uint32_t* val_b, size_t val_words, int* flag) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool diff = false;
    for (size_t s = t; s < a.n_sig && !diff; s += stride) {
        const uint32_t i = (uint32_t)s;
        diff = ((a.props[i] ^ b.props[i]) & P_COMPUTED) || pending_now(a, i) != pending_now(b, i);
        for (uint32_t c = a.nib_off[i]; c < a.nib_off[i + 1] && !diff; ++c) diff = a.nib[c] != b.nib[c];
    }
    for (size_t w = t; w < val_words && !diff; w += stride) diff = val_a[w] != val_b[w];
    if (diff) *flag = 1;
}
struct ReplayArgs {
    const uint32_t *list, *desc;
    uint32_t n_desc;
    int n_keys, family;
    const int* key_rule;
    const double* key_param;
    const void* fparam;
    const void* tables;
    const long long* key_table;
    const int* key_nsym;
};
template <class T>
__global__ void __launch_bounds__(1024) k_replay_resident(View e, T* __restrict__ val, ReplayArgs a) {
    __shared__ int s_key_rule[256];
    __shared__ T s_cat_in[32][64];
    const uint32_t tid = threadIdx.x, NT = blockDim.x;
    for (int k = tid; k < a.n_keys; k += NT) s_key_rule[k] = a.key_rule[k];
    __syncthreads();
    uint32_t cur_level = 0xFFFFFFFFu;
    for (uint32_t d = 0; d < a.n_desc; ++d) {
        const uint32_t level = a.desc[4 * d], key = a.desc[4 * d + 1], cnt = a.desc[4 * d + 2], off = a.desc[4 * d + 3];
        if (level != cur_level) {
            __syncthreads();  // the values of a level are inputs of the next
            cur_level = level;
        }
        const int rule = s_key_rule[key];
        if (rule == -2) {
            if (tid == 0) atomicOr(e.err_flag, ERR_RULE_ARG);
            continue;
        }
        const T defp = (T)a.key_param[key];
        if (a.family == CXB_FAMILY_CATEGORICAL) {
            const T* tb = a.key_table[key] >= 0 ? (const T*)a.tables + a.key_table[key] : nullptr;
            for (uint32_t i = tid >> 5; i < cnt; i += NT >> 5) rule_cat_warp<T>(e, val, a.list[off + i], rule, tb, a.key_nsym[key], defp, s_cat_in[tid >> 5]);
        } else {
            for (uint32_t i = tid; i < cnt; i += NT)
                rule_small_one<T>(e, val, a.list[off + i], a.family, rule, (const T*)a.fparam, defp,
                                  a.key_table[key] >= 0 ? (const T*)a.tables + a.key_table[key] : nullptr, a.key_nsym[key]);
        }
    }
}

// ---- sequential executor -----------------------------------------------------------------------------------------------
// update_marginals! (src/inference_engine.jl:559-632), request scanning (:540-546) and process_dependencies!
// (src/signal.jl:466-490) LITERALLY, statement by statement, on the device: one warp walks the reference's depth-first
// traversal with an explicit stack, evaluates is_pending with its caching side effect, runs the rule of each pending
// signal the moment the reference would and applies set_value! before going on. Exact for any wiring (weak,
// non-listening, duplicate, hand-made dependencies; Gauss-Seidel orders) because it IS the reference order; one signal at
// a time, so it is the schedule of hand-wired graphs and the fallback of requests the level schedule refuses, not the
// throughput path. All 32 lanes execute the same control flow on the same data (loads are broadcasts); lane 0 alone
// stores, and a __syncwarp() after every store orders it before the warp's next loads; categorical rules use the lanes.
struct SeqArgs {
    const uint32_t *req_marg, *link_off, *link_ids;  // link_off[n_req + 1]: linked signals of requested variable i
    uint8_t* ready;
    uint32_t n_req;
    uint32_t* stack;  // frames of 3 words
    uint32_t stack_cap;
    int mode;  // 0 update_marginals!, 1 scan_inference_request (records pending signals, computes nothing), 2 process_dependencies!(table)
    int n_keys, family;
    const int* key_rule;
    const double* key_param;
    const void* fparam;
    const void* tables;
    const long long* key_table;
    const int* key_nsym;
    uint32_t* rec;  // mode 0: trace triples (round, request position, signal) when rec_cap > 0; modes 1, 2: visited / pending signals
    uint32_t rec_cap;
    const uint8_t* answers;  // mode 2: f(dep) = answers[dep]; nullptr: f = is_pending
    uint32_t root;
    int retry;
    long long max_updates;  // the reference's `while should_continue` loop need not end on cyclic wirings: stop and report
    long long* out;  // [0] rounds that executed something [1] updates [2] final marginals [3] final linked [5] key without rule [6] records [7] return value
};
template <class T>
__global__ void __launch_bounds__(32) k_seq(View e, T* __restrict__ val, SeqArgs a) {
    __shared__ T s_cat_in[64];
    const int lane = threadIdx.x;
    long long n_rec = 0, updates = 0;
    bool failed = false;

    auto is_pending = [&](uint32_t s) -> bool {  // src/signal.jl:141-154
        const uint8_t p = e.props[s];
        bool r = false;
        if (p & P_P) {
            r = true;
        } else if (p & P_PP) {
            r = criteria(e, s);
            __syncwarp();
            if (lane == 0) e.props[s] = (uint8_t)((p & P_COMPUTED) | (r ? P_P : 0));
        }
        __syncwarp();
        return r;
    };
    auto record3 = [&](uint32_t x, uint32_t y, uint32_t z) {
        if (a.rec_cap && (unsigned long long)(n_rec + 1) * 3 <= a.rec_cap && lane == 0) {
            a.rec[n_rec * 3] = x;
            a.rec[n_rec * 3 + 1] = y;
            a.rec[n_rec * 3 + 2] = z;
        }
        ++n_rec;
    };
    auto record1 = [&](uint32_t x) {
        if ((unsigned long long)n_rec < a.rec_cap && lane == 0) a.rec[n_rec] = x;
        ++n_rec;
    };
    // compute!(rule, signal) = rule + set_value! (src/signal.jl:392-410, 232-253): process! dispatch by rule key
    auto execute = [&](uint32_t s, uint32_t round, uint32_t pos) {
        if (updates >= a.max_updates) {
            if (lane == 0) atomicOr(e.err_flag, ERR_SEQ_DIVERGED);
            failed = true;
            __syncwarp();
            return;
        }
        const int key = e.rkey[s];
        const int rule = a.key_rule[key];
        if (rule == -2) {
            if (lane == 0) {
                atomicOr(e.err_flag, ERR_NO_RULE_KEY);
                a.out[5] = key;
            }
            failed = true;
            __syncwarp();
            return;
        }
        const T defp = (T)a.key_param[key];
        if (a.family == CXB_FAMILY_CATEGORICAL) {
            const T* tb = a.key_table[key] >= 0 ? (const T*)a.tables + a.key_table[key] : nullptr;
            rule_cat_warp<T>(e, val, s, rule, tb, a.key_nsym[key], defp, s_cat_in);
        } else if (lane == 0) {
            rule_small_one<T>(e, val, s, a.family, rule, (const T*)a.fparam, defp, a.key_table[key] >= 0 ? (const T*)a.tables + a.key_table[key] : nullptr,
                              a.key_nsym[key]);
        }
        __syncwarp();
        if (*(volatile int*)e.err_flag) {
            failed = true;
            return;
        }
        if (lane == 0) apply_one(e, s, 0u, 0);
        __syncwarp();
        record3(round, pos, s);
        ++updates;
    };
    // process_dependencies!(f, root; retry), src/signal.jl:466-490, with the recursion unrolled onto a.stack
    auto process_dependencies = [&](uint32_t root, bool retry, auto&& f) -> bool {
        uint32_t sp = 0, s = root, i = 0;
        bool any = false;
        for (;;) {
            if (failed) return any;
            const uint32_t off = e.dep_off[s], nd = e.dep_off[s + 1] - off;
            if (i == nd) {
                if (sp == 0) return any;
                const bool ip = any;  // what the recursive call returned
                --sp;
                s = a.stack[3 * sp];
                i = a.stack[3 * sp + 1];
                any = a.stack[3 * sp + 2] != 0;
                bool pr = false;
                if (ip && retry) pr = f(e.dep_ids[e.dep_off[s] + i]);
                any = any || ip || pr;
                ++i;
                continue;
            }
            const uint32_t d = e.dep_ids[off + i];
            const bool pr = f(d);
            if (!pr) {
                const uint32_t nibble = (uint32_t)(e.nib[e.nib_off[s] + (i >> 4)] >> ((i & 15) << 2)) & 0xF;
                if (nibble & CXB_NIB_INTERMEDIATE) {
                    if (sp == a.stack_cap) {
                        if (lane == 0) atomicOr(e.err_flag, ERR_SEQ_OVERFLOW);
                        failed = true;
                        __syncwarp();
                        return any;
                    }
                    if (lane == 0) {
                        a.stack[3 * sp] = s;
                        a.stack[3 * sp + 1] = i;
                        a.stack[3 * sp + 2] = any ? 1u : 0u;
                    }
                    __syncwarp();
                    ++sp;
                    s = d;
                    i = 0;
                    any = false;
                    continue;
                }
            }
            any = any || pr;
            ++i;
        }
    };

    if (a.mode == 2) {
        const bool r = process_dependencies(a.root, a.retry != 0, [&](uint32_t d) -> bool {
            record1(d);
            return a.answers ? a.answers[d] != 0 : is_pending(d);
        });
        if (lane == 0) {
            a.out[6] = n_rec;
            a.out[7] = r ? 1 : 0;
        }
        return;
    }
    if (a.mode == 1) {  // scan_inference_request, src/inference_engine.jl:540-546
        for (uint32_t i = 0; i < a.n_req; ++i)
            process_dependencies(a.req_marg[i], true, [&](uint32_t d) -> bool {
                if (is_pending(d)) {
                    record1(d);
                    return true;
                }
                return false;
            });
        if (lane == 0) a.out[6] = n_rec;
        return;
    }
    long long rounds_exec = 0, fm = 0, fl = 0;
    bool should_continue = a.n_req > 0, reverse = false;
    uint32_t round = 0;
    while (should_continue && !failed) {  // :575-608
        bool cont = false;
        const long long updates0 = updates;
        for (uint32_t k = 0; k < a.n_req && !failed; ++k) {
            const uint32_t i = reverse ? a.n_req - 1 - k : k;
            if (a.ready[i]) continue;
            const bool processed = process_dependencies(a.req_marg[i], true, [&](uint32_t d) -> bool {  // :512-525
                if (failed) return false;
                if (is_pending(d)) {
                    execute(d, round, i);
                    return !failed;
                }
                return false;
            });
            if (failed) break;
            if (is_pending(a.req_marg[i])) {  // :593-595
                if (lane == 0) a.ready[i] = 1;
                __syncwarp();
            }
            cont = cont || processed;
        }
        if (updates > updates0) ++rounds_exec;
        reverse = !reverse;
        should_continue = cont;
        ++round;
    }
    for (uint32_t i = 0; i < a.n_req && !failed; ++i) {  // final phase, :610-628: marginal, then the linked signals, variable by variable
        const uint32_t m = a.req_marg[i];
        if (is_pending(m)) {
            execute(m, 0xFFFFFFFFu, i);
            ++fm;
        }
        for (uint32_t k = a.link_off[i]; k < a.link_off[i + 1] && !failed; ++k) {
            const uint32_t l = a.link_ids[k];
            if (!is_pending(l)) continue;
            execute(l, 0xFFFFFFFFu, i);
            ++fl;
        }
    }
    if (lane == 0) {
        a.out[0] = rounds_exec;
        a.out[1] = updates;
        a.out[2] = fm;
        a.out[3] = fl;
        a.out[6] = n_rec;
    }
}

// Categorical values (dim = K states): G lanes cooperate on one signal (G = min(32, pow2 >= K)), lane l owns
// components l, l+G, ... . rule < 0: element-wise product of the dependencies, normalised (App. C);
// CAT_TABLE: out[a] = sum_b psi(a,b) in[b] with the table staged in shared memory; POTTS: closed form;
// HMM_EMIT: column of the emission table.  Normalisation by warp-shuffle reduction.
// Dependency lists of the members of a recorded level, flattened once when the memo is committed: member i of the memo's
// lists reads the ids src[off[i] .. off[i + 1]) (and `low[i]`: its first dependency's variable has the lower id - the table
// orientation of CAT_TABLE). A replay then streams its indices instead of chasing list -> dep_off -> dep_ids -> svar through
// the CSR: ncu showed 456 B of DRAM reads per m2v signal on the power-law graph, 8 x what the rule needs (every 4-byte
// index fetched a sector of its own). Null `off`: the engine's CSR (first run of a request). Same arithmetic either way.
struct FlatDeps {
    const uint32_t* off;
    const uint32_t* src;
    const uint8_t* low;
};
__global__ void k_flat_count(View e, const uint32_t* list, uint32_t n, uint32_t* cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t s = list[i];
        cnt[i] = e.dep_off[s + 1] - e.dep_off[s];
    }
    if (i == n) cnt[i] = 0;
}
__global__ void k_flat_fill(View e, const uint32_t* list, uint32_t n, const uint32_t* off, uint32_t* src, uint8_t* low) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = list[i], d0 = e.dep_off[s], nd = e.dep_off[s + 1] - d0, o = off[i];
    for (uint32_t j = 0; j < nd; ++j) src[o + j] = e.dep_ids[d0 + j];
    low[i] = (nd > 0 && e.svar[e.dep_ids[d0]] < e.svar[s]) ? 1 : 0;
}
template <class T, int G, int MAXC>
__global__ void k_rule_cat(View e, T* __restrict__ val, const uint32_t* list, uint32_t n, int rule,
                           const T* __restrict__ table, const T* __restrict__ table_t, int n_sym, T potts_w, const FlatDeps flat) {
    extern __shared__ unsigned char smem_raw[];
    T* sh = reinterpret_cast<T*>(smem_raw);
    if (*(volatile int*)e.err_flag) return;  // refused level: no value is written (uniform for the whole grid)
    const int K = e.dim;
    const int groups_per_block = blockDim.x / G;
    T* sh_table = sh;                                   // K*K (both orientations) when CAT_TABLE
    T* sh_in = sh + (rule == CXB_RULE_CAT_TABLE ? 2 * (size_t)K * K : 0);  // per group K staging
    if (rule == CXB_RULE_CAT_TABLE) {
        for (int t = threadIdx.x; t < K * K; t += blockDim.x) {
            sh_table[t] = table[t];            // [x_lo][x_hi]
            sh_table[K * K + t] = table_t[t];  // [x_hi][x_lo]
        }
    }
    __syncthreads();
    const int g = threadIdx.x / G, lane = threadIdx.x % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    // persistent blocks: the table is staged once per block and serves every batch of groups_per_block signals the block takes
    // (one block per 32 signals re-staged it 7,800 times per launch on the power-law graph)
    for (uint32_t i0 = blockIdx.x * groups_per_block; i0 < n; i0 += gridDim.x * groups_per_block) {
    const uint32_t i = i0 + g;
    const bool active = i < n;
    uint32_t s = active ? list[i] : 0;
    const uint32_t* dep_ids = flat.off ? flat.src : e.dep_ids;
    uint32_t off = 0, nd = 0;
    if (active) {
        if (flat.off) {
            off = flat.off[i];
            nd = flat.off[i + 1] - off;
        } else {
            off = e.dep_off[s];
            nd = e.dep_off[s + 1] - off;
        }
    }
    T acc[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) acc[c] = T(0);
    if (active && nd == 0) atomicOr(e.err_flag, ERR_RULE_ARG);
    if (active && nd > 0) {
        const T* a = val + (size_t)dep_ids[off] * K;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            int k = lane + c * G;
            if (k < K) acc[c] = a[k];
        }
    }
    if (rule < 0) {
        for (uint32_t j = 1; j < nd; ++j) {
            const T* b = val + (size_t)dep_ids[off + j] * K;
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
                int k = lane + c * G;
                if (k < K) acc[c] = acc[c] * b[k];
            }
        }
    } else if (rule == CXB_RULE_POTTS) {
        T part = T(0);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) part += acc[c];
        for (int o = G / 2; o > 0; o >>= 1) part += __shfl_xor_sync(gmask, part, o, G);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) acc[c] = part + potts_w * acc[c];  // sum_b in[b] + (e^beta - 1) in[a]
    } else if (rule == CXB_RULE_CAT_TABLE) {
        // stage the incoming message, then every lane contracts its own output components
        T* my_in = sh_in + (size_t)g * K;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            int k = lane + c * G;
            if (k < K) my_in[k] = acc[c];
        }
        __syncwarp(gmask);
        bool u_is_low = false;
        if (active && nd > 0) u_is_low = flat.off ? flat.low[i] != 0 : e.svar[dep_ids[off]] < e.svar[s];
        // u low : out[a] = sum_b psi[b][a] in[b] -> table  [b*K + a];  u high: out[a] = sum_b psi[a][b] in[b] -> table_t[b*K + a]
        const T* tb = sh_table + (u_is_low ? 0 : (size_t)K * K);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            int a = lane + c * G;
            T sum = T(0);
            if (a < K)
                for (int b = 0; b < K; ++b) sum += tb[b * K + a] * my_in[b];
            acc[c] = sum;
        }
        __syncwarp(gmask);
    } else if (rule == CXB_RULE_HMM_EMIT) {
        int o = 0;
        if (active && nd > 0) {
            o = (int)val[(size_t)dep_ids[off] * K];
            if (o < 0 || o >= n_sym) {
                atomicOr(e.err_flag, ERR_RULE_ARG);
                o = 0;
            }
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            int a = lane + c * G;
            acc[c] = (a < K) ? table[(size_t)a * n_sym + o] : T(0);
        }
    }
    // normalise to sum 1
    T part = T(0);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) part += acc[c];
    for (int o = G / 2; o > 0; o >>= 1) part += __shfl_xor_sync(gmask, part, o, G);
    if (active && nd > 0) {
        T* o = val + (size_t)s * K;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            int k = lane + c * G;
            if (k < K) o[k] = acc[c] / part;
        }
    }
    }  // batches of this block
}

// ---- closed-form plan: disjoint random-walk chains -----------------------------------------------------------------------
// A memoised schedule whose graph is a set of linear-Gaussian random-walk chains (test/inference_engine_tests.jl:436-462,
// canonical rules of SURVEY Appendix C) does not need its 2T-1 recorded levels: one thread per chain runs the forward
// filter and the backward smoother in registers and writes the same 6T-4 signals. The arithmetic is rule_small_one's,
// expression for expression (GAUSS_OBS, GAUSS_RW, left-to-right canonical sums in dependency order), so the values are
// bit-identical to the level schedule's. Index arrays (built once from the wiring) say where each signal lives in `val`.
struct ChainPlanArgs {
    const uint32_t *i_y, *i_obs, *i_pred, *i_fwd, *i_bwd, *i_back, *i_marg;  // [pos] signal ids; 0xFFFFFFFF = does not exist
    const void *par_r, *par_q;  // [pos] noise variance of lik_t / tr_t (engine dtype)
    const uint32_t* base;       // [n_chains] position of (chain, t = 0)
    const uint32_t* len;        // [n_chains] T of the chain
    uint32_t stride;            // position of (chain, t) = base[chain] + t * stride
    uint32_t n_chains;
    int along_t;                // positions and signal ids are contiguous along t (else along the chains)
};
// blockIdx.y = 0: forward filter (m2v(x_t, lik_t), m2v(x_t, tr_{t-1}), m2f(x_t, tr_t)); blockIdx.y = 1: backward recursion
// (m2v(x_t, tr_t), m2f(x_t, tr_{t-1})), which recomputes the observation message from y with the same expression instead of
// waiting for the forward recursion. The two recursions of a chain run concurrently; the marginals follow in k_chain_plan_marg.
//
// A block owns 32 chains and walks them in tiles of 32 steps (tau = steps from the recursion's start). The only serial part
// is the two divisions per step, so the tile is split in three phases and the block's W warps take tiles round-robin:
//   A (all lanes, off the critical path): gather y / noise through the index tables with the lanes along the direction in
//     which the positions AND the signal ids are contiguous (along_t: lane = step, else lane = chain), compute the
//     observation message, park (oL, oh, q) in shared memory as [step][chain];
//   B (lane = chain): wait for the (L, h) the previous tile handed over (named barrier, producer/consumer), run 32 steps
//     from shared memory, hand (L, h) to the next tile's warp;
//   C (all lanes, off the critical path): scatter the messages to `val` with the phase-A lane mapping.
// While one warp is in B the other W-1 prefetch / drain their tiles, so a chain advances at the latency of its arithmetic.
template <class T>
struct ChainTile {
    T oL[32][33], oh[32][33], q[32][33], pL[32][33], ph[32][33];  // [step in tile][chain lane], padded: both lane mappings conflict-free
};
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void store_pair(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store_pair(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
template <class T, int W>
__global__ void __launch_bounds__(W * 32) k_chain_plan(T* __restrict__ val, ChainPlanArgs a) {
    extern __shared__ __align__(16) unsigned char chain_smem[];
    ChainTile<T>* tiles = reinterpret_cast<ChainTile<T>*>(chain_smem);
    T* hand = reinterpret_cast<T*>(tiles + W);  // [2][32]: (L, h) of the 32 chains after the last finished tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool bwd = blockIdx.y == 1, along_t = a.along_t != 0;
    const uint32_t c0 = blockIdx.x * 32, NONE = 0xFFFFFFFFu, st = a.stride;
    const uint32_t nc = min(32u, a.n_chains - c0);
    const uint32_t my_len = (uint32_t)lane < nc ? a.len[c0 + lane] : 0u;
    const uint32_t my_base = (uint32_t)lane < nc ? a.base[c0 + lane] : 0u;
    const uint32_t n_tiles = (__reduce_max_sync(0xFFFFFFFFu, my_len) + 31) / 32;
    const T* __restrict__ pr_r = (const T*)a.par_r;
    const T* __restrict__ pr_q = (const T*)a.par_q;
    ChainTile<T>& S = tiles[warp];
    const int n_it = along_t ? (int)nc : 32;  // along_t: iteration = chain, lane = step; else iteration = step, lane = chain
    constexpr int CH = sizeof(T) == 4 ? 16 : 8;
    for (uint32_t k = warp; k < n_tiles; k += W) {
        const uint32_t tau0 = k * 32;
        // ---- A: gather + observation message (CH positions per lane in flight: two memory latencies per chunk)
        for (int i0 = 0; i0 < n_it; i0 += CH) {
            uint32_t iy[CH], io[CH];
            T rr[CH], qq[CH], yy[CH];
            bool on[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int i = i0 + u, cc = along_t ? i : lane, tt = along_t ? lane : i;
                const uint32_t len_c = __shfl_sync(0xFFFFFFFFu, my_len, cc & 31), base_c = __shfl_sync(0xFFFFFFFFu, my_base, cc & 31);
                const uint32_t tau = tau0 + tt;
                on[u] = i < n_it && tau < len_c;
                const uint32_t pos = on[u] ? base_c + (bwd ? len_c - 1 - tau : tau) * st : 0u;
                iy[u] = on[u] ? a.i_y[pos] : 0u;
                io[u] = on[u] && !bwd ? a.i_obs[pos] : NONE;
                rr[u] = on[u] ? pr_r[pos] : T(1);
                qq[u] = on[u] && tau > 0 ? (bwd ? pr_q[pos] : pr_q[pos - st]) : T(0);
            }
#pragma unroll
            for (int u = 0; u < CH; ++u) yy[u] = on[u] ? val[(size_t)iy[u] * 2] : T(0);
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                if (!on[u]) continue;
                const int i = i0 + u, cc = along_t ? i : lane, tt = along_t ? lane : i;
                const T p = rr[u];
                const T oL = T(1) / p, oh = yy[u] / p;  // GAUSS_OBS
                S.oL[tt][cc] = oL;
                S.oh[tt][cc] = oh;
                S.q[tt][cc] = qq[u];
                if (io[u] != NONE) store_pair(val + (size_t)io[u] * 2, oL, oh);
            }
        }
        __syncwarp();
        // ---- B: the recursion, lane = chain
        T L = 0, h = 0;
        if (k > 0) {
            named_bar_sync(1 + (int)((k - 1) % W), 64);
            L = hand[lane];
            h = hand[32 + lane];
        }
        const uint32_t n_here = my_len > tau0 ? min(32u, my_len - tau0) : 0u;
#pragma unroll 4
        for (uint32_t tt = 0; tt < n_here; ++tt) {
            const T oL = S.oL[tt][lane], oh = S.oh[tt][lane];
            if (tau0 + tt > 0) {  // GAUSS_RW from the neighbour's m2f = (L, h), then lik (+) it
                const T q = S.q[tt][lane];
                const T den = T(1) + q * L;
                const T pL = L / den, ph = h / den;
                S.pL[tt][lane] = pL;
                S.ph[tt][lane] = ph;
                L = oL + pL;
                h = oh + ph;
            } else {
                L = oL;
                h = oh;
            }
            S.oL[tt][lane] = L;
            S.oh[tt][lane] = h;
        }
        if (k + 1 < n_tiles) {
            hand[lane] = L;
            hand[32 + lane] = h;
            __threadfence_block();
            named_bar_arrive(1 + (int)(k % W), 64);
        }
        __syncwarp();
        // ---- C: scatter (index loads of a chunk first, then its stores)
        for (int i0 = 0; i0 < n_it; i0 += CH) {
            uint32_t irw[CH], iout[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int i = i0 + u, cc = along_t ? i : lane, tt = along_t ? lane : i;
                const uint32_t len_c = __shfl_sync(0xFFFFFFFFu, my_len, cc & 31), base_c = __shfl_sync(0xFFFFFFFFu, my_base, cc & 31);
                const uint32_t tau = tau0 + tt;
                const bool on = i < n_it && tau < len_c;
                const uint32_t pos = on ? base_c + (bwd ? len_c - 1 - tau : tau) * st : 0u;
                irw[u] = on && tau > 0 ? (bwd ? a.i_bwd[pos] : a.i_pred[pos]) : NONE;
                iout[u] = on ? (bwd ? a.i_back[pos] : a.i_fwd[pos]) : NONE;
            }
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int i = i0 + u, cc = (along_t ? i : lane) & 31, tt = (along_t ? lane : i) & 31;
                if (irw[u] != NONE) store_pair(val + (size_t)irw[u] * 2, S.pL[tt][cc], S.ph[tt][cc]);
                if (iout[u] != NONE) store_pair(val + (size_t)iout[u] * 2, S.oL[tt][cc], S.oh[tt][cc]);
            }
        }
        __syncwarp();
    }
}
// marginal(x_t) = lik (+) tr_{t-1} (+) tr_t, left to right: one thread per (chain, t) position
template <class T>
__global__ void k_chain_plan_marg(T* __restrict__ val, ChainPlanArgs a, size_t n_pos) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pos) return;
    const uint32_t io = a.i_obs[p], ip = a.i_pred[p], ib = a.i_bwd[p], ig = a.i_marg[p];
    T gL = val[(size_t)io * 2], gh = val[(size_t)io * 2 + 1];
    if (ip != 0xFFFFFFFFu) {
        gL = gL + val[(size_t)ip * 2];
        gh = gh + val[(size_t)ip * 2 + 1];
    }
    if (ib != 0xFFFFFFFFu) {
        gL = gL + val[(size_t)ib * 2];
        gh = gh + val[(size_t)ib * 2 + 1];
    }
    val[(size_t)ig * 2] = gL;
    val[(size_t)ig * 2 + 1] = gh;
}

// =================================================================================================================
struct RuleDef {
    int kind = CXB_RULE_NONE;
    std::vector<double> params;
};

// one recorded run of the level schedule (see k_replay_resident)
struct Memo {
    std::vector<int64_t> req_ids;
    DBuf<uint8_t> pre_props, post_props;
    DBuf<uint64_t> pre_nib, post_nib;
    DBuf<uint32_t> lists, desc;  // members level by level; descriptors (level, key, count, offset)
    DBuf<uint32_t> flat_off, flat_src;  // flattened dependency lists of the members (FlatDeps), per-level replay of categorical graphs
    DBuf<uint8_t> flat_low;
    bool flat_ok = false;
    std::vector<uint32_t> h_desc;
    uint32_t n_desc = 0, n_list = 0;
    cxb_update_stats stats{};
    unsigned long long kind_count[6] = {0, 0, 0, 0, 0, 0};
    uint64_t last_use = 0;
    size_t bytes = 0;
    int plan = 0;  // 0: replay the recorded levels; > 0: a closed-form plan computes the values (DeviceEngine::run_plan)
    int64_t prepared = -1;  // recorded from this prepared request (ids need not be compared again)
    bool seq_only = false;  // certification failed (or the level schedule refused): this (request, flag state) is answered sequentially
    bool certified = false; // the recorded level run was compared with the sequential executor's answer and found equal
};
template <class T>
inline void swap_buf(DBuf<T>& a, DBuf<T>& b) {
    std::swap(a.p, b.p);
    std::swap(a.cap, b.cap);
}

struct DeviceEngine {
    int device = 0, dtype = CXB_F64, dim = 1, family = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    HostGraph g;
    Csr csr;
    bool structure_dirty = true, device_live = false, rules_dirty = true, trace_on = false;
    bool host_state_valid = true;            // host mirrors (props, C/F nibbles, values) equal the device state
    std::vector<unsigned char> host_vals;    // value mirror, valid with host_state_valid
    std::vector<uint32_t> lnk_off, lnk_ids;  // linked signals per variable id (CSR), rebuilt when links change
    bool links_dirty = true;
    std::vector<uint32_t> host_mark;  // epoch-stamped scratch for the set_values independence check
    uint32_t host_mark_tag = 0;
    std::map<int32_t, RuleDef> rules;     // by factor type
    std::vector<int32_t> key_ftype;       // key-1 -> factor type
    std::vector<double> fparam;           // per id, NaN = unset
    std::vector<uint8_t> var_family;      // per id, 0xFF = the engine family (cxb_set_variable_families)
    bool has_var_family = false;
    DBuf<uint8_t> d_sfam;
    uint32_t req_epoch = 0, lvl_epoch = 0;
    size_t n_uploaded = 0;
    // schedules (cxb_set_schedule): see include/cortex_b200.h
    int schedule = CXB_SCHEDULE_AUTO, last_ran = 0;
    bool hand_wired = false;  // cxb_create_signal / cxb_add_dependency were used: AUTO runs the sequential executor
    bool strict = true;       // strict rules A / B / D / E of the level schedule (CXB_STRICT=0: the round-1 rules)
    bool has_nl = false;      // any non-listening dependency in the graph (strict rule E keeps marks only then)
    DBuf<uint32_t> d_mark2, d_nl_epoch, d_nl_list, d_link_off, d_hz_m, d_hz_l, d_seq_stack, d_seq_rec;
    DBuf<uint8_t> d_pend_req, d_answers;
    std::vector<uint32_t> h_hz_m, h_hz_l;
    // pre-request snapshot of the dynamic state: a refused request is rolled back (side-effect free) and may then be
    // answered by the sequential executor
    DBuf<uint8_t> d_snap_props;
    DBuf<uint64_t> d_snap_nib;
    DBuf<unsigned char> d_snap_val;
    bool snap_valid = false;
    std::vector<int64_t> tr_var;  // trace: TracedInferenceExecution.variable_id
    // memoised schedules (CXB_MEMO=0 turns them off)
    std::vector<std::unique_ptr<Memo>> memos;
    bool memo_on = true;
    uint64_t memo_clock = 0;
    std::unique_ptr<Memo> rec;  // the recording of the request that is running
    DBuf<int> d_memo_flags;
    HBuf<int> h_memo_flags;
    static constexpr size_t MAX_MEMOS = 4;
    DBuf<unsigned char> d_val_level;  // certification: the level schedule's values, kept while the sequential executor runs
    long long n_certified = 0, n_cert_failed = 0;

    // device state
    DBuf<uint32_t> d_dep_off, d_dep_ids, d_nib_off, d_lis_off, d_lis_ids, d_lis_slot, d_done, d_visit, d_probe;
    bool has_weak = false;  // any weak dependency in the graph: levels probe through done signals (bfs_visit)
    DBuf<uint64_t> d_nib;
    DBuf<uint8_t> d_lis_listen, d_props, d_kind, d_rkey;
    DBuf<uint32_t> d_front_epoch, d_key_base;
    std::vector<uint32_t> key_base;  // per rule key: start of its partition of the frontier buffer
    bool cur_use_keys = true;
    uint32_t cur_total = 0, n_links = 0;
    int bfs_guess = 2;               // BFS steps launched before the host looks (adapts to the graph)
    DBuf<int32_t> d_svar, d_sfac;
    DBuf<unsigned char> d_val, d_fparam, d_tables, d_stage_val;
    DBuf<uint32_t> d_list_a, d_list_b, d_front, d_key_count, d_req_marg, d_link_ids, d_stage_ids;
    DBuf<uint8_t> d_ready;
    DBuf<int> d_flags;  // [0] err flag, [1] scratch int
    DBuf<uint32_t> d_counters;
    DBuf<unsigned long long> d_kind_count;
    // resident level loop (k_update_resident): per-key rule table, result slots
    DBuf<int> d_key_rule;
    DBuf<double> d_key_param;
    DBuf<long long> d_res_out;
    HBuf<long long> h_res_out;
    std::vector<int> h_key_rule, h_key_nsym;
    std::vector<double> h_key_param;
    std::vector<long long> h_key_table;
    DBuf<long long> d_key_table;
    DBuf<int> d_key_nsym;
    HBuf<uint32_t> h_counts;
    HBuf<unsigned char> h_stage;
    HBuf<int> h_flags;
    HBuf<unsigned long long> h_kind_count;
    std::map<int32_t, std::pair<size_t, size_t>> table_off;  // factor type -> (offset, count) in d_tables (elements)

    // request
    const int64_t* req_src = nullptr;
    bool req_src_prepared = false;
    int64_t current_prepared = -1;
    std::vector<int64_t> req_ids;
    std::vector<uint32_t> h_req_marg, h_link_off, h_link_ids;
    bool req_uploaded = false;
    uint32_t n_req = 0;

    // trace of the last update
    std::vector<int64_t> tr_level, tr_sid, tr_ns;  // tr_ns: device time of the level (CUDA events), shared evenly by its members
    cudaEvent_t tr_ev0 = nullptr, tr_ev1 = nullptr;
    cxb_update_stats stats{};

    size_t esz() const { return dtype == CXB_F32 ? 4 : 8; }

    ~DeviceEngine() {
        if (tr_ev0) cudaEventDestroy(tr_ev0);
        if (tr_ev1) cudaEventDestroy(tr_ev1);
        if (stream) cudaStreamDestroy(stream);
    }

    int32_t init() {
        int count = 0;
        cudaError_t e0 = cudaGetDeviceCount(&count);
        if (e0 != cudaSuccess || count == 0) {
            err = "no CUDA device available (cortex_b200 has no CPU fallback)";
            return CXB_ERR_CUDA;
        }
        if (device < 0 || device >= count) {
            err = "bad device index";
            return CXB_ERR_BAD_ARG;
        }
        CXB_CUDA(cudaSetDevice(device));
        CXB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        if (const char* e = getenv("CXB_STRICT")) strict = atoi(e) != 0;
        if (const char* e = getenv("CXB_MEMO")) memo_on = atoi(e) != 0;
        CXB_CUDA(d_flags.reserve(4));
        CXB_CUDA(d_counters.reserve(8));
        CXB_CUDA(d_kind_count.reserve(8));
        CXB_CUDA(h_counts.reserve(1024));
        CXB_CUDA(h_flags.reserve(4));
        CXB_CUDA(h_kind_count.reserve(8));
        CXB_CUDA(cudaMemsetAsync(d_flags.p, 0, 4 * sizeof(int), stream));
        CXB_CUDA(cudaMemsetAsync(d_counters.p, 0, 8 * sizeof(uint32_t), stream));
        CXB_CUDA(cudaMemsetAsync(d_kind_count.p, 0, 8 * sizeof(unsigned long long), stream));
        return CXB_OK;
    }

    View view() {
        View v;
        v.n_sig = (uint32_t)g.n_sig();
        v.dim = dim;
        v.dep_off = d_dep_off.p;
        v.dep_ids = d_dep_ids.p;
        v.nib_off = d_nib_off.p;
        v.nib = d_nib.p;
        v.lis_off = d_lis_off.p;
        v.lis_ids = d_lis_ids.p;
        v.lis_slot = d_lis_slot.p;
        v.lis_listen = d_lis_listen.p;
        v.props = d_props.p;
        v.kind = d_kind.p;
        v.rkey = d_rkey.p;
        v.svar = d_svar.p;
        v.sfac = d_sfac.p;
        v.sfam = has_var_family ? d_sfam.p : nullptr;
        v.done_epoch = d_done.p;
        v.visit_epoch = d_visit.p;
        v.probe_epoch = has_weak ? d_probe.p : nullptr;
        v.mark2 = d_mark2.p;
        v.nl_epoch = (has_nl && strict) ? d_nl_epoch.p : nullptr;
        v.nl_list = d_nl_list.p;
        v.nl_cnt = d_counters.p + 4;
        v.strict = strict ? 1 : 0;
        v.front_epoch = d_front_epoch.p;
        v.front = d_front.p;
        v.key_base = d_key_base.p;
        v.key_cnt = d_key_count.p;
        v.use_keys = cur_use_keys ? 1 : 0;
        v.err_flag = d_flags.p;
        v.kind_count = d_kind_count.p;
        return v;
    }

    template <class T>
    int32_t up(DBuf<T>& d, const T* src, size_t n) {
        CXB_CUDA(d.reserve(n));
        if (n) CXB_CUDA(cudaMemcpyAsync(d.p, src, n * sizeof(T), cudaMemcpyHostToDevice, stream));
        return CXB_OK;
    }

    // bring the dynamic state back to the host mirrors before a structural change
    int32_t sync_host() {
        if (!device_live || host_state_valid) return CXB_OK;
        CXB_CUDA(cudaSetDevice(device));
        int32_t st = download_state(host_vals);
        if (st) return st;
        host_state_valid = true;
        return CXB_OK;
    }
    int32_t download_state(std::vector<unsigned char>& values_out) {
        size_t N = n_uploaded;
        CXB_CUDA(cudaMemcpyAsync(g.props.data(), d_props.p, N, cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaMemcpyAsync(csr.nib.data(), d_nib.p, csr.nib.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
        values_out.resize(N * dim * esz());
        if (N) CXB_CUDA(cudaMemcpyAsync(values_out.data(), d_val.p, values_out.size(), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        // csr still describes the uploaded structure: only the first E_old log entries existed then
        HostGraph& gg = g;
        size_t E_old = csr.edge_pos.size();
        for (size_t e = 0; e < E_old; ++e) {
            uint32_t s = (uint32_t)gg.e_sig[e];
            uint32_t slot = csr.edge_pos[e] - csr.dep_off[s];
            uint64_t nibble = (csr.nib[csr.nib_off[s] + (slot >> 4)] >> ((slot & 15) << 2)) & 0xF;
            gg.e_nib[e] = (uint8_t)(nibble & 0xC);
        }
        return CXB_OK;
    }

    int32_t build_keys() {
        // rule key per signal: 0 = family reduce, 1.. = m2v of factor type, 255 = no rule (Unspecified / Joint)
        std::map<int32_t, int> key_of;
        key_ftype.clear();
        for (int64_t f : g.factors)
            if (!key_of.count(g.ftype[f])) {
                key_of[g.ftype[f]] = (int)key_ftype.size() + 1;
                key_ftype.push_back(g.ftype[f]);
            }
        if (key_ftype.size() > 250) {
            err = "more than 250 distinct factor types";
            return CXB_ERR_BAD_ARG;
        }
        size_t N = (size_t)g.n_sig();
        std::vector<uint8_t> rkey(N);
        std::vector<int32_t> svar(N), sfac(N);
        for (size_t s = 0; s < N; ++s) {
            uint8_t k = g.kind[s];
            svar[s] = (int32_t)g.svar[s];
            sfac[s] = (int32_t)g.sfac[s];
            if (k == CXB_KIND_M2F || k == CXB_KIND_PRODUCT || k == CXB_KIND_MARGINAL)
                rkey[s] = KEY_COMBINE;
            else if ((k == CXB_KIND_M2V || k == CXB_KIND_JOINT) && g.sfac[s] >= 0 && g.sfac[s] < g.n_ids && g.is_factor[g.sfac[s]])
                rkey[s] = (uint8_t)key_of[g.ftype[g.sfac[s]]];  // a JointMarginal is computed by its factor's rule
            else
                rkey[s] = (uint8_t)key_no_rule();
        }
        key_base.assign((size_t)n_keys() + 1, 0);
        for (size_t s2 = 0; s2 < N; ++s2) ++key_base[(size_t)rkey[s2] + 1];
        for (int k = 0; k < n_keys(); ++k) key_base[k + 1] += key_base[k];
        int32_t st;
        if ((st = up(d_key_base, key_base.data(), key_base.size()))) return st;
        if ((st = up(d_rkey, rkey.data(), N))) return st;
        if ((st = up(d_svar, svar.data(), N))) return st;
        if ((st = up(d_sfac, sfac.data(), N))) return st;
        if ((st = up(d_kind, g.kind.data(), N))) return st;
        CXB_CUDA(cudaStreamSynchronize(stream));
        return CXB_OK;
    }

    int32_t upload_rules() {
        // tables: CAT_TABLE psi and its transpose, HMM emission; per-factor params
        std::vector<double> all;
        table_off.clear();
        for (auto& kv : rules) {
            const RuleDef& r = kv.second;
            if (r.kind == CXB_RULE_CAT_TABLE) {
                if ((int64_t)r.params.size() != (int64_t)dim * dim) {
                    err = "CAT_TABLE needs value_dim*value_dim parameters";
                    return CXB_ERR_BAD_ARG;
                }
                table_off[kv.first] = {all.size(), r.params.size()};
                all.insert(all.end(), r.params.begin(), r.params.end());
                for (int a = 0; a < dim; ++a)
                    for (int b = 0; b < dim; ++b) all.push_back(r.params[(size_t)b * dim + a]);  // transpose
            } else if (r.kind == CXB_RULE_PROGRAM) {
                if (r.params.empty() || r.params[0] < 0 || (size_t)r.params[0] + 1 > r.params.size()) {
                    err = "PROGRAM needs {n_consts, consts..., code...} parameters";
                    return CXB_ERR_BAD_ARG;
                }
                table_off[kv.first] = {all.size(), r.params.size()};
                all.insert(all.end(), r.params.begin(), r.params.end());
            } else if (r.kind == CXB_RULE_HMM_EMIT) {
                if (r.params.empty() || (int64_t)r.params.size() != 1 + (int64_t)dim * (int64_t)r.params[0]) {
                    err = "HMM_EMIT needs {M, E[K][M]} parameters";
                    return CXB_ERR_BAD_ARG;
                }
                table_off[kv.first] = {all.size(), r.params.size() - 1};
                all.insert(all.end(), r.params.begin() + 1, r.params.end());
            }
        }
        size_t n = all.size();
        std::vector<unsigned char> raw(std::max<size_t>(n, 1) * esz());
        for (size_t i = 0; i < n; ++i) {
            if (dtype == CXB_F32)
                ((float*)raw.data())[i] = (float)all[i];
            else
                ((double*)raw.data())[i] = all[i];
        }
        int32_t st;
        if ((st = up(d_tables, raw.data(), raw.size()))) return st;
        std::vector<unsigned char> fp((size_t)std::max<int64_t>(g.n_ids, 1) * esz());
        for (int64_t i = 0; i < g.n_ids; ++i) {
            double v = (i < (int64_t)fparam.size()) ? fparam[i] : NAN;
            if (dtype == CXB_F32)
                ((float*)fp.data())[i] = (float)v;
            else
                ((double*)fp.data())[i] = v;
        }
        if ((st = up(d_fparam, fp.data(), fp.size()))) return st;
        if (has_var_family) {  // per-signal family = family of the signal's variable
            const size_t N = (size_t)g.n_sig();
            std::vector<uint8_t> sfam(std::max<size_t>(N, 1), (uint8_t)family);
            for (size_t s2 = 0; s2 < N; ++s2) {
                const int64_t v = g.svar[s2];
                if (v >= 0 && v < (int64_t)var_family.size() && var_family[v] != 0xFF) sfam[s2] = var_family[v];
            }
            if ((st = up(d_sfam, sfam.data(), sfam.size()))) return st;
        }
        CXB_CUDA(cudaStreamSynchronize(stream));
        rules_dirty = false;
        ++rules_version;
        return CXB_OK;
    }

    // (re)build the CSR and upload everything; keeps the dynamic state across structural changes
    int32_t ensure_device() {
        if (!stream) {
            err = "engine not initialised";
            return CXB_ERR_STATE;
        }
        CXB_CUDA(cudaSetDevice(device));
        int32_t st;
        if (structure_dirty) {
            size_t N_old = 0;
            if (device_live) {
                if ((st = sync_host())) return st;
                N_old = n_uploaded;
            }
            std::vector<unsigned char>& old_vals = host_vals;
            build_csr(g, csr);
            size_t N = (size_t)g.n_sig(), E = csr.dep_ids.size();
            if ((st = up(d_dep_off, csr.dep_off.data(), N + 1))) return st;
            if ((st = up(d_dep_ids, csr.dep_ids.data(), E))) return st;
            if ((st = up(d_nib_off, csr.nib_off.data(), N + 1))) return st;
            if ((st = up(d_nib, csr.nib.data(), csr.nib.size()))) return st;
            if ((st = up(d_lis_off, csr.lis_off.data(), N + 1))) return st;
            if ((st = up(d_lis_ids, csr.lis_ids.data(), E))) return st;
            if ((st = up(d_lis_slot, csr.lis_slot.data(), E))) return st;
            if ((st = up(d_lis_listen, csr.lis_listen.data(), E))) return st;
            if ((st = up(d_props, g.props.data(), N))) return st;
            // values: keep the old ones, zero the new signals
            {
                std::vector<unsigned char> vals(std::max<size_t>(N, 1) * dim * esz(), 0);
                if (N_old) std::memcpy(vals.data(), old_vals.data(), std::min(old_vals.size(), vals.size()));
                if ((st = up(d_val, vals.data(), vals.size()))) return st;
                CXB_CUDA(cudaStreamSynchronize(stream));
            }
            size_t Np = std::max<size_t>(N, 1);
            CXB_CUDA(d_done.reserve(Np));
            CXB_CUDA(d_visit.reserve(Np));
            CXB_CUDA(d_front_epoch.reserve(Np));
            has_weak = false;
            for (uint8_t fl : g.e_flags) has_weak = has_weak || (fl & CXB_NIB_WEAK);
            if (has_weak) {
                CXB_CUDA(d_probe.reserve(Np));
                CXB_CUDA(cudaMemsetAsync(d_probe.p, 0, Np * 4, stream));
            }
            // a signal enters a level's lists once, once more with MULTI_BIT (strict rule D), and once more as a probe
            CXB_CUDA(d_list_a.reserve((has_weak ? 3 : 2) * Np));
            CXB_CUDA(d_list_b.reserve((has_weak ? 3 : 2) * Np));
            CXB_CUDA(d_front.reserve(Np));
            CXB_CUDA(d_mark2.reserve(Np));
            size_t n_nl = 0;
            for (uint8_t fl : g.e_flags) n_nl += (fl & E_LISTEN) ? 0 : 1;
            has_nl = n_nl > 0;
            if (has_nl) {
                CXB_CUDA(d_nl_epoch.reserve(Np));
                CXB_CUDA(d_nl_list.reserve(n_nl));
            }
            if ((st = reset_epochs())) return st;
            snap_valid = false;
            memos.clear();  // recorded schedules belong to the old structure
            lists.clear();  // prepared signal lists too (handles become invalid: prepare again after a structural change)
            chain_plan.tried = chain_plan.ok = false;
            if ((st = build_keys())) return st;
            CXB_CUDA(d_key_count.reserve(512));
            CXB_CUDA(cudaStreamSynchronize(stream));
            n_uploaded = N;
            structure_dirty = false;
            device_live = true;
            rules_dirty = true;
            req_uploaded = false;
        }
        if (rules_dirty && (st = upload_rules())) return st;
        host_state_valid = false;  // whatever runs next may mutate the device state
        return CXB_OK;
    }

    // epoch stamps replace clearing; before a counter can wrap (the level epoch has 29 bits in mark2) everything is cleared
    int32_t reset_epochs() {
        const size_t Np = std::max<size_t>((size_t)g.n_sig(), 1);
        CXB_CUDA(cudaMemsetAsync(d_done.p, 0, Np * 4, stream));
        CXB_CUDA(cudaMemsetAsync(d_visit.p, 0, Np * 4, stream));
        CXB_CUDA(cudaMemsetAsync(d_front_epoch.p, 0, Np * 4, stream));
        CXB_CUDA(cudaMemsetAsync(d_mark2.p, 0, Np * 4, stream));
        if (has_weak) CXB_CUDA(cudaMemsetAsync(d_probe.p, 0, Np * 4, stream));
        if (has_nl) CXB_CUDA(cudaMemsetAsync(d_nl_epoch.p, 0, Np * 4, stream));
        CXB_CUDA(cudaMemsetAsync(d_counters.p, 0, 8 * sizeof(uint32_t), stream));
        req_epoch = lvl_epoch = 0;
        return CXB_OK;
    }
    int32_t take_snapshot() {
        snap_valid = false;
        const size_t N = n_uploaded, vb = N * dim * esz();
        if (d_snap_props.reserve(std::max<size_t>(N, 1)) != cudaSuccess || d_snap_nib.reserve(std::max<size_t>(csr.nib.size(), 1)) != cudaSuccess ||
            d_snap_val.reserve(std::max<size_t>(vb, 1)) != cudaSuccess) {
            cudaGetLastError();  // not enough memory for a copy of the state: run without rollback
            return CXB_OK;
        }
        if (N) CXB_CUDA(cudaMemcpyAsync(d_snap_props.p, d_props.p, N, cudaMemcpyDeviceToDevice, stream));
        if (!csr.nib.empty()) CXB_CUDA(cudaMemcpyAsync(d_snap_nib.p, d_nib.p, csr.nib.size() * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
        if (vb) CXB_CUDA(cudaMemcpyAsync(d_snap_val.p, d_val.p, vb, cudaMemcpyDeviceToDevice, stream));
        snap_valid = true;
        return CXB_OK;
    }
    int32_t restore_snapshot() {
        if (!snap_valid) return CXB_OK;
        const size_t N = n_uploaded, vb = N * dim * esz();
        if (N) CXB_CUDA(cudaMemcpyAsync(d_props.p, d_snap_props.p, N, cudaMemcpyDeviceToDevice, stream));
        if (!csr.nib.empty()) CXB_CUDA(cudaMemcpyAsync(d_nib.p, d_snap_nib.p, csr.nib.size() * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
        if (vb) CXB_CUDA(cudaMemcpyAsync(d_val.p, d_snap_val.p, vb, cudaMemcpyDeviceToDevice, stream));
        CXB_CUDA(cudaMemsetAsync(d_flags.p, 0, 4 * sizeof(int), stream));
        CXB_CUDA(cudaMemsetAsync(d_counters.p, 0, 8 * sizeof(uint32_t), stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        return CXB_OK;
    }

    int n_keys() const { return (int)key_ftype.size() + 2; }
    int key_no_rule() const { return (int)key_ftype.size() + 1; }

    const uint32_t* rule_list_base = nullptr;  // where the members of the level being evaluated live (frontier buffer / a memo's lists)
    FlatDeps rule_flat{nullptr, nullptr, nullptr};  // the memo's flattened dependency lists, indexed like its lists (replay only)
    template <class T>
    int32_t launch_rules_t(uint32_t total) {
        View v = view();
        T* val = (T*)d_val.p;
        const int nk = n_keys();
        bool categorical = family == CXB_FAMILY_CATEGORICAL;
        if (!categorical && dim > 8) {
            err = "value_dim > 8 is only supported for the categorical family";
            return CXB_ERR_BAD_ARG;
        }
        for (int k = 0; k < nk; ++k) {
            uint32_t cnt = h_counts.p[k], off = h_counts.p[nk + k];
            if (!cnt) continue;
            if (k == key_no_rule()) {
                err = "Unprocessed signal variant (no rule for an Unspecified / JointMarginal signal)";
                return CXB_ERR_NO_RULE;
            }
            int rule = -1;
            const RuleDef* rd = nullptr;
            if (k != KEY_COMBINE) {
                auto it = rules.find(key_ftype[k - 1]);
                if (it == rules.end() || it->second.kind == CXB_RULE_NONE) {
                    err = "The function `compute_message_to_variable!` is not implemented for factor type " +
                          std::to_string(key_ftype[k - 1]);
                    return CXB_ERR_NO_RULE;
                }
                rd = &it->second;
                rule = rd->kind;
            }
            const uint32_t* list = rule_list_base + off;
            bool cat_rule = rule == CXB_RULE_CAT_TABLE || rule == CXB_RULE_POTTS || rule == CXB_RULE_HMM_EMIT;
            if (categorical && (rule < 0 || cat_rule)) {
                const T* tb = nullptr;
                const T* tbt = nullptr;
                int n_sym = 0;
                T w = T(0);
                if (rule == CXB_RULE_CAT_TABLE) {
                    tb = (const T*)d_tables.p + table_off[key_ftype[k - 1]].first;
                    tbt = tb + (size_t)dim * dim;
                } else if (rule == CXB_RULE_HMM_EMIT) {
                    tb = (const T*)d_tables.p + table_off[key_ftype[k - 1]].first;
                    n_sym = (int)rd->params[0];
                } else if (rule == CXB_RULE_POTTS) {
                    w = (T)(std::exp(rd->params.empty() ? 0.0 : rd->params[0]) - 1.0);
                }
                int G = 1;
                while (G < dim && G < 32) G <<= 1;
                int maxc = (dim + G - 1) / G;
                const int threads = 256;
                int gpb = threads / G;
                size_t smem = ((rule == CXB_RULE_CAT_TABLE ? 2 * (size_t)dim * dim : 0) + (size_t)gpb * dim) * sizeof(T);
                if (smem > 200 * 1024 || maxc > 16) {
                    err = "categorical value_dim too large for the generic engine (use the structured HMM engine)";
                    return CXB_ERR_BAD_ARG;
                }
                unsigned grid = std::min(cdiv(cnt, gpb), 148u * 8u);  // persistent: 8 blocks of 256 threads per SM
                FlatDeps fd{nullptr, nullptr, nullptr};
                if (rule_flat.off) fd = FlatDeps{rule_flat.off + off, rule_flat.src, rule_flat.low + off};
#define CAT_LAUNCH(GG, MC)                                                                                        \
    do {                                                                                                          \
        if (smem > 48 * 1024)                                                                                     \
            cudaFuncSetAttribute(k_rule_cat<T, GG, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        CXB_LAUNCH((k_rule_cat<T, GG, MC>), grid, threads, smem, stream, v, val, list, cnt, rule, tb, tbt, n_sym, w, fd); \
    } while (0)
                if (G == 1) CAT_LAUNCH(1, 1);
                else if (G == 2) CAT_LAUNCH(2, 1);
                else if (G == 4) CAT_LAUNCH(4, 1);
                else if (G == 8) CAT_LAUNCH(8, 1);
                else if (G == 16) CAT_LAUNCH(16, 1);
                else if (maxc == 1) CAT_LAUNCH(32, 1);
                else if (maxc == 2) CAT_LAUNCH(32, 2);
                else if (maxc <= 4) CAT_LAUNCH(32, 4);
                else if (maxc <= 8) CAT_LAUNCH(32, 8);
                else CAT_LAUNCH(32, 16);
#undef CAT_LAUNCH
            } else if (!categorical && !cat_rule) {
                T defp = (T)((rd && !rd->params.empty()) ? rd->params[0] : 1.0);
                const T* prog = nullptr;
                int n_prog = 0;
                if (rule == CXB_RULE_PROGRAM) {  // default parameter = consts[0]; the program sits with the tables
                    defp = (T)((rd->params.size() > 1 && rd->params[0] >= 1) ? rd->params[1] : 1.0);
                    prog = (const T*)d_tables.p + table_off[key_ftype[k - 1]].first;
                    n_prog = (int)table_off[key_ftype[k - 1]].second;
                }
                CXB_LAUNCH(k_rule_small<T>, cdiv(cnt, 256), 256, 0, stream, v, val, list, cnt, family, rule,
                           (const T*)d_fparam.p, defp, prog, n_prog);
            } else {
                err = "rule kind does not match the engine's value family";
                return CXB_ERR_BAD_ARG;
            }
        }
        (void)total;
        return CXB_OK;
    }
    // ---- frontier plumbing ------------------------------------------------------------------------------------
    // device counters: [0],[1] = ping-pong BFS list lengths; key cursors live in d_key_count[0..nk)
    int32_t begin_level(bool use_keys) {
        ++lvl_epoch;
        cur_use_keys = use_keys;
        CXB_CUDA(cudaMemsetAsync(d_key_count.p, 0, (size_t)n_keys() * sizeof(uint32_t), stream));
        CXB_CUDA(cudaMemsetAsync(d_counters.p, 0, 2 * sizeof(uint32_t), stream));
        return CXB_OK;
    }
    unsigned stride_grid(size_t n_max) const { return std::max(1u, std::min(cdiv(std::max<size_t>(n_max, 1), 256), 148u * 16u)); }

    // breadth-first reachability from the seeds in d_list_a (length in d_counters[0]): pushes pending signals to the
    // frontier. `bfs_guess` steps are launched back to back; the host only looks at the result afterwards.
    int32_t bfs(int visit_flags) {  // bit 0: skip (probe) signals done in this request, bit 1: first level of a request (strict rule A)
        View v = view();
        uint32_t N = (uint32_t)g.n_sig();
        int which = 0;  // list holding the current input
        for (;;) {
            for (int it = 0; it < bfs_guess; ++it) {
                uint32_t* in = which ? d_list_b.p : d_list_a.p;
                uint32_t* out = which ? d_list_a.p : d_list_b.p;
                CXB_CUDA(cudaMemsetAsync(d_counters.p + (which ^ 1), 0, sizeof(uint32_t), stream));
                CXB_LAUNCH(k_bfs, stride_grid(N), 256, 0, stream, v, in, d_counters.p + which, out, d_counters.p + (which ^ 1),
                           lvl_epoch, req_epoch, visit_flags);
                which ^= 1;
            }
            // one round trip: remaining BFS work + per-key frontier sizes + error flag
            CXB_CUDA(cudaMemcpyAsync(h_counts.p + 600, d_counters.p + which, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            int32_t st = fetch_frontier();
            if (st) return st;
            if (h_counts.p[600] == 0) break;
            ++bfs_guess;  // deeper than expected: keep going and remember
        }
        return CXB_OK;
    }
    // per-key frontier sizes (+ error flag) to the host; fills h_counts[0..nk) = counts, [nk..2nk) = segment bases
    int32_t fetch_frontier() {
        int nk = n_keys();
        CXB_CUDA(cudaMemcpyAsync(h_counts.p, d_key_count.p, nk * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaMemcpyAsync(h_flags.p, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        cur_total = 0;
        for (int k = 0; k < nk; ++k) {
            h_counts.p[nk + k] = cur_use_keys ? key_base[k] : 0;
            cur_total += h_counts.p[k];
        }
        return flags_to_status(h_flags.p[0]);
    }
    int32_t flags_to_status(int f) {
        if (!f) return CXB_OK;
        cudaMemsetAsync(d_flags.p, 0, sizeof(int), stream);
        if (f & ERR_RULE_ARG) {
            err = "rule kernel rejected its arguments (no dependencies, symbol out of range or unknown rule kind)";
            return CXB_ERR_NO_RULE;
        }
        if (f & ERR_SEQ_DIVERGED) {
            err = "update_marginals!: the sequential loop does not terminate on this request (signals keep refreshing each other round "
                  "after round)";
            return CXB_ERR_STATE;
        }
        if (f & ERR_SEQ_OVERFLOW) {
            err = "sequential schedule: the traversal is deeper than the number of signals (a cycle of intermediate dependencies; the reference "
                  "recurses without end here)";
            return CXB_ERR_STATE;
        }
        if (f & ERR_INDEPENDENCE)
            err = "level-synchronous schedule out of contract: frontier member depends on another member";
        else if (f & ERR_WORK_BENEATH_PENDING)
            err = "level-synchronous schedule out of contract: a requested marginal was already pending when the request arrived and there is "
                  "pending work beneath it (the reference gives it exactly one traversal: order-dependent)";
        else if (f & ERR_STALE_BENEATH)
            err = "level-synchronous schedule out of contract: a signal reached by the request is not pending but holds leftover freshness "
                  "from an earlier, incomplete request (order-dependent in the reference)";
        else if (f & ERR_REVISITED)
            err = "level-synchronous schedule out of contract: a pending signal reached more than once has a pending dependency "
                  "(order-dependent in the reference)";
        else if (f & ERR_NONLISTEN)
            err = "level-synchronous schedule out of contract: a non-listening notification decides a pending state "
                  "(in the reference it depends on the order inside this level / request)";
        else if (f & ERR_FINAL_ORDER)
            err = "level-synchronous schedule out of contract: a final-phase signal depends on another final-phase signal across the "
                  "reference's per-variable order (marginal, then linked signals, variable by variable)";
        else if (f & ERR_LEFTOVER_FRESH)
            err = "level-synchronous schedule out of contract: a requested marginal holds leftover freshness from an earlier, incomplete "
                  "request (order-dependent in the reference)";
        else if (f & ERR_WEAK_BENEATH_DONE)
            err = "level-synchronous schedule out of contract: a pending weak dependency lies beneath a signal already computed in "
                  "this request (the reference would recompute that signal: order-dependent)";
        else
            err = "level-synchronous schedule out of contract: a dependency is recomputed after its listener within one "
                  "request (order-dependent in the reference)";
        return CXB_ERR_OUT_OF_CONTRACT;
    }
    int32_t check_flags() {
        CXB_CUDA(cudaMemcpyAsync(h_flags.p, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        return flags_to_status(h_flags.p[0]);
    }

    int32_t launch_rules(uint32_t total, const uint32_t* list_base = nullptr, const Memo* flat_of = nullptr) {
        rule_list_base = list_base ? list_base : d_front.p;
        rule_flat = (flat_of && flat_of->flat_ok) ? FlatDeps{flat_of->flat_off.p, flat_of->flat_src.p, flat_of->flat_low.p} : FlatDeps{nullptr, nullptr, nullptr};
        return dtype == CXB_F32 ? launch_rules_t<float>(total) : launch_rules_t<double>(total);
    }

    // evaluate + apply the fetched level (segments described by h_counts). Errors raised by the kernels of this
    // level surface at the next fetch_frontier / check_flags.
    int32_t run_level(int check_mode, int64_t level_tag) {
        uint32_t total = cur_total;
        if (!total) return CXB_OK;
        View v = view();
        int32_t st;
        int nk = n_keys();
        std::vector<Segs> chunks;
        Segs sg{};
        for (int k = 0; k < nk; ++k) {
            if (!h_counts.p[k]) continue;
            if (sg.n == 32) {
                chunks.push_back(sg);
                sg = Segs{};
            }
            sg.base[sg.n] = h_counts.p[nk + k];
            sg.cnt[sg.n] = h_counts.p[k];
            sg.total += h_counts.p[k];
            ++sg.n;
        }
        if (sg.n) chunks.push_back(sg);
        if (trace_on) {
            if (!tr_ev0) {
                CXB_CUDA(cudaEventCreate(&tr_ev0));
                CXB_CUDA(cudaEventCreate(&tr_ev1));
            }
            CXB_CUDA(cudaEventRecord(tr_ev0, stream));
        }
        for (auto& c : chunks) CXB_LAUNCH(k_check_independent, cdiv(c.total, 256), 256, 0, stream, v, c, lvl_epoch, req_epoch, check_mode);
        if (rec) {  // record the level for the memoised schedule (the frontier buffer is reused by the next level)
            if ((size_t)rec->n_list + total > rec->lists.cap) {
                rec.reset();  // more executions than the recording has room for: give up on this one
            } else {
                const uint32_t level = rec->h_desc.empty() ? 0u : rec->h_desc[rec->h_desc.size() - 4] + 1;
                for (int k = 0; k < nk; ++k) {
                    const uint32_t cnt = h_counts.p[k];
                    if (!cnt) continue;
                    CXB_CUDA(cudaMemcpyAsync(rec->lists.p + rec->n_list, d_front.p + h_counts.p[nk + k], cnt * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                             stream));
                    const uint32_t d4[4] = {level, (uint32_t)k, cnt, rec->n_list};
                    rec->h_desc.insert(rec->h_desc.end(), d4, d4 + 4);
                    rec->n_list += cnt;
                }
            }
        }
        if ((st = launch_rules(total))) return st;  // rule and set_value! kernels return at once when a check refused the level
        for (auto& c : chunks) CXB_LAUNCH(k_apply, cdiv(c.total, 256), 256, 0, stream, v, c, req_epoch, check_mode);
        if (check_mode == 1 && v.nl_epoch) {  // strict rule E, second half
            CXB_LAUNCH(k_check_nl, std::min(cdiv(std::max<size_t>(d_nl_list.cap, 1), 256), 1184u), 256, 0, stream, v, req_epoch);
            CXB_CUDA(cudaMemsetAsync(d_counters.p + 4, 0, sizeof(uint32_t), stream));
        }
        stats.updates += total;
        if (trace_on) {
            CXB_CUDA(cudaEventRecord(tr_ev1, stream));
            std::vector<uint32_t> ids;
            for (int k = 0; k < nk; ++k) {
                uint32_t cnt = h_counts.p[k];
                if (!cnt) continue;
                size_t o = ids.size();
                ids.resize(o + cnt);
                CXB_CUDA(cudaMemcpyAsync(ids.data() + o, d_front.p + h_counts.p[nk + k], cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                         stream));
            }
            CXB_CUDA(cudaStreamSynchronize(stream));
            std::sort(ids.begin(), ids.end());
            float level_ms = 0.0f;
            CXB_CUDA(cudaEventElapsedTime(&level_ms, tr_ev0, tr_ev1));  // rules + set_value! side effects of this level, on the device
            const int64_t each = std::max<int64_t>((int64_t)(level_ms * 1e6 / std::max<size_t>(ids.size(), 1)), 1);
            for (uint32_t s : ids) {
                tr_level.push_back(level_tag);
                tr_sid.push_back(s);
                tr_ns.push_back(each);
                tr_var.push_back(g.svar[s]);
            }
        }
        return CXB_OK;
    }

    int32_t request(int64_t n, const int64_t* ids) {
        int32_t st = ensure_device();
        if (st) return st;
        if (n < 0 || (n > 0 && !ids)) {
            err = "request_inference_for: bad id list";
            return CXB_ERR_BAD_ARG;
        }
        bool same = req_uploaded && (int64_t)req_ids.size() == n &&
                    (n == 0 || (ids == req_src && req_src_prepared) || !std::memcmp(req_ids.data(), ids, n * 8));
        req_src = ids;
        req_src_prepared = current_prepared >= 0;  // a prepared request's id vector is immutable: same pointer = same ids
        if (!same) {
            // everything is validated and built in temporaries; the cached request is replaced only after a successful upload
            req_uploaded = false;
            std::vector<uint32_t> marg((size_t)n), loff((size_t)n + 1, 0), lids;
            // linked signals per variable, in link order (src/model_engine.jl:80-83)
            if (links_dirty) {
                lnk_off.assign((size_t)g.n_ids + 1, 0);
                for (auto& l : g.links) ++lnk_off[(size_t)l.first + 1];
                for (int64_t i = 0; i < g.n_ids; ++i) lnk_off[i + 1] += lnk_off[i];
                lnk_ids.resize(g.links.size());
                std::vector<uint32_t> cur(lnk_off.begin(), lnk_off.end() - 1);
                for (auto& l : g.links) lnk_ids[cur[(size_t)l.first]++] = (uint32_t)l.second;
                links_dirty = false;
            }
            for (int64_t i = 0; i < n; ++i) {
                int64_t v = ids[i];
                if (v < 0 || v >= g.n_ids || g.is_factor[v] || g.marg_of[v] < 0) {
                    err = "request_inference_for: not a variable id";
                    return CXB_ERR_BAD_ARG;
                }
                marg[i] = (uint32_t)g.marg_of[v];
                for (uint32_t k = lnk_off[v]; k < lnk_off[v + 1]; ++k) lids.push_back(lnk_ids[k]);
                loff[i + 1] = (uint32_t)lids.size();
            }
            std::vector<uint32_t> hzm, hzl;
            final_phase_hazards(marg, loff, lids, hzm, hzl);
            if ((st = up(d_req_marg, marg.data(), (size_t)n))) return st;
            if ((st = up(d_link_ids, lids.data(), lids.size()))) return st;
            if ((st = up(d_link_off, loff.data(), loff.size()))) return st;
            if ((st = up(d_hz_m, hzm.data(), hzm.size()))) return st;
            if ((st = up(d_hz_l, hzl.data(), hzl.size()))) return st;
            CXB_CUDA(d_ready.reserve(std::max<size_t>(n, 1)));
            CXB_CUDA(d_pend_req.reserve(std::max<size_t>(n, 1)));
            CXB_CUDA(cudaStreamSynchronize(stream));
            req_ids.assign(ids, ids + n);
            h_req_marg.swap(marg);
            h_link_ids.swap(lids);
            h_link_off.swap(loff);
            h_hz_m.swap(hzm);
            h_hz_l.swap(hzl);
            req_uploaded = true;
        }
        n_req = (uint32_t)n;
        n_links = (uint32_t)h_link_ids.size();
        if (req_epoch > 0xFFFF0000u || lvl_epoch > 0x1F000000u)  // before an epoch counter can wrap: clear the stamps
            if ((st = reset_epochs())) return st;
        ++req_epoch;
        CXB_CUDA(cudaMemsetAsync(d_ready.p, 0, std::max<size_t>(n, 1), stream));
        if (n_req + n_links)
            CXB_LAUNCH(k_request, cdiv((size_t)n_req + n_links, 256), 256, 0, stream, view(), d_req_marg.p, d_link_ids.p, n_req, n_links);
        return CXB_OK;
    }
    // Rule F of the level schedule (the oracle's update_lvl states it in the same words). The reference's final phase goes
    // variable by variable - marginal(v_i), then the linked signals of v_i (src/inference_engine.jl:610-628) - while a level
    // schedule computes all pending marginals, then all pending linked signals. The orders differ only when one final-phase
    // candidate depends on another across that order. From the request and the dependency lists (host, once per distinct
    // request) two lists are made: signals that must NOT be pending when the marginal level starts (F1: a marginal that a
    // linked signal of an EARLIER requested variable depends on; F2: a marginal that depends on a linked signal of an earlier
    // variable; F4: a requested marginal another requested marginal depends on) and when the linked level starts (F3: a
    // linked signal another linked signal depends on). The device tests them with k_check_not_pending.
    void final_phase_hazards(const std::vector<uint32_t>& marg, const std::vector<uint32_t>& loff, const std::vector<uint32_t>& lids,
                             std::vector<uint32_t>& hzm, std::vector<uint32_t>& hzl) {
        const size_t n = marg.size();
        if (n == 0) return;
        // kinds that can be a hazard target at all: dependencies are looked up only when their kind is among them
        unsigned link_kinds = 0;
        for (uint32_t l : lids) link_kinds |= 1u << g.kind[l];
        std::vector<std::pair<uint32_t, uint32_t>> req_pos, link_pos;  // (signal, last request position) / (signal, first position)
        auto build_req = [&]() {
            req_pos.reserve(n);
            for (size_t i = 0; i < n; ++i) req_pos.emplace_back(marg[i], (uint32_t)i);
            std::sort(req_pos.begin(), req_pos.end());
            size_t w = 0;
            for (size_t r = 0; r < req_pos.size(); ++r) {  // keep the LAST position of a repeated marginal
                if (w && req_pos[w - 1].first == req_pos[r].first) req_pos[w - 1].second = req_pos[r].second;
                else req_pos[w++] = req_pos[r];
            }
            req_pos.resize(w);
        };
        auto build_link = [&]() {
            link_pos.reserve(lids.size());
            for (size_t i = 0; i < n; ++i)
                for (uint32_t k = loff[i]; k < loff[i + 1]; ++k) link_pos.emplace_back(lids[k], (uint32_t)i);
            std::sort(link_pos.begin(), link_pos.end());
            size_t w = 0;
            for (size_t r = 0; r < link_pos.size(); ++r)  // sorted: the FIRST position of a repeated signal comes first
                if (!w || link_pos[w - 1].first != link_pos[r].first) link_pos[w++] = link_pos[r];
            link_pos.resize(w);
        };
        auto find = [](const std::vector<std::pair<uint32_t, uint32_t>>& v, uint32_t sid) -> int64_t {
            auto it = std::lower_bound(v.begin(), v.end(), std::make_pair(sid, 0u));
            return (it != v.end() && it->first == sid) ? (int64_t)it->second : -1;
        };
        const unsigned marg_kind = 1u << CXB_KIND_MARGINAL;
        bool req_built = false, link_built = false;
        for (size_t i = 0; i < n; ++i) {
            const uint32_t m = marg[i];
            for (uint32_t k = csr.dep_off[m]; k < csr.dep_off[m + 1]; ++k) {
                const uint32_t d = csr.dep_ids[k];
                const unsigned kb = 1u << g.kind[d];
                if ((kb & marg_kind) && d != m) {  // F4
                    if (!req_built) build_req(), req_built = true;
                    if (find(req_pos, d) >= 0) hzm.push_back(d);
                }
                if (kb & link_kinds) {  // F2
                    if (!link_built) build_link(), link_built = true;
                    const int64_t j = find(link_pos, d);
                    if (j >= 0 && j < (int64_t)i) hzm.push_back(m);
                }
            }
            for (uint32_t q = loff[i]; q < loff[i + 1]; ++q) {
                const uint32_t l = lids[q];
                for (uint32_t k = csr.dep_off[l]; k < csr.dep_off[l + 1]; ++k) {
                    const uint32_t d = csr.dep_ids[k];
                    const unsigned kb = 1u << g.kind[d];
                    if (kb & marg_kind) {  // F1
                        if (!req_built) build_req(), req_built = true;
                        const int64_t j = find(req_pos, d);
                        if (j > (int64_t)i) hzm.push_back(d);
                    }
                    if ((kb & link_kinds) && d != l) {  // F3
                        if (!link_built) build_link(), link_built = true;
                        if (find(link_pos, d) >= 0) hzl.push_back(d);
                    }
                }
            }
        }
        std::sort(hzm.begin(), hzm.end());
        hzm.erase(std::unique(hzm.begin(), hzm.end()), hzm.end());
        std::sort(hzl.begin(), hzl.end());
        hzl.erase(std::unique(hzl.begin(), hzl.end()), hzl.end());
    }

    // one level of the frontier: seeds -> BFS -> per-key segments on the host
    int32_t find_frontier(bool use_done, bool use_keys, bool first_level = false) {
        int32_t st = begin_level(use_keys);
        if (st) return st;
        const int flags = (use_done ? 1 : 0) | (first_level ? 2 : 0);
        if (first_level && strict && n_pend_at_req > 0) {
            // strict rule B: the marginals that were pending when the request arrived are traversed first, alone - the
            // reference gives such a variable one traversal and takes its marginal, so nothing may be pending beneath it
            CXB_LAUNCH(k_seeds, cdiv(n_req, 256), 256, 0, stream, d_req_marg.p, d_pend_req.p, (uint8_t)1, n_req, d_list_a.p, d_counters.p);
            if ((st = bfs(flags))) return st;
            if (cur_total) {
                err = "level-synchronous schedule out of contract: a requested marginal was already pending when the request arrived and "
                      "there is pending work beneath it (the reference gives it exactly one traversal: order-dependent)";
                return CXB_ERR_OUT_OF_CONTRACT;
            }
            CXB_CUDA(cudaMemsetAsync(d_counters.p, 0, 2 * sizeof(uint32_t), stream));
            CXB_LAUNCH(k_seeds, cdiv(n_req, 256), 256, 0, stream, d_req_marg.p, d_pend_req.p, (uint8_t)0, n_req, d_list_a.p, d_counters.p);
            return bfs(flags);
        }
        if (n_req) CXB_LAUNCH(k_seeds, cdiv(n_req, 256), 256, 0, stream, d_req_marg.p, d_ready.p, (uint8_t)0, n_req, d_list_a.p, d_counters.p);
        return bfs(flags);
    }
    int n_pend_at_req = 0;

    int32_t scan(std::vector<int64_t>& out) {
        if (!req_uploaded) {
            err = "scan_inference_request: no request";
            return CXB_ERR_STATE;
        }
        int32_t st = find_frontier(false, false);
        if (st) return st;
        std::vector<uint32_t> ids(cur_total);
        if (cur_total) CXB_CUDA(cudaMemcpy(ids.data(), d_front.p, cur_total * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        std::sort(ids.begin(), ids.end());  // reported in ascending signal id
        out.assign(ids.begin(), ids.end());
        return CXB_OK;
    }

    // the whole update on the device in one launch (k_update_resident): small graphs, small-value families, no trace
    bool resident_ok() {
        if (const char* e = getenv("CXB_ENGINE_RESIDENT"))
            if (!atoi(e)) return false;
        const bool small_family = family != CXB_FAMILY_CATEGORICAL && dim <= 8;
        const bool small_categorical = family == CXB_FAMILY_CATEGORICAL && dim <= 64;
        return !trace_on && (small_family || small_categorical) && g.n_sig() <= 65536;
    }
    int32_t update_resident(unsigned long long launches0) {
        const int nk = n_keys();
        int32_t st = upload_key_tables();
        if (st) return st;
        cur_use_keys = true;
        ResidentArgs a{};
        a.req_marg = d_req_marg.p;
        a.link_ids = d_link_ids.p;
        a.ready = d_ready.p;
        a.n_req = n_req;
        a.n_links = n_links;
        a.list_a = d_list_a.p;
        a.list_b = d_list_b.p;
        a.req_epoch = req_epoch;
        a.lvl_epoch0 = lvl_epoch;
        a.n_keys = nk;
        a.family = family;
        a.key_rule = d_key_rule.p;
        a.key_param = d_key_param.p;
        a.fparam = d_fparam.p;
        a.tables = d_tables.p;
        a.key_table = d_key_table.p;
        a.key_nsym = d_key_nsym.p;
        a.out = d_res_out.p;
        if (rec) {
            a.rec_list = rec->lists.p;
            a.rec_desc = rec->desc.p;
            a.rec_list_cap = (uint32_t)std::min<size_t>(rec->lists.cap, 0xFFFFFFF0u);
            a.rec_desc_cap = (uint32_t)std::min<size_t>(rec->desc.cap / 4, 0x3FFFFFF0u);
        }
        a.pend_at_req = d_pend_req.p;
        a.hz_m = d_hz_m.p;
        a.hz_l = d_hz_l.p;
        a.n_hz_m = (uint32_t)h_hz_m.size();
        a.n_hz_l = (uint32_t)h_hz_l.size();
        int threads = 1024;  // measured on the T = 1000 chain: 1024 threads 19 ms, 256 threads 27 ms (the per-level passes over the requested marginals dominate)
        if (const char* e = getenv("CXB_ENGINE_RESIDENT_THREADS")) threads = std::max(32, std::min(1024, atoi(e) / 32 * 32));
        if (dtype == CXB_F32)
            CXB_LAUNCH(k_update_resident<float>, 1, threads, 0, stream, view(), (float*)d_val.p, a);
        else
            CXB_LAUNCH(k_update_resident<double>, 1, threads, 0, stream, view(), (double*)d_val.p, a);
        CXB_CUDA(cudaMemcpyAsync(h_res_out.p, d_res_out.p, 8 * sizeof(long long), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaMemcpyAsync(h_kind_count.p, d_kind_count.p, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaMemcpyAsync(h_flags.p, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        lvl_epoch = (uint32_t)h_res_out.p[4];
        if (rec) {
            if (h_res_out.p[6] < 0) {
                rec.reset();
            } else {
                rec->n_desc = (uint32_t)h_res_out.p[6];
                rec->n_list = (uint32_t)h_res_out.p[7];
            }
        }
        stats.levels = h_res_out.p[0];
        stats.updates = h_res_out.p[1];
        stats.final_marginals = h_res_out.p[2];
        stats.final_linked = h_res_out.p[3];
        for (int k = 0; k < 6; ++k) stats.updates_by_kind[k] = (int64_t)h_kind_count.p[k];
        stats.kernel_launches = (int64_t)(g_kernel_launches - launches0);
        const int f = h_flags.p[0];
        if (f & ERR_NO_RULE_KEY) return no_rule_status();
        return flags_to_status(f);
    }

    // ---- memoised schedules -----------------------------------------------------------------------------------------------
    size_t flag_bytes() const { return n_uploaded + csr.nib.size() * sizeof(uint64_t); }
    void begin_recording() {
        rec.reset(new Memo());
        const size_t N = std::max<size_t>(n_uploaded, 1);
        // a signal runs at most once in the loop phase and once in the final phase; a descriptor holds >= 1 member
        if (rec->lists.reserve(2 * N) != cudaSuccess || (resident_ok() && rec->desc.reserve(4 * (2 * N + 8)) != cudaSuccess) ||
            rec->post_props.reserve(N) != cudaSuccess || rec->post_nib.reserve(std::max<size_t>(csr.nib.size(), 1)) != cudaSuccess) {
            cudaGetLastError();
            rec.reset();  // no room for a recording: run without
        }
    }
    // flattened dependency lists of every recorded member (FlatDeps): count, exclusive scan, fill
    void build_flat_deps(Memo& m) {
        const uint32_t n = m.n_list;
        View v = view();
        if (m.flat_off.reserve((size_t)n + 1) != cudaSuccess || m.flat_low.reserve(n) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        k_flat_count<<<cdiv((size_t)n + 1, 256), 256, 0, stream>>>(v, m.lists.p, n, m.flat_off.p);
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, m.flat_off.p, m.flat_off.p, (int)(n + 1), stream);
        DBuf<unsigned char> tmp;
        if (tmp.reserve(std::max<size_t>(tmp_bytes, 1)) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, m.flat_off.p, m.flat_off.p, (int)(n + 1), stream);
        uint32_t total = 0;
        if (cudaMemcpyAsync(&total, m.flat_off.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess || m.flat_src.reserve(std::max<size_t>(total, 1)) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        k_flat_fill<<<cdiv((size_t)n, 256), 256, 0, stream>>>(v, m.lists.p, n, m.flat_off.p, m.flat_src.p, m.flat_low.p);
        if (cudaGetLastError() != cudaSuccess) return;
        g_kernel_launches += 2;
        m.flat_ok = true;
    }
    int32_t commit_recording(int64_t n, const int64_t* ids) {
        Memo& m = *rec;
        if (last_ran != CXB_SCHEDULE_LEVEL) return CXB_OK;
        const size_t N = n_uploaded;
        if (!resident_ok()) {  // per-level path: the descriptors were collected on the host
            m.n_desc = (uint32_t)(m.h_desc.size() / 4);
            CXB_CUDA(m.desc.reserve(std::max<size_t>(m.h_desc.size(), 4)));
            if (!m.h_desc.empty())
                CXB_CUDA(cudaMemcpyAsync(m.desc.p, m.h_desc.data(), m.h_desc.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        } else {  // resident path: they were written by the kernel; the per-level replay of large graphs is not used for these
            m.h_desc.clear();
        }
        if (N) CXB_CUDA(cudaMemcpyAsync(m.post_props.p, d_props.p, N, cudaMemcpyDeviceToDevice, stream));
        if (!csr.nib.empty()) CXB_CUDA(cudaMemcpyAsync(m.post_nib.p, d_nib.p, csr.nib.size() * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
        if (!resident_ok() && family == CXB_FAMILY_CATEGORICAL && m.n_list > 0 && !(getenv("CXB_MEMO_FLAT") && !atoi(getenv("CXB_MEMO_FLAT"))))
            build_flat_deps(m);  // best effort: without them the replay walks the CSR
        CXB_CUDA(cudaStreamSynchronize(stream));
        // the rollback snapshot taken before the request IS the pre-state: hand its flag buffers over
        swap_buf(m.pre_props, d_snap_props);
        swap_buf(m.pre_nib, d_snap_nib);
        snap_valid = false;
        m.req_ids.assign(ids, ids + n);
        m.prepared = current_prepared;
        m.stats = stats;
        for (int k = 0; k < 6; ++k) m.kind_count[k] = h_kind_count.p[k];
        m.last_use = ++memo_clock;
        m.bytes = 2 * flag_bytes() + (m.lists.cap + m.desc.cap + m.flat_off.cap + m.flat_src.cap) * sizeof(uint32_t) + m.flat_low.cap;
        m.plan = recognise_plan(m);
        if (memos.size() >= MAX_MEMOS) {  // evict the least recently used
            size_t lru = 0;
            for (size_t i = 1; i < memos.size(); ++i)
                if (memos[i]->last_use < memos[lru]->last_use) lru = i;
            memos.erase(memos.begin() + lru);
        }
        memos.push_back(std::move(rec));
        return CXB_OK;
    }
    // the level schedule refused this request on this flag state (the rollback snapshot = the pre-state): a negative memo
    void remember_sequential_only(int64_t n, const int64_t* ids) {
        std::unique_ptr<Memo> m(new Memo());
        swap_buf(m->pre_props, d_snap_props);
        swap_buf(m->pre_nib, d_snap_nib);
        snap_valid = false;
        m->req_ids.assign(ids, ids + n);
        m->prepared = current_prepared;
        m->seq_only = true;
        m->last_use = ++memo_clock;
        if (memos.size() >= MAX_MEMOS) {
            size_t lru = 0;
            for (size_t i = 1; i < memos.size(); ++i)
                if (memos[i]->last_use < memos[lru]->last_use) lru = i;
            memos.erase(memos.begin() + lru);
        }
        memos.push_back(std::move(m));
    }
    // ---- certification --------------------------------------------------------------------------------------------------------
    // The contract checks of the level schedule are sufficient on the benchmark families and on everything the fuzzers of
    // round 1 produced, but not complete: tests/fuzz_bp_graphs.py (random default-resolver graphs with loops and hubs, messages
    // overwritten by the user, random links and request orders) still finds about 1 script in 50 on which an ACCEPTED request
    // differs from the reference (a later round of the reference re-enters signals computed earlier in the request). AUTO
    // therefore CERTIFIES a level run the first time it records it: the pre-request state is restored, the sequential executor
    // answers the same request, and the two observable states are compared on the device (k_state_equiv). Equal: the memo is
    // kept and later identical requests replay it - exactness then rests on "same request + same flag state => same schedule",
    // not on the rule set. Different: the sequential result stands and the (request, flag state) pair is remembered as
    // sequential-only. Graphs above CERTIFY_LIMIT signals (the benchmark-size graphs, wired by the default resolver and
    // driven by the proven protocols) are not certified: there the rules are the guarantee.
    // k_seq costs ~4 us per executed signal: 2^18 signals bound a certification to about a second. CXB_CERTIFY_LIMIT (signals)
    // moves the bound, CXB_CERTIFY=0 switches certification off.
    int64_t CERTIFY_LIMIT = getenv("CXB_CERTIFY_LIMIT") ? atoll(getenv("CXB_CERTIFY_LIMIT")) : (1 << 18);
    int32_t certify_last_memo(int64_t n, const int64_t* ids) {
        Memo& m = *memos.back();
        const size_t N = n_uploaded, NC = csr.nib.size(), vb = N * dim * esz();
        const cxb_update_stats level_stats = stats;
        unsigned long long level_kinds[6];
        for (int k = 0; k < 6; ++k) level_kinds[k] = h_kind_count.p[k];
        // keep the level result's values; flags of the level result = m.post_*; restore the pre-request state
        CXB_CUDA(d_val_level.reserve(std::max<size_t>(vb, 4)));
        if (vb) CXB_CUDA(cudaMemcpyAsync(d_val_level.p, d_val.p, vb, cudaMemcpyDeviceToDevice, stream));
        if (N) CXB_CUDA(cudaMemcpyAsync(d_props.p, m.pre_props.p, N, cudaMemcpyDeviceToDevice, stream));
        if (NC) CXB_CUDA(cudaMemcpyAsync(d_nib.p, m.pre_nib.p, NC * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
        if (vb) CXB_CUDA(cudaMemcpyAsync(d_val.p, d_snap_val.p, vb, cudaMemcpyDeviceToDevice, stream));  // the snapshot's values are still there
        CXB_CUDA(cudaMemsetAsync(d_kind_count.p, 0, 8 * sizeof(unsigned long long), stream));
        stats = cxb_update_stats{};
        const bool trace_was = trace_on;
        trace_on = false;
        int32_t st = update_seq(n, ids);
        trace_on = trace_was;
        if (st) {  // e.g. the reference loop does not terminate here: nothing stands; the engine is back at the pre-request state
            const std::string why = err;
            if (N) CXB_CUDA(cudaMemcpyAsync(d_props.p, m.pre_props.p, N, cudaMemcpyDeviceToDevice, stream));
            if (NC) CXB_CUDA(cudaMemcpyAsync(d_nib.p, m.pre_nib.p, NC * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
            if (vb) CXB_CUDA(cudaMemcpyAsync(d_val.p, d_snap_val.p, vb, cudaMemcpyDeviceToDevice, stream));
            CXB_CUDA(cudaMemsetAsync(d_flags.p, 0, 4 * sizeof(int), stream));
            CXB_CUDA(cudaStreamSynchronize(stream));
            memos.pop_back();
            err = why;
            return st;
        }
        View a = view(), b = view();
        b.props = m.post_props.p;
        b.nib = m.post_nib.p;
        CXB_CUDA(d_memo_flags.reserve(MAX_MEMOS));
        CXB_CUDA(h_memo_flags.reserve(MAX_MEMOS));
        CXB_CUDA(cudaMemsetAsync(d_memo_flags.p, 0, sizeof(int), stream));
        const unsigned grid = std::max(1u, std::min(cdiv(std::max<size_t>(N, vb / 4), 256), 148u * 8u));
        CXB_LAUNCH(k_state_equiv, grid, 256, 0, stream, a, b, (const uint32_t*)d_val.p, (const uint32_t*)d_val_level.p, vb / 4, d_memo_flags.p);
        CXB_CUDA(cudaMemcpyAsync(h_memo_flags.p, d_memo_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        if (!h_memo_flags.p[0]) {  // certified: the engine holds the (identical) state; report the level run
            ++n_certified;
            m.certified = true;
            stats = level_stats;
            for (int k = 0; k < 6; ++k) h_kind_count.p[k] = level_kinds[k];
            last_ran = CXB_SCHEDULE_LEVEL;
            return CXB_OK;
        }
        ++n_cert_failed;
        m.seq_only = true;  // the sequential result stands (stats and last_ran are update_seq's)
        m.lists.release();
        m.desc.release();
        m.flat_off.release();
        m.flat_src.release();
        m.flat_low.release();
        m.flat_ok = false;
        m.post_props.release();
        m.post_nib.release();
        m.plan = 0;
        return CXB_OK;
    }

    // ---- closed-form plans -----------------------------------------------------------------------------------------------
    // A plan computes the VALUES of a recorded schedule with a kernel written for the structure (the flags still come from
    // the recording). Plan 1: disjoint random-walk chains (k_chain_plan). 0 = none: replay the recorded levels.
    struct ChainPlan {
        bool tried = false, ok = false;
        DBuf<uint32_t> i_y, i_obs, i_pred, i_fwd, i_bwd, i_back, i_marg, base, len;
        DBuf<unsigned char> par_r, par_q;
        std::vector<int64_t> pos_lik, pos_tr;  // factor id per position (-1: none), to refresh the parameters
        std::vector<int64_t> xs_sorted;        // the state variables, ascending (the request must name exactly these)
        uint32_t stride = 1, n_chains = 0;
        bool along_t = true;
        size_t n_pos = 0;
        int64_t upd_m2v = 0, upd_m2f = 0, upd_marg = 0;
        uint64_t params_version = ~0ull;
    } chain_plan;
    uint64_t rules_version = 0;  // bumped by upload_rules

    bool build_chain_plan() {
        ChainPlan& P = chain_plan;
        P.tried = true;
        P.ok = false;
        if (family != CXB_FAMILY_GAUSS_CANON || dim != 2 || has_var_family || !g.links.empty()) return false;
        if ((int64_t)g.n_sig() != g.n_var + 2 * g.n_conn) return false;  // free signals: not the plain BP wiring
        auto rule_kind = [&](int64_t f) {
            auto it = rules.find(g.ftype[f]);
            return it == rules.end() ? (int)CXB_RULE_NONE : it->second.kind;
        };
        auto deg = [&](int64_t id) { return g.adj_off[id + 1] - g.adj_off[id]; };
        const uint32_t NONE = 0xFFFFFFFFu;
        // classify: y (degree 1, on a GAUSS_OBS factor), x (one GAUSS_OBS factor + at most two GAUSS_RW factors)
        std::vector<int64_t> lik_of(g.n_ids, -1), y_of(g.n_ids, -1), tr_a(g.n_ids, -1), tr_b(g.n_ids, -1);
        for (int64_t f : g.factors) {
            if (deg(f) != 2) return false;
            const int64_t u = g.adj_nbr[g.adj_off[f]], w = g.adj_nbr[g.adj_off[f] + 1];
            const int kind = rule_kind(f);
            if (kind == CXB_RULE_GAUSS_OBS) {
                const bool u_is_y = deg(u) == 1, w_is_y = deg(w) == 1;
                if (u_is_y == w_is_y) return false;
                const int64_t x = u_is_y ? w : u, y = u_is_y ? u : w;
                if (lik_of[x] >= 0) return false;
                lik_of[x] = f;
                y_of[x] = y;
            } else if (kind == CXB_RULE_GAUSS_RW) {
                for (int64_t x : {u, w}) {
                    if (tr_a[x] < 0) tr_a[x] = f;
                    else if (tr_b[x] < 0) tr_b[x] = f;
                    else return false;
                }
            } else {
                return false;
            }
        }
        std::vector<int64_t> xs;
        for (int64_t v : g.variables) {
            if (lik_of[v] >= 0) {
                if (deg(v) != 1 + (tr_a[v] >= 0) + (tr_b[v] >= 0)) return false;
                xs.push_back(v);
            } else if (deg(v) != 1 || tr_a[v] >= 0) {
                return false;  // neither a state nor an observation variable
            }
        }
        if (xs.empty()) return false;
        // walk the paths from their ends; the end whose transition factor has the smaller id comes first
        auto other = [&](int64_t f, int64_t x) {
            const int64_t u = g.adj_nbr[g.adj_off[f]], w = g.adj_nbr[g.adj_off[f] + 1];
            return u == x ? w : u;
        };
        std::vector<uint8_t> seen(g.n_ids, 0);
        std::vector<std::vector<int64_t>> chains;
        std::vector<int64_t> ends;
        for (int64_t x : xs)
            if (tr_b[x] < 0) ends.push_back(x);
        std::sort(ends.begin(), ends.end(), [&](int64_t p, int64_t q) { return std::make_pair(tr_a[p], p) < std::make_pair(tr_a[q], q); });
        for (int64_t e0 : ends) {
            if (seen[e0]) continue;
            std::vector<int64_t> path;
            int64_t x = e0, via = -1;
            while (true) {
                seen[x] = 1;
                path.push_back(x);
                const int64_t nf = tr_a[x] != via && tr_a[x] >= 0 ? tr_a[x] : (tr_b[x] != via && tr_b[x] >= 0 ? tr_b[x] : -1);
                if (nf < 0) break;
                const int64_t nx = other(nf, x);
                if (seen[nx]) return false;
                via = nf;
                x = nx;
            }
            chains.push_back(std::move(path));
        }
        size_t total = 0;
        for (auto& c : chains) total += c.size();
        if (total != xs.size()) return false;  // a cycle of transitions
        bool same_len = true;
        for (auto& c : chains) same_len = same_len && c.size() == chains[0].size();
        const size_t B = chains.size();
        P.n_chains = (uint32_t)B;
        P.n_pos = total;
        // Position layout = the direction in which the graph's own signal ids are contiguous, so that the kernel's gathers
        // and scatters coalesce: chain after chain (lanes along t) unless the chains were created interleaved.
        bool interleaved = false;
        if (same_len && B >= 32 && chains[0].size() > 1) {
            auto obs_sig = [&](size_t b, size_t t) { return (int64_t)g.m2v_of_conn(g.conn_of(chains[b][t], lik_of[chains[b][t]])); };
            interleaved = std::llabs(obs_sig(1, 0) - obs_sig(0, 0)) < std::llabs(obs_sig(0, 1) - obs_sig(0, 0));
        }
        P.along_t = !interleaved;
        P.stride = interleaved ? (uint32_t)B : 1u;
        std::vector<uint32_t> base(B), len(B), iy(total), io(total), ip(total, NONE), im(total, NONE), ib(total, NONE), ik(total, NONE), ig(total);
        P.pos_lik.assign(total, -1);
        P.pos_tr.assign(total, -1);
        size_t off = 0;
        P.upd_m2v = P.upd_m2f = P.upd_marg = 0;
        auto deps_are = [&](uint32_t s2, std::initializer_list<uint32_t> want) {
            size_t n = 0;
            for (uint32_t w : want) n += w != NONE;
            if (csr.dep_off[s2 + 1] - csr.dep_off[s2] != n) return false;
            size_t k = csr.dep_off[s2];
            for (uint32_t w : want)
                if (w != NONE && csr.dep_ids[k++] != w) return false;
            return true;
        };
        for (size_t b = 0; b < B; ++b) {
            const auto& c = chains[b];
            const size_t Tn = c.size();
            base[b] = interleaved ? (uint32_t)b : (uint32_t)off;
            len[b] = (uint32_t)Tn;
            auto pos = [&](size_t t) { return interleaved ? t * B + b : off + t; };
            for (size_t t = 0; t < Tn; ++t) {
                const int64_t x = c[t], lik = lik_of[x];
                const int64_t trn = t + 1 < Tn ? (other(tr_a[x], x) == c[t + 1] ? tr_a[x] : tr_b[x]) : -1;
                const int64_t trp = t > 0 ? (other(tr_a[x], x) == c[t - 1] ? tr_a[x] : tr_b[x]) : -1;
                const size_t p = pos(t);
                iy[p] = (uint32_t)g.m2f_of_conn(g.conn_of(y_of[x], lik));
                io[p] = (uint32_t)g.m2v_of_conn(g.conn_of(x, lik));
                ig[p] = (uint32_t)g.marg_of[x];
                if (trp >= 0) {
                    ip[p] = (uint32_t)g.m2v_of_conn(g.conn_of(x, trp));
                    ik[p] = (uint32_t)g.m2f_of_conn(g.conn_of(x, trp));
                }
                if (trn >= 0) {
                    ib[p] = (uint32_t)g.m2v_of_conn(g.conn_of(x, trn));
                    im[p] = (uint32_t)g.m2f_of_conn(g.conn_of(x, trn));
                }
                P.pos_lik[p] = lik;
                P.pos_tr[p] = trn;
            }
            // the wiring must be exactly the one the kernel's arithmetic stands for (dependency ORDER included)
            for (size_t t = 0; t < Tn; ++t) {
                const size_t p = pos(t);
                if (!deps_are(io[p], {iy[p]})) return false;
                if (t > 0 && !deps_are(ip[p], {im[pos(t - 1)]})) return false;
                if (t + 1 < Tn && !deps_are(ib[p], {ik[pos(t + 1)]})) return false;
                if (im[p] != NONE && !deps_are(im[p], {io[p], ip[p]})) return false;
                if (ik[p] != NONE && !deps_are(ik[p], {io[p], ib[p]})) return false;
                if (!deps_are(ig[p], {io[p], ip[p], ib[p]})) return false;
                if (csr.dep_off[iy[p] + 1] != csr.dep_off[iy[p]]) return false;  // the observation is an input
            }
            P.upd_m2v += (int64_t)(3 * Tn - 2);
            P.upd_m2f += (int64_t)(2 * Tn - 2);
            P.upd_marg += (int64_t)Tn;
            off += Tn;
        }
        auto upl = [&](DBuf<uint32_t>& d, const std::vector<uint32_t>& v) { return up(d, v.data(), v.size()) == CXB_OK; };
        if (!(upl(P.i_y, iy) && upl(P.i_obs, io) && upl(P.i_pred, ip) && upl(P.i_fwd, im) && upl(P.i_bwd, ib) && upl(P.i_back, ik) && upl(P.i_marg, ig) &&
              upl(P.base, base) && upl(P.len, len)))
            return false;
        if (P.par_r.reserve(total * esz()) != cudaSuccess || P.par_q.reserve(total * esz()) != cudaSuccess) return false;
        if (cudaStreamSynchronize(stream) != cudaSuccess) return false;
        P.xs_sorted = xs;
        std::sort(P.xs_sorted.begin(), P.xs_sorted.end());
        P.params_version = ~0ull;
        P.ok = true;
        return true;
    }
    int recognise_plan(const Memo& m) {
        if (getenv("CXB_PLAN") && !atoi(getenv("CXB_PLAN"))) return 0;
        if (!chain_plan.tried) build_chain_plan();
        const ChainPlan& P = chain_plan;
        if (!P.ok) return 0;
        std::vector<int64_t> ids(m.req_ids);
        std::sort(ids.begin(), ids.end());
        if (ids != P.xs_sorted) return 0;
        // the recorded run must have executed exactly the signals the kernel writes
        if (m.stats.updates_by_kind[CXB_KIND_M2V] != P.upd_m2v || m.stats.updates_by_kind[CXB_KIND_M2F] != P.upd_m2f ||
            m.stats.updates_by_kind[CXB_KIND_MARGINAL] != P.upd_marg || m.stats.updates != P.upd_m2v + P.upd_m2f + P.upd_marg)
            return 0;
        return 1;
    }
    int32_t run_plan(const Memo&) {
        ChainPlan& P = chain_plan;
        if (P.params_version != rules_version) {  // noise variances: per-factor parameter, else the rule's default
            std::vector<unsigned char> r(P.n_pos * esz()), q(P.n_pos * esz());
            auto param = [&](int64_t f) {
                if (f < 0) return 0.0;
                const double v = f < (int64_t)fparam.size() ? fparam[f] : NAN;
                if (v == v) return v;
                auto it = rules.find(g.ftype[f]);
                return (it != rules.end() && !it->second.params.empty()) ? it->second.params[0] : 1.0;
            };
            for (size_t p = 0; p < P.n_pos; ++p) {
                const double rv = param(P.pos_lik[p]), qv = param(P.pos_tr[p]);
                if (dtype == CXB_F32) {
                    ((float*)r.data())[p] = (float)rv;
                    ((float*)q.data())[p] = (float)qv;
                } else {
                    ((double*)r.data())[p] = rv;
                    ((double*)q.data())[p] = qv;
                }
            }
            CXB_CUDA(cudaMemcpyAsync(P.par_r.p, r.data(), r.size(), cudaMemcpyHostToDevice, stream));
            CXB_CUDA(cudaMemcpyAsync(P.par_q.p, q.data(), q.size(), cudaMemcpyHostToDevice, stream));
            CXB_CUDA(cudaStreamSynchronize(stream));
            P.params_version = rules_version;
        }
        ChainPlanArgs a{};
        a.i_y = P.i_y.p;
        a.i_obs = P.i_obs.p;
        a.i_pred = P.i_pred.p;
        a.i_fwd = P.i_fwd.p;
        a.i_bwd = P.i_bwd.p;
        a.i_back = P.i_back.p;
        a.i_marg = P.i_marg.p;
        a.par_r = P.par_r.p;
        a.par_q = P.par_q.p;
        a.base = P.base.p;
        a.len = P.len.p;
        a.stride = P.stride;
        a.n_chains = P.n_chains;
        a.along_t = P.along_t ? 1 : 0;
        const dim3 grid(cdiv(P.n_chains, 32), 2);  // y: forward / backward recursion of the same chains, side by side
        constexpr int W = 4;
        if (dtype == CXB_F32) {
            const size_t smem = W * sizeof(ChainTile<float>) + 64 * sizeof(float);
            auto kf = k_chain_plan<float, W>;
            CXB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CXB_LAUNCH((k_chain_plan<float, W>), grid, W * 32, smem, stream, (float*)d_val.p, a);
            CXB_LAUNCH(k_chain_plan_marg<float>, cdiv(P.n_pos, 256), 256, 0, stream, (float*)d_val.p, a, P.n_pos);
        } else {
            const size_t smem = W * sizeof(ChainTile<double>) + 64 * sizeof(double);
            auto kf = k_chain_plan<double, W>;
            CXB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CXB_LAUNCH((k_chain_plan<double, W>), grid, W * 32, smem, stream, (double*)d_val.p, a);
            CXB_LAUNCH(k_chain_plan_marg<double>, cdiv(P.n_pos, 256), 256, 0, stream, (double*)d_val.p, a, P.n_pos);
        }
        return CXB_OK;
    }

    // Same request and bit-identical flag state as a recorded run? Then replay it.
    int32_t try_replay(int64_t n, const int64_t* ids, bool& hit) {
        hit = false;
        std::vector<Memo*> cand;
        for (auto& m : memos)
            if ((int64_t)m->req_ids.size() == n && ((current_prepared >= 0 && m->prepared == current_prepared) || !std::memcmp(m->req_ids.data(), ids, (size_t)n * 8))) {
                if (current_prepared >= 0) m->prepared = current_prepared;  // next time the handle alone identifies the ids
                cand.push_back(m.get());
            }
        if (cand.empty()) return CXB_OK;
        const size_t N = n_uploaded, NC = csr.nib.size();
        CXB_CUDA(d_memo_flags.reserve(MAX_MEMOS));
        CXB_CUDA(h_memo_flags.reserve(MAX_MEMOS));
        CXB_CUDA(cudaMemsetAsync(d_memo_flags.p, 0, MAX_MEMOS * sizeof(int), stream));
        const unsigned grid = std::max(1u, std::min(cdiv(std::max(N, NC), 256), 148u * 8u));
        for (size_t c = 0; c < cand.size(); ++c)
            CXB_LAUNCH(k_state_differs, grid, 256, 0, stream, d_props.p, cand[c]->pre_props.p, N, d_nib.p, cand[c]->pre_nib.p, NC, d_memo_flags.p + c);
        CXB_CUDA(cudaMemcpyAsync(h_memo_flags.p, d_memo_flags.p, MAX_MEMOS * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        // under AUTO a graph small enough to be certified replays only what WAS certified (a memo recorded under an explicit
        // LEVEL schedule is not)
        const bool need_cert = schedule == CXB_SCHEDULE_AUTO && small_values() && g.n_sig() <= CERTIFY_LIMIT &&
                               !(getenv("CXB_CERTIFY") && !atoi(getenv("CXB_CERTIFY")));
        Memo* m = nullptr;
        for (size_t c = 0; c < cand.size() && !m; ++c)
            if (!h_memo_flags.p[c] && (!need_cert || cand[c]->certified || cand[c]->seq_only)) m = cand[c];
        if (!m) return CXB_OK;
        if (m->seq_only && schedule != CXB_SCHEDULE_AUTO) return CXB_OK;  // an explicit LEVEL request gets the level schedule (and its refusal)
        int32_t st;
        if (m->seq_only) {  // certification found the level schedule wrong for this request and flag state
            if ((st = update_seq(n, ids))) return st;
            m->last_use = ++memo_clock;
            hit = true;
            return CXB_OK;
        }
        if (m->plan > 0) {
            if ((st = run_plan(*m))) return st;
            last_ran = CXB_RAN_PLAN;
        } else {
            if ((st = replay_levels(*m))) return st;
            last_ran = CXB_RAN_REPLAY;
        }
        if (N) CXB_CUDA(cudaMemcpyAsync(d_props.p, m->post_props.p, N, cudaMemcpyDeviceToDevice, stream));
        if (NC) CXB_CUDA(cudaMemcpyAsync(d_nib.p, m->post_nib.p, NC * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
        if ((st = check_flags())) return st;  // a rule rejected its arguments (the only error a replay can meet)
        stats = m->stats;
        m->last_use = ++memo_clock;
        hit = true;
        return CXB_OK;
    }
    int32_t replay_levels(const Memo& m) {
        int32_t st;
        if (resident_ok()) {
            if ((st = upload_key_tables())) return st;
            ReplayArgs a{};
            a.list = m.lists.p;
            a.desc = m.desc.p;
            a.n_desc = m.n_desc;
            a.n_keys = n_keys();
            a.family = family;
            a.key_rule = d_key_rule.p;
            a.key_param = d_key_param.p;
            a.fparam = d_fparam.p;
            a.tables = d_tables.p;
            a.key_table = d_key_table.p;
            a.key_nsym = d_key_nsym.p;
            cur_use_keys = true;
            if (dtype == CXB_F32)
                CXB_LAUNCH(k_replay_resident<float>, 1, 1024, 0, stream, view(), (float*)d_val.p, a);
            else
                CXB_LAUNCH(k_replay_resident<double>, 1, 1024, 0, stream, view(), (double*)d_val.p, a);
            return CXB_OK;
        }
        // large graphs: one batched rule kernel per (level, rule key), back to back on the stream
        const int nk = n_keys();
        size_t d = 0;
        const size_t nd = m.h_desc.size() / 4;
        while (d < nd) {
            const uint32_t level = m.h_desc[4 * d];
            for (int k = 0; k < 2 * nk; ++k) h_counts.p[k] = 0;
            uint32_t total = 0;
            for (; d < nd && m.h_desc[4 * d] == level; ++d) {
                const uint32_t key = m.h_desc[4 * d + 1];
                h_counts.p[key] = m.h_desc[4 * d + 2];
                h_counts.p[nk + key] = m.h_desc[4 * d + 3];
                total += m.h_desc[4 * d + 2];
            }
            if ((st = launch_rules(total, m.lists.p, &m))) return st;
        }
        return CXB_OK;
    }

    // Can the sequential executor (and the resident level loop) evaluate this model's rules? (small fixed-size values,
    // categorical values up to 64 states)
    bool small_values() const {
        return (family != CXB_FAMILY_CATEGORICAL && dim <= 8) || (family == CXB_FAMILY_CATEGORICAL && dim <= 64);
    }
    static constexpr int64_t SEQ_FALLBACK_LIMIT = 1 << 20;  // signals: above it a refused request stays refused under AUTO (k_seq: ~microseconds per executed signal)

    // update_marginals!(engine, ids), src/inference_engine.jl:559-632: schedule selection (include/cortex_b200.h)
    int32_t update(int64_t n, const int64_t* ids) {
        const unsigned long long launches0 = g_kernel_launches;
        stats = cxb_update_stats{};
        tr_level.clear();
        tr_sid.clear();
        tr_ns.clear();
        tr_var.clear();
        last_ran = 0;
        int32_t st = ensure_device();
        if (st) return st;
        CXB_CUDA(cudaMemsetAsync(d_kind_count.p, 0, 8 * sizeof(unsigned long long), stream));
        if (schedule == CXB_SCHEDULE_SEQUENTIAL || (schedule == CXB_SCHEDULE_AUTO && hand_wired && small_values())) {
            if (!small_values()) {
                err = "the sequential schedule evaluates small fixed-size values and categorical values up to 64 states only";
                return CXB_ERR_BAD_ARG;
            }
            st = update_seq(n, ids);
            stats.kernel_launches = (int64_t)(g_kernel_launches - launches0);
            return st;
        }
        const bool memo_ok = memo_on && !trace_on && n > 0 && ids;
        if (memo_ok) {
            bool hit = false;
            if ((st = try_replay(n, ids, hit))) return st;
            if (hit) {
                stats.kernel_launches = (int64_t)(g_kernel_launches - launches0);
                return CXB_OK;
            }
        }
        if ((st = take_snapshot())) return st;  // a refused request is rolled back
        if (memo_ok && snap_valid) begin_recording();
        st = update_level(n, ids, launches0);
        if (rec) {
            const bool committed = st == CXB_OK && last_ran == CXB_SCHEDULE_LEVEL;
            if (st == CXB_OK) st = commit_recording(n, ids);
            rec.reset();
            if (st == CXB_OK && committed && schedule == CXB_SCHEDULE_AUTO && !memos.empty() && small_values() && g.n_sig() <= CERTIFY_LIMIT &&
                !(getenv("CXB_CERTIFY") && !atoi(getenv("CXB_CERTIFY"))))
                st = certify_last_memo(n, ids);
        }
        if (st == CXB_ERR_OUT_OF_CONTRACT && snap_valid) {
            const std::string why = err;
            int32_t st2 = restore_snapshot();
            if (st2) return st2;
            err = why;
            if (schedule == CXB_SCHEDULE_AUTO && small_values() && g.n_sig() <= SEQ_FALLBACK_LIMIT) {
                if (memo_ok) remember_sequential_only(n, ids);  // next time this request meets this flag state: straight to k_seq
                stats = cxb_update_stats{};
                tr_level.clear();
                tr_sid.clear();
                tr_ns.clear();
                CXB_CUDA(cudaMemsetAsync(d_kind_count.p, 0, 8 * sizeof(unsigned long long), stream));
                st = update_seq(n, ids);
            }
        }
        stats.kernel_launches = (int64_t)(g_kernel_launches - launches0);
        return st;
    }

    // per-key rule tables of the single-launch kernels (k_update_resident, k_seq)
    int32_t upload_key_tables() {
        const int nk = n_keys();
        h_key_rule.assign((size_t)nk, -2);
        h_key_param.assign((size_t)nk, 1.0);
        h_key_rule[KEY_COMBINE] = -1;
        h_key_table.assign((size_t)nk, -1);
        h_key_nsym.assign((size_t)nk, 0);
        for (int k = 1; k < nk - 1; ++k) {
            auto it = rules.find(key_ftype[k - 1]);
            if (it == rules.end() || it->second.kind == CXB_RULE_NONE) continue;
            const int kind = it->second.kind;
            const bool cat_rule = kind == CXB_RULE_CAT_TABLE || kind == CXB_RULE_POTTS || kind == CXB_RULE_HMM_EMIT;
            if (cat_rule != (family == CXB_FAMILY_CATEGORICAL)) {
                err = "rule kind does not match the engine's value family";
                return CXB_ERR_BAD_ARG;
            }
            h_key_rule[k] = kind;
            if (!it->second.params.empty()) h_key_param[k] = it->second.params[0];
            if (kind == CXB_RULE_POTTS) h_key_param[k] = std::exp(it->second.params.empty() ? 0.0 : it->second.params[0]) - 1.0;
            if (kind == CXB_RULE_CAT_TABLE || kind == CXB_RULE_HMM_EMIT) h_key_table[k] = (long long)table_off[key_ftype[k - 1]].first;
            if (kind == CXB_RULE_HMM_EMIT) h_key_nsym[k] = (int)it->second.params[0];
            if (kind == CXB_RULE_PROGRAM) {  // the program lives with the tables; key_nsym carries its length
                h_key_table[k] = (long long)table_off[key_ftype[k - 1]].first;
                h_key_nsym[k] = (int)table_off[key_ftype[k - 1]].second;
                h_key_param[k] = (it->second.params.size() > 1 && it->second.params[0] >= 1) ? it->second.params[1] : 1.0;
            }
        }
        int32_t st;
        if ((st = up(d_key_rule, h_key_rule.data(), (size_t)nk))) return st;
        if ((st = up(d_key_param, h_key_param.data(), (size_t)nk))) return st;
        if ((st = up(d_key_table, h_key_table.data(), (size_t)nk))) return st;
        if ((st = up(d_key_nsym, h_key_nsym.data(), (size_t)nk))) return st;
        CXB_CUDA(d_res_out.reserve(8));
        CXB_CUDA(h_res_out.reserve(8));
        CXB_CUDA(cudaMemsetAsync(d_res_out.p, 0, 8 * sizeof(long long), stream));
        return CXB_OK;
    }

    // one launch of the sequential executor (k_seq). mode 0: update_marginals!, 1: scan, 2: process_dependencies!(table)
    int32_t run_seq(int mode, uint32_t root, bool retry, const uint8_t* answers, size_t rec_cap) {
        int32_t st = upload_key_tables();
        if (st) return st;
        const size_t N = std::max<size_t>((size_t)g.n_sig(), 1);
        CXB_CUDA(d_seq_stack.reserve(3 * (N + 1)));
        CXB_CUDA(d_seq_rec.reserve(std::max<size_t>(rec_cap, 1)));
        if (answers) {
            CXB_CUDA(d_answers.reserve(N));
            CXB_CUDA(cudaMemcpyAsync(d_answers.p, answers, (size_t)g.n_sig(), cudaMemcpyHostToDevice, stream));
        }
        cur_use_keys = true;
        SeqArgs a{};
        a.req_marg = d_req_marg.p;
        a.link_off = d_link_off.p;
        a.link_ids = d_link_ids.p;
        a.ready = d_ready.p;
        a.n_req = mode == 2 ? 0u : n_req;
        a.stack = d_seq_stack.p;
        a.stack_cap = (uint32_t)N + 1;
        a.mode = mode;
        a.n_keys = n_keys();
        a.family = family;
        a.key_rule = d_key_rule.p;
        a.key_param = d_key_param.p;
        a.fparam = d_fparam.p;
        a.tables = d_tables.p;
        a.key_table = d_key_table.p;
        a.key_nsym = d_key_nsym.p;
        a.rec = d_seq_rec.p;
        a.rec_cap = (uint32_t)std::min<size_t>(rec_cap, 0xFFFFFFFFu);
        a.answers = answers ? d_answers.p : nullptr;
        a.root = root;
        a.retry = retry ? 1 : 0;
        a.max_updates = 8 * (long long)g.n_sig() + 4096 + 1;  // the oracle's seq_execution_cap
        a.out = d_res_out.p;
        if (dtype == CXB_F32)
            CXB_LAUNCH(k_seq<float>, 1, 32, 0, stream, view(), (float*)d_val.p, a);
        else
            CXB_LAUNCH(k_seq<double>, 1, 32, 0, stream, view(), (double*)d_val.p, a);
        CXB_CUDA(cudaMemcpyAsync(h_res_out.p, d_res_out.p, 8 * sizeof(long long), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaMemcpyAsync(h_kind_count.p, d_kind_count.p, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaMemcpyAsync(h_flags.p, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        return CXB_OK;
    }
    int32_t no_rule_status() {
        cudaMemsetAsync(d_flags.p, 0, sizeof(int), stream);
        const int k = (int)h_res_out.p[5];
        if (k == key_no_rule())
            err = "Unprocessed signal variant (no rule for an Unspecified / JointMarginal signal)";
        else
            err = "The function `compute_message_to_variable!` is not implemented for factor type " + std::to_string(key_ftype[k - 1]);
        return CXB_ERR_NO_RULE;
    }
    // the reference loop literally (k_seq); with a trace: the reference's round numbers, execution order and variable ids
    int32_t update_seq(int64_t n, const int64_t* ids) {
        int32_t st = request(n, ids);
        if (st) return st;
        const size_t N = std::max<size_t>((size_t)g.n_sig(), 1);
        const size_t rec_cap = trace_on ? 3 * std::max<size_t>(8 * N, 65536) : 0;
        if (trace_on) {
            if (!tr_ev0) {
                CXB_CUDA(cudaEventCreate(&tr_ev0));
                CXB_CUDA(cudaEventCreate(&tr_ev1));
            }
            CXB_CUDA(cudaEventRecord(tr_ev0, stream));
        }
        if ((st = run_seq(0, 0, true, nullptr, rec_cap))) return st;
        last_ran = CXB_SCHEDULE_SEQUENTIAL;
        stats.levels = h_res_out.p[0];
        stats.updates = h_res_out.p[1];
        stats.final_marginals = h_res_out.p[2];
        stats.final_linked = h_res_out.p[3];
        for (int k = 0; k < 6; ++k) stats.updates_by_kind[k] = (int64_t)h_kind_count.p[k];
        if (trace_on) {
            CXB_CUDA(cudaEventRecord(tr_ev1, stream));
            const size_t cnt = std::min<size_t>((size_t)h_res_out.p[6], rec_cap / 3);
            std::vector<uint32_t> rec(3 * cnt);
            if (cnt) CXB_CUDA(cudaMemcpyAsync(rec.data(), d_seq_rec.p, rec.size() * 4, cudaMemcpyDeviceToHost, stream));
            CXB_CUDA(cudaStreamSynchronize(stream));
            float ms = 0.0f;
            CXB_CUDA(cudaEventElapsedTime(&ms, tr_ev0, tr_ev1));
            const int64_t each = std::max<int64_t>((int64_t)(ms * 1e6 / std::max<size_t>(cnt, 1)), 1);
            for (size_t k = 0; k < cnt; ++k) {
                tr_level.push_back(rec[3 * k] == 0xFFFFFFFFu ? -1 : (int64_t)rec[3 * k]);
                tr_var.push_back(req_ids[rec[3 * k + 1]]);
                tr_sid.push_back(rec[3 * k + 2]);
                tr_ns.push_back(each);
            }
        }
        const int f = h_flags.p[0];
        if (f & ERR_NO_RULE_KEY) return no_rule_status();
        return flags_to_status(f);
    }
    // scan_inference_request in the reference's own order (duplicates included)
    int32_t scan_dfs(std::vector<int64_t>& out) {
        if (!req_uploaded) {
            err = "scan_inference_request: no request";
            return CXB_ERR_STATE;
        }
        if (!small_values()) {
            err = "scan in depth-first order needs the sequential executor (small values)";
            return CXB_ERR_BAD_ARG;
        }
        size_t cap = std::max<size_t>(4 * (size_t)g.n_sig(), 1024);
        for (;;) {
            int32_t st = run_seq(1, 0, true, nullptr, cap);
            if (st) return st;
            if ((st = flags_to_status(h_flags.p[0]))) return st;
            const size_t cnt = (size_t)h_res_out.p[6];
            if (cnt > cap) {  // scanning only caches is_pending: running it again with more room is harmless
                cap = cnt;
                continue;
            }
            std::vector<uint32_t> rec(cnt);
            if (cnt) CXB_CUDA(cudaMemcpy(rec.data(), d_seq_rec.p, cnt * 4, cudaMemcpyDeviceToHost));
            out.assign(rec.begin(), rec.end());
            return CXB_OK;
        }
    }
    int32_t process_dependencies_table(int64_t s, bool retry, const uint8_t* answers, std::vector<int64_t>& visited, bool& processed) {
        int32_t st = ensure_device();
        if (st) return st;
        size_t cap = std::max<size_t>(8 * (size_t)g.n_sig(), 1024);
        if ((st = run_seq(2, (uint32_t)s, retry, answers, cap))) return st;
        if ((st = flags_to_status(h_flags.p[0]))) return st;
        const size_t cnt = std::min<size_t>((size_t)h_res_out.p[6], cap);
        std::vector<uint32_t> rec(cnt);
        if (cnt) CXB_CUDA(cudaMemcpy(rec.data(), d_seq_rec.p, cnt * 4, cudaMemcpyDeviceToHost));
        visited.assign(rec.begin(), rec.end());
        processed = h_res_out.p[7] != 0;
        return CXB_OK;
    }

    // the level-synchronous schedule under its contract checks
    int32_t update_level(int64_t n, const int64_t* ids, unsigned long long launches0) {
        int32_t st = request(n, ids);
        if (st) return st;
        last_ran = CXB_SCHEDULE_LEVEL;
        CXB_CUDA(cudaMemsetAsync(d_flags.p + 2, 0, sizeof(int), stream));
        if (n_req) CXB_LAUNCH(k_request_check, cdiv(n_req, 256), 256, 0, stream, view(), d_req_marg.p, n_req, d_pend_req.p, d_flags.p + 2);
        if (resident_ok()) return update_resident(launches0);
        CXB_CUDA(cudaMemcpyAsync(h_flags.p + 2, d_flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost, stream));
        if ((st = check_flags())) return st;  // refused at request time: nothing is traversed or computed
        n_pend_at_req = h_flags.p[2];
        int64_t level = 0;
        while (n_req) {
            if ((st = find_frontier(true, true, level == 0))) return st;
            if (!cur_total) break;
            if ((st = run_level(1, level))) return st;
            CXB_LAUNCH(k_ready, cdiv(n_req, 256), 256, 0, stream, view(), d_req_marg.p, d_ready.p, n_req);
            ++stats.levels;
            ++level;
        }
        if ((st = check_flags())) return st;  // checks of the last level
        for (int mode = 0; mode < 2 && n_req; ++mode) {  // final phase: marginals, then linked signals
            uint32_t cnt = mode == 0 ? n_req : n_links;
            const std::vector<uint32_t>& hz = mode == 0 ? h_hz_m : h_hz_l;
            if (!hz.empty()) {  // rule F
                CXB_LAUNCH(k_check_not_pending, cdiv(hz.size(), 256), 256, 0, stream, view(), mode == 0 ? d_hz_m.p : d_hz_l.p, (uint32_t)hz.size());
                if ((st = check_flags())) return st;
            }
            if ((st = begin_level(true))) return st;
            if (cnt)
                CXB_LAUNCH(k_final_flags, cdiv(cnt, 256), 256, 0, stream, view(), d_req_marg.p, d_link_ids.p, n_req, n_links, mode, lvl_epoch);
            if ((st = fetch_frontier())) return st;
            (mode == 0 ? stats.final_marginals : stats.final_linked) = cur_total;
            if ((st = run_level(2, mode == 0 ? -1 : -2))) return st;
        }
        CXB_CUDA(cudaMemcpyAsync(h_kind_count.p, d_kind_count.p, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        if ((st = check_flags())) return st;
        for (int k = 0; k < 6; ++k) stats.updates_by_kind[k] = (int64_t)h_kind_count.p[k];
        return CXB_OK;
    }

    // ---- prepared signal lists and requests (the "prepare once, run many" form of the bulk calls) ----------------------
    // cxb_set_values / cxb_update_marginals take host id arrays and float64 values: at 10^6-10^7 signals per call the host
    // side (id validation, independence test, float64 -> engine dtype, staging, id upload) costs far more than the kernels.
    // A prepared list / request does that work ONCE and keeps the ids on the device; values then arrive in the engine dtype
    // from host OR device memory. (The reference's InferenceRequest object, src/inference_engine.jl:265-323, is the same idea.)
    struct PreparedList {
        DBuf<uint32_t> ids;
        uint32_t n = 0;
    };
    std::vector<std::unique_ptr<PreparedList>> lists;
    std::vector<std::vector<int64_t>> prepared_requests;
    int64_t prepare_list(int64_t n, const int64_t* sids) {
        if (ensure_device()) return -1;
        const int64_t N = g.n_sig();
        if (n <= 0 || !sids) {
            err = "prepare_signals: empty list";
            return -1;
        }
        if (host_mark.size() < (size_t)N) host_mark.assign((size_t)N, 0);
        if (++host_mark_tag == 0) {
            std::fill(host_mark.begin(), host_mark.end(), 0u);
            host_mark_tag = 1;
        }
        std::vector<uint32_t> ids32((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            if (sids[i] < 0 || sids[i] >= N) {
                err = "prepare_signals: bad signal id";
                return -1;
            }
            if (host_mark[sids[i]] == host_mark_tag) {
                err = "prepare_signals: a signal appears twice (sequential set_value! semantics need cxb_set_values)";
                return -1;
            }
            host_mark[sids[i]] = host_mark_tag;
            ids32[i] = (uint32_t)sids[i];
        }
        for (int64_t i = 0; i < n; ++i)
            for (uint32_t k = csr.dep_off[ids32[i]]; k < csr.dep_off[ids32[i] + 1]; ++k)
                if (host_mark[csr.dep_ids[k]] == host_mark_tag) {
                    err = "prepare_signals: the signals depend on each other (sequential set_value! semantics need cxb_set_values)";
                    return -1;
                }
        std::unique_ptr<PreparedList> L(new PreparedList());
        L->n = (uint32_t)n;
        if (up(L->ids, ids32.data(), ids32.size()) || cudaStreamSynchronize(stream) != cudaSuccess) return -1;
        lists.push_back(std::move(L));
        return (int64_t)lists.size() - 1;
    }
    // set_value! of every signal of a prepared list: values[n][value_dim] in the ENGINE dtype, host or device memory
    int32_t set_values_prepared(int64_t list, const void* values, bool on_device) {
        int32_t st = ensure_device();
        if (st) return st;
        if (list < 0 || list >= (int64_t)lists.size() || !values) {
            err = "set_values_prepared: bad list handle / null values";
            return CXB_ERR_BAD_ARG;
        }
        PreparedList& L = *lists[(size_t)list];
        const size_t bytes = (size_t)L.n * dim * esz();
        const void* src = values;
        if (!on_device) {
            CXB_CUDA(d_stage_val.reserve(bytes));
            CXB_CUDA(cudaMemcpyAsync(d_stage_val.p, values, bytes, cudaMemcpyHostToDevice, stream));
            src = d_stage_val.p;
        }
        const unsigned grid = cdiv((size_t)L.n * dim, 256);
        if (dtype == CXB_F32)
            CXB_LAUNCH(k_write_values<float>, grid, 256, 0, stream, (float*)d_val.p, dim, L.ids.p, (const float*)src, L.n);
        else
            CXB_LAUNCH(k_write_values<double>, grid, 256, 0, stream, (double*)d_val.p, dim, L.ids.p, (const double*)src, L.n);
        CXB_LAUNCH(k_apply_list, cdiv(L.n, 256), 256, 0, stream, view(), L.ids.p, L.n, req_epoch);
        if (!on_device) CXB_CUDA(cudaStreamSynchronize(stream));  // the caller's host buffer is free again
        return CXB_OK;
    }
    // get_value of every signal of a prepared list: out[n][value_dim] in the engine dtype, host or device memory
    int32_t get_values_prepared(int64_t list, void* out, bool on_device) {
        int32_t st = ensure_device();
        if (st) return st;
        if (list < 0 || list >= (int64_t)lists.size() || !out) {
            err = "get_values_prepared: bad list handle / null output";
            return CXB_ERR_BAD_ARG;
        }
        PreparedList& L = *lists[(size_t)list];
        const size_t bytes = (size_t)L.n * dim * esz();
        void* dst = out;
        if (!on_device) {
            CXB_CUDA(d_stage_val.reserve(bytes));
            dst = d_stage_val.p;
        }
        const unsigned grid = cdiv((size_t)L.n * dim, 256);
        if (dtype == CXB_F32)
            CXB_LAUNCH(k_gather_values<float>, grid, 256, 0, stream, (const float*)d_val.p, dim, L.ids.p, (float*)dst, L.n);
        else
            CXB_LAUNCH(k_gather_values<double>, grid, 256, 0, stream, (const double*)d_val.p, dim, L.ids.p, (double*)dst, L.n);
        if (!on_device) {
            CXB_CUDA(cudaMemcpyAsync(out, dst, bytes, cudaMemcpyDeviceToHost, stream));
            CXB_CUDA(cudaStreamSynchronize(stream));
        }
        return CXB_OK;
    }
    int64_t prepare_request(int64_t n, const int64_t* ids) {
        if (ensure_device()) return -1;
        if (n < 0 || (n > 0 && !ids)) {
            err = "prepare_request: bad id list";
            return -1;
        }
        for (int64_t i = 0; i < n; ++i)
            if (ids[i] < 0 || ids[i] >= g.n_ids || g.is_factor[ids[i]] || g.marg_of[ids[i]] < 0) {
                err = "prepare_request: not a variable id";
                return -1;
            }
        prepared_requests.emplace_back(ids, ids + n);
        return (int64_t)prepared_requests.size() - 1;
    }

    // bulk set_value!: sequential semantics; members that depend on each other are applied one by one
    int32_t set_values(int64_t n, const int64_t* sids, const double* values, int64_t stride) {
        int32_t st = ensure_device();
        if (st) return st;
        if (n == 0) return CXB_OK;
        int64_t N = g.n_sig();
        for (int64_t i = 0; i < n; ++i)
            if (sids[i] < 0 || sids[i] >= N) {
                err = "set_values: bad signal id";
                return CXB_ERR_BAD_ARG;
            }
        // members that depend on each other (or repeat) must be applied one by one, in order (sequential set_value!)
        bool independent = true;
        if (n > 1) {
            if (host_mark.size() < (size_t)N) host_mark.assign((size_t)N, 0);
            if (++host_mark_tag == 0) {
                std::fill(host_mark.begin(), host_mark.end(), 0u);
                host_mark_tag = 1;
            }
            for (int64_t i = 0; i < n && independent; ++i) {
                if (host_mark[sids[i]] == host_mark_tag) independent = false;
                host_mark[sids[i]] = host_mark_tag;
            }
            for (int64_t i = 0; i < n && independent; ++i) {
                uint32_t s = (uint32_t)sids[i];
                for (uint32_t k = csr.dep_off[s]; k < csr.dep_off[s + 1]; ++k)
                    if (host_mark[csr.dep_ids[k]] == host_mark_tag) {
                        independent = false;
                        break;
                    }
            }
        }
        if (!independent) {
            for (int64_t i = 0; i < n; ++i)
                if ((st = set_values_batch(1, sids + i, values + i * stride, stride))) return st;
            return CXB_OK;
        }
        return set_values_batch(n, sids, values, stride);
    }
    int32_t set_values_batch(int64_t n, const int64_t* sids, const double* values, int64_t stride) {
        size_t bytes = (size_t)n * dim * esz();
        CXB_CUDA(h_stage.reserve(bytes + (size_t)n * 4));
        uint32_t* hid = (uint32_t*)(h_stage.p);
        unsigned char* hv = h_stage.p + (size_t)n * 4;
        for (int64_t i = 0; i < n; ++i) {
            hid[i] = (uint32_t)sids[i];
            for (int k = 0; k < dim; ++k) {
                double v = values[i * stride + k];
                if (dtype == CXB_F32)
                    ((float*)hv)[i * dim + k] = (float)v;
                else
                    ((double*)hv)[i * dim + k] = v;
            }
        }
        CXB_CUDA(d_stage_ids.reserve((size_t)n));
        CXB_CUDA(d_stage_val.reserve(bytes));
        CXB_CUDA(cudaMemcpyAsync(d_stage_ids.p, hid, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
        CXB_CUDA(cudaMemcpyAsync(d_stage_val.p, hv, bytes, cudaMemcpyHostToDevice, stream));
        unsigned grid = cdiv((size_t)n * dim, 256);
        if (dtype == CXB_F32)
            CXB_LAUNCH(k_write_values<float>, grid, 256, 0, stream, (float*)d_val.p, dim, d_stage_ids.p, (const float*)d_stage_val.p,
                       (uint32_t)n);
        else
            CXB_LAUNCH(k_write_values<double>, grid, 256, 0, stream, (double*)d_val.p, dim, d_stage_ids.p,
                       (const double*)d_stage_val.p, (uint32_t)n);
        CXB_LAUNCH(k_apply_list, cdiv(n, 256), 256, 0, stream, view(), d_stage_ids.p, (uint32_t)n, req_epoch);
        CXB_CUDA(cudaStreamSynchronize(stream));  // the pinned staging buffer is reused by the next call
        return CXB_OK;
    }
    int32_t get_values(int64_t n, const int64_t* sids, double* out, int64_t stride) {
        int32_t st = ensure_device();
        if (st) return st;
        if (n == 0) return CXB_OK;
        int64_t N = g.n_sig();
        size_t bytes = (size_t)n * dim * esz();
        CXB_CUDA(h_stage.reserve(bytes + (size_t)n * 4));
        uint32_t* hid = (uint32_t*)(h_stage.p);
        unsigned char* hv = h_stage.p + (size_t)n * 4;
        for (int64_t i = 0; i < n; ++i) {
            if (sids[i] < 0 || sids[i] >= N) {
                err = "get_values: bad signal id";
                return CXB_ERR_BAD_ARG;
            }
            hid[i] = (uint32_t)sids[i];
        }
        CXB_CUDA(d_stage_ids.reserve((size_t)n));
        CXB_CUDA(d_stage_val.reserve(bytes));
        CXB_CUDA(cudaMemcpyAsync(d_stage_ids.p, hid, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
        unsigned grid = cdiv((size_t)n * dim, 256);
        if (dtype == CXB_F32)
            CXB_LAUNCH(k_gather_values<float>, grid, 256, 0, stream, (const float*)d_val.p, dim, d_stage_ids.p, (float*)d_stage_val.p,
                       (uint32_t)n);
        else
            CXB_LAUNCH(k_gather_values<double>, grid, 256, 0, stream, (const double*)d_val.p, dim, d_stage_ids.p,
                       (double*)d_stage_val.p, (uint32_t)n);
        CXB_CUDA(cudaMemcpyAsync(hv, d_stage_val.p, bytes, cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        for (int64_t i = 0; i < n; ++i)
            for (int k = 0; k < dim; ++k)
                out[i * stride + k] = dtype == CXB_F32 ? (double)((float*)hv)[i * dim + k] : ((double*)hv)[i * dim + k];
        return CXB_OK;
    }
    int32_t is_pending(int64_t s, int& out) {
        int32_t st = ensure_device();
        if (st) return st;
        CXB_LAUNCH(k_pending_single, 1, 1, 0, stream, view(), (uint32_t)s, d_flags.p + 1);
        CXB_CUDA(cudaMemcpyAsync(h_flags.p + 1, d_flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        out = h_flags.p[1];
        return CXB_OK;
    }
    int32_t is_computed(int64_t s, int& out) {
        int32_t st = ensure_device();
        if (st) return st;
        uint8_t p = 0;
        CXB_CUDA(cudaMemcpyAsync(&p, d_props.p + s, 1, cudaMemcpyDeviceToHost, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        out = (p & P_COMPUTED) ? 1 : 0;
        return CXB_OK;
    }
    // compute!(strategy = registered rule / family reduce, signal; force, skip_if_no_listeners), src/signal.jl:392-410
    int32_t compute(int64_t s, bool force, bool skip) {
        int32_t st = ensure_device();
        if (st) return st;
        if (skip && g.lis_count[s] == 0) return CXB_OK;
        int pend = 0;
        if ((st = is_pending(s, pend))) return st;
        if (!force && !pend) {
            err = "Signal is not pending. Cannot compute a non-pending signal. Use `force=true` to force computation.";
            return CXB_ERR_NOT_PENDING;
        }
        if (g.dep_count[s] == 0) {
            err = "compute!: the signal has no dependencies to reduce";
            return CXB_ERR_NO_RULE;
        }
        uint32_t id = (uint32_t)s;
        CXB_CUDA(cudaMemcpyAsync(d_front.p, &id, 4, cudaMemcpyHostToDevice, stream));
        CXB_CUDA(cudaStreamSynchronize(stream));
        // a bare compute! on an Unspecified signal uses the family reduce as strategy
        int key = g.kind[s] == CXB_KIND_M2V ? -2 : KEY_COMBINE;
        for (int k = 0; k < 2 * n_keys(); ++k) h_counts.p[k] = 0;
        if (key == -2) {
            if (g.sfac[s] < 0 || g.sfac[s] >= g.n_ids) {
                err = "compute!: the message has no factor to take a rule from";
                return CXB_ERR_NO_RULE;
            }
            for (size_t k = 0; k < key_ftype.size(); ++k)
                if (key_ftype[k] == g.ftype[g.sfac[s]]) key = (int)k + 1;
            if (key == -2) key = key_no_rule();
        }
        h_counts.p[key] = 1;
        if ((st = launch_rules(1))) return st;
        CXB_LAUNCH(k_apply_list, 1, 32, 0, stream, view(), d_front.p, 1u, req_epoch);
        return check_flags();
    }
};

}  // namespace cxb

// =================================================================================================================
// C ABI (include/cortex_b200.h) — generic engine part
// =================================================================================================================
using cxb::DeviceEngine;
static inline DeviceEngine* E(cxb_engine* h) { return reinterpret_cast<DeviceEngine*>(h); }
#define CHECK_SIG(h, s)                                             \
    if ((s) < 0 || (s) >= (int64_t)E(h)->g.n_sig()) {               \
        E(h)->err = "bad signal id";                                \
        return CXB_ERR_BAD_ARG;                                     \
    }

extern "C" {

const char* cxb_version(void) { return "cortex_b200 0.1.0 (sm_100a)"; }
uint64_t cxb_kernel_launches(void) { return (uint64_t)cxb::g_kernel_launches; }

int32_t cxb_create(int32_t device, int32_t dtype, int32_t value_dim, int32_t family, cxb_engine** out) try {
    if (!out || value_dim < 1 || (dtype != CXB_F32 && dtype != CXB_F64)) return CXB_ERR_BAD_ARG;
    *out = nullptr;
    DeviceEngine* e = new DeviceEngine();
    e->device = device;
    e->dtype = dtype;
    e->dim = value_dim;
    e->family = family;
    int32_t st = e->init();
    if (st) {
        fprintf(stderr, "cxb_create: %s\n", e->err.c_str());
        delete e;
        return st;
    }
    *out = reinterpret_cast<cxb_engine*>(e);
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
void cxb_destroy(cxb_engine* h) {
    if (h) {
        cudaSetDevice(E(h)->device);
        delete E(h);
    }
}
const char* cxb_last_error(cxb_engine* h) { return h ? E(h)->err.c_str() : "null handle"; }

int32_t cxb_graph_build(cxb_engine* h, int64_t n_ids, const uint8_t* is_factor, const int32_t* factor_type, int64_t n_edges,
                        const int64_t* edge_var, const int64_t* edge_fac) try {
    DeviceEngine* e = E(h);
    int32_t st = e->g.build(n_ids, is_factor, factor_type, n_edges, edge_var, edge_fac, e->err);
    if (st) return st;
    e->fparam.assign((size_t)n_ids, NAN);
    e->structure_dirty = true;
    return st;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_register_rule(cxb_engine* h, int32_t factor_type, int32_t rule_kind, const double* params, int64_t n_params) try {
    cxb::RuleDef r;
    r.kind = rule_kind;
    if (params && n_params > 0) r.params.assign(params, params + n_params);
    E(h)->rules[factor_type] = r;
    E(h)->rules_dirty = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_set_factor_params(cxb_engine* h, int64_t n, const int64_t* factor_ids, const double* values) try {
    DeviceEngine* e = E(h);
    for (int64_t i = 0; i < n; ++i) {
        int64_t f = factor_ids[i];
        if (f < 0 || f >= e->g.n_ids || !e->g.is_factor[f]) {
            e->err = "set_factor_params: not a factor id";
            return CXB_ERR_BAD_ARG;
        }
        e->fparam[f] = values[i];
    }
    e->rules_dirty = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_set_variable_families(cxb_engine* h, int64_t n, const int64_t* variable_ids, const int32_t* families) try {
    DeviceEngine* e = E(h);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t v = variable_ids[i];
        if (v < 0 || v >= e->g.n_ids || e->g.is_factor[v]) {
            e->err = "set_variable_families: not a variable id";
            return CXB_ERR_BAD_ARG;
        }
        if (families[i] < 0 || families[i] > CXB_FAMILY_POINT || families[i] == CXB_FAMILY_CATEGORICAL) {
            e->err = "set_variable_families: family must be one of the fixed-size (non-categorical) families";
            return CXB_ERR_BAD_ARG;
        }
        if ((int64_t)e->var_family.size() < e->g.n_ids) e->var_family.resize((size_t)e->g.n_ids, 0xFF);
        e->var_family[v] = (uint8_t)families[i];
    }
    if (n > 0) e->has_var_family = true;
    e->rules_dirty = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int64_t cxb_create_signal(cxb_engine* h) try {
    DeviceEngine* e = E(h);
    if (e->sync_host()) return -1;  // pull the dynamic state before growing the structure
    e->structure_dirty = true;
    e->hand_wired = true;
    return e->g.new_signal();
} CXB_ABI_CATCH(-1)
int32_t cxb_add_dependency(cxb_engine* h, int64_t s, int64_t d, int32_t flags) try {
    DeviceEngine* e = E(h);
    CHECK_SIG(h, s);
    CHECK_SIG(h, d);
    {
        int32_t st = e->sync_host();
        if (st) return st;
    }
    e->g.add_dependency((int32_t)s, (int32_t)d, flags & CXB_DEP_WEAK, !(flags & CXB_DEP_NO_LISTEN),
                        !(flags & CXB_DEP_NO_CHECK_COMPUTED), flags & CXB_DEP_INTERMEDIATE);
    e->structure_dirty = true;
    e->hand_wired = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_resolve_dependencies(cxb_engine* h, int32_t resolver) try {
    DeviceEngine* e = E(h);
    {
        int32_t st = e->sync_host();
        if (st) return st;
    }
    e->structure_dirty = true;
    return e->g.resolve(resolver, e->err);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_resolve_factor_dependencies(cxb_engine* h, int32_t resolver, int64_t factor_id) try {
    DeviceEngine* e = E(h);
    if (int32_t st = e->sync_host()) return st;
    e->structure_dirty = true;
    return e->g.resolve_one(resolver, factor_id, true, e->err);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_resolve_variable_dependencies(cxb_engine* h, int32_t resolver, int64_t variable_id) try {
    DeviceEngine* e = E(h);
    if (int32_t st = e->sync_host()) return st;
    e->structure_dirty = true;
    return e->g.resolve_one(resolver, variable_id, false, e->err);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_set_signal_variant(cxb_engine* h, int64_t s, int32_t kind, int64_t variable_id, int64_t factor_id) try {
    DeviceEngine* e = E(h);
    CHECK_SIG(h, s);
    if (kind < CXB_KIND_UNSPECIFIED || kind > CXB_KIND_JOINT) {
        e->err = "set_signal_variant: unknown kind";
        return CXB_ERR_BAD_ARG;
    }
    if (variable_id >= e->g.n_ids || factor_id >= e->g.n_ids || (variable_id >= 0 && e->g.is_factor[variable_id]) ||
        (factor_id >= 0 && !e->g.is_factor[factor_id])) {
        e->err = "set_signal_variant: bad variable / factor id";
        return CXB_ERR_BAD_ARG;
    }
    if ((kind == CXB_KIND_M2V || kind == CXB_KIND_M2F) && (variable_id < 0 || factor_id < 0)) {
        e->err = "set_signal_variant: a message variant needs a variable id and a factor id";
        return CXB_ERR_BAD_ARG;
    }
    if ((kind == CXB_KIND_MARGINAL || kind == CXB_KIND_PRODUCT) && variable_id < 0) {
        e->err = "set_signal_variant: the variant needs a variable id";
        return CXB_ERR_BAD_ARG;
    }
    if (variable_id < -1 || factor_id < -1) {
        e->err = "set_signal_variant: bad variable / factor id";
        return CXB_ERR_BAD_ARG;
    }
    if (int32_t st = e->sync_host()) return st;
    e->g.kind[s] = (uint8_t)kind;
    e->g.svar[s] = variable_id;
    e->g.sfac[s] = factor_id;
    e->structure_dirty = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_link_signal(cxb_engine* h, int64_t v, int64_t s) try {
    DeviceEngine* e = E(h);
    CHECK_SIG(h, s);
    if (v < 0 || v >= e->g.n_ids || e->g.is_factor[v]) {
        e->err = "link_signal: not a variable id";
        return CXB_ERR_BAD_ARG;
    }
    e->g.links.emplace_back(v, (int32_t)s);
    e->req_uploaded = false;
    e->links_dirty = true;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_link_signals(cxb_engine* h, int64_t n, const int64_t* vs, const int64_t* ss) try {
    for (int64_t i = 0; i < n; ++i) {
        int32_t st = cxb_link_signal(h, vs[i], ss[i]);
        if (st) return st;
    }
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int64_t cxb_n_signals(cxb_engine* h) try { return E(h)->g.n_sig(); } CXB_ABI_CATCH(-1)
int64_t cxb_signal_id(cxb_engine* h, int32_t kind, int64_t v, int64_t f) try {
    DeviceEngine* e = E(h);
    if (v < 0 || v >= e->g.n_ids || e->g.is_factor[v]) return -1;
    if (kind == CXB_KIND_MARGINAL) return e->g.marg_of[v];
    if (f < 0 || f >= e->g.n_ids || !e->g.is_factor[f]) return -1;
    int32_t c = e->g.conn_of(v, f);
    if (c < 0) return -1;
    if (kind == CXB_KIND_M2V) return e->g.m2v_of_conn(c);
    if (kind == CXB_KIND_M2F) return e->g.m2f_of_conn(c);
    return -1;
} CXB_ABI_CATCH(-1)
int32_t cxb_signal_info(cxb_engine* h, int64_t s, int64_t out[5]) try {
    CHECK_SIG(h, s);
    const cxb::HostGraph& g = E(h)->g;
    out[0] = g.kind[s];
    out[1] = g.svar[s];
    out[2] = g.sfac[s];
    out[3] = g.r0[s];
    out[4] = g.r1[s];
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
// dependency lists with the live 4-bit props: the structure comes from the log, the C/F bits from the device
int64_t cxb_get_dependencies(cxb_engine* h, int64_t s, int64_t* out_ids, uint8_t* out_nib, int64_t cap) try {
    DeviceEngine* e = E(h);
    if (s < 0 || s >= (int64_t)e->g.n_sig()) return -1;
    int64_t nd = e->g.dep_count[s];
    if (cap <= 0 || nd == 0) return nd;
    if (e->ensure_device()) return -1;
    uint32_t off = e->csr.dep_off[s], noff = e->csr.nib_off[s], nch = e->csr.nib_off[s + 1] - noff;
    std::vector<uint64_t> ch(nch);
    if (cudaMemcpy(ch.data(), e->d_nib.p + noff, nch * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (int64_t i = 0; i < nd && i < cap; ++i) {
        if (out_ids) out_ids[i] = e->csr.dep_ids[off + i];
        if (out_nib) out_nib[i] = (uint8_t)((ch[i >> 4] >> ((i & 15) << 2)) & 0xF);
    }
    return nd;
} CXB_ABI_CATCH(-1)
int64_t cxb_get_listeners(cxb_engine* h, int64_t s, int64_t* out_ids, uint8_t* out_listen, int64_t cap) try {
    DeviceEngine* e = E(h);
    if (s < 0 || s >= (int64_t)e->g.n_sig()) return -1;
    int64_t nl = e->g.lis_count[s];
    if (cap <= 0 || nl == 0) return nl;
    if (e->structure_dirty) {  // structure only: rebuild the CSR on the host, nothing needs the device
        if (e->ensure_device()) return -1;
    }
    uint32_t off = e->csr.lis_off[s];
    for (int64_t i = 0; i < nl && i < cap; ++i) {
        if (out_ids) out_ids[i] = e->csr.lis_ids[off + i];
        if (out_listen) out_listen[i] = e->csr.lis_listen[off + i];
    }
    return nl;
} CXB_ABI_CATCH(-1)
int64_t cxb_get_warnings(cxb_engine* h, int64_t* out, int64_t cap) try {
    const auto& w = E(h)->g.warnings;
    for (int64_t i = 0; i < (int64_t)w.size() && i < cap; ++i) out[i] = w[i];
    return (int64_t)w.size();
} CXB_ABI_CATCH(-1)
int32_t cxb_set_values(cxb_engine* h, int64_t n, const int64_t* sids, const double* values, int64_t stride) try {
    return E(h)->set_values(n, sids, values, stride);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_get_values(cxb_engine* h, int64_t n, const int64_t* sids, double* out, int64_t stride) try {
    return E(h)->get_values(n, sids, out, stride);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_is_pending(cxb_engine* h, int64_t s) try {
    if (s < 0 || s >= (int64_t)E(h)->g.n_sig()) return -1;
    int r = 0;
    if (E(h)->is_pending(s, r)) return -1;
    return r;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_is_computed(cxb_engine* h, int64_t s) try {
    if (s < 0 || s >= (int64_t)E(h)->g.n_sig()) return -1;
    int r = 0;
    if (E(h)->is_computed(s, r)) return -1;
    return r;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_compute(cxb_engine* h, int64_t s, int32_t force, int32_t skip) try {
    CHECK_SIG(h, s);
    return E(h)->compute(s, force != 0, skip != 0);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_request_inference(cxb_engine* h, int64_t n, const int64_t* ids) try {
    DeviceEngine* e = E(h);
    int32_t st = e->request(n, ids);
    if (st) return st;
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) return CXB_ERR_CUDA;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int64_t cxb_scan(cxb_engine* h, int64_t* out, int64_t cap) try {
    std::vector<int64_t> v;
    if (E(h)->scan(v)) return -1;
    for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[i];
    return (int64_t)v.size();
} CXB_ABI_CATCH(-1)
int32_t cxb_update_marginals(cxb_engine* h, int64_t n, const int64_t* ids, cxb_update_stats* stats) try {
    int32_t st = E(h)->update(n, ids);
    if (stats) *stats = E(h)->stats;
    return st;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int64_t cxb_prepare_signals(cxb_engine* h, int64_t n, const int64_t* signals) try { return E(h)->prepare_list(n, signals); } CXB_ABI_CATCH(-1)
int32_t cxb_set_values_prepared(cxb_engine* h, int64_t list, const void* values, int32_t values_on_device) try {
    return E(h)->set_values_prepared(list, values, values_on_device != 0);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_get_values_prepared(cxb_engine* h, int64_t list, void* out, int32_t out_on_device) try {
    return E(h)->get_values_prepared(list, out, out_on_device != 0);
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int64_t cxb_prepare_request(cxb_engine* h, int64_t n, const int64_t* variable_ids) try { return E(h)->prepare_request(n, variable_ids); } CXB_ABI_CATCH(-1)
int32_t cxb_update_marginals_prepared(cxb_engine* h, int64_t request, cxb_update_stats* stats) try {
    DeviceEngine* e = E(h);
    if (request < 0 || request >= (int64_t)e->prepared_requests.size()) {
        e->err = "update_marginals_prepared: bad request handle";
        return CXB_ERR_BAD_ARG;
    }
    const std::vector<int64_t>& ids = e->prepared_requests[(size_t)request];
    e->current_prepared = request;
    int32_t st = e->update((int64_t)ids.size(), ids.data());
    e->current_prepared = -1;
    if (stats) *stats = e->stats;
    return st;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
void* cxb_stream(cxb_engine* h) { return (void*)E(h)->stream; }
int32_t cxb_set_schedule(cxb_engine* h, int32_t schedule) try {
    if (schedule < CXB_SCHEDULE_AUTO || schedule > CXB_SCHEDULE_SEQUENTIAL) {
        E(h)->err = "set_schedule: unknown schedule";
        return CXB_ERR_BAD_ARG;
    }
    E(h)->schedule = schedule;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int32_t cxb_last_schedule(cxb_engine* h) try { return E(h)->last_ran; } CXB_ABI_CATCH(-1)
int64_t cxb_scan_dfs(cxb_engine* h, int64_t* out, int64_t cap) try {
    std::vector<int64_t> v;
    if (E(h)->scan_dfs(v)) return -1;
    for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[i];
    return (int64_t)v.size();
} CXB_ABI_CATCH(-1)
int64_t cxb_process_dependencies_table(cxb_engine* h, int64_t s, int32_t retry, const uint8_t* answers, int64_t* out_visited,
                                       int64_t cap, int32_t* processed_out) try {
    if (s < 0 || s >= (int64_t)E(h)->g.n_sig()) return -1;
    std::vector<int64_t> v;
    bool processed = false;
    if (E(h)->process_dependencies_table(s, retry != 0, answers, v, processed)) return -1;
    if (processed_out) *processed_out = processed ? 1 : 0;
    for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out_visited[i] = v[i];
    return (int64_t)v.size();
} CXB_ABI_CATCH(-1)
int64_t cxb_trace_get_variables(cxb_engine* h, int64_t* out, int64_t cap) try {
    DeviceEngine* e = E(h);
    for (int64_t i = 0; i < (int64_t)e->tr_var.size() && i < cap; ++i)
        if (out) out[i] = e->tr_var[i];
    return (int64_t)e->tr_var.size();
} CXB_ABI_CATCH(-1)
int32_t cxb_trace_enable(cxb_engine* h, int32_t on) try {
    E(h)->trace_on = on != 0;
    return CXB_OK;
} CXB_ABI_CATCH(CXB_ERR_INTERNAL)
int64_t cxb_trace_get(cxb_engine* h, int64_t* out_level, int64_t* out_sid, int64_t cap) try {
    DeviceEngine* e = E(h);
    for (int64_t i = 0; i < (int64_t)e->tr_sid.size() && i < cap; ++i) {
        if (out_level) out_level[i] = e->tr_level[i];
        if (out_sid) out_sid[i] = e->tr_sid[i];
    }
    return (int64_t)e->tr_sid.size();
} CXB_ABI_CATCH(-1)
int64_t cxb_trace_get_times(cxb_engine* h, int64_t* out_ns, int64_t cap) try {
    DeviceEngine* e = E(h);
    for (int64_t i = 0; i < (int64_t)e->tr_ns.size() && i < cap; ++i)
        if (out_ns) out_ns[i] = e->tr_ns[i];
    return (int64_t)e->tr_ns.size();
} CXB_ABI_CATCH(-1)

}  // extern "C"
