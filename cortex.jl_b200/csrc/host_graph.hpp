// host_graph.hpp — host-side graph ingestion and dependency wiring of the product library.
//
// Build-time only: turns the model-engine backend's answers into flat arrays and
// reproduces DefaultDependencyResolver (src/dependencies.jl:5-173) as an append-only *dependency
// log* (signal, dependency, flags) from which the device CSR is built by two stable counting sorts
// (by signal -> dependency lists in add_dependency! order; by dependency -> listener lists in
// add_dependency! order, src/signal.jl:310-320).  No per-signal heap objects: config-5-sized graphs
// (1e8 signals) stay a few GB of flat vectors.
//
// Dynamic state (props, C/F nibble bits, values) lives on the DEVICE once uploaded; the host keeps
// a mirror that is only valid while `host_state_valid` (engine.cu downloads it before any
// structural mutation after upload).
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cortex_b200.h"

namespace cxb {

constexpr uint8_t P_PP = 1, P_P = 2, P_COMPUTED = 4;  // per-signal props byte
constexpr uint8_t E_LISTEN = 0x10;                    // log flag: listen = true

struct HostGraph {
    // ---- model graph (shared id space, neighbours ascending by id) -------------------------------
    int64_t n_ids = 0, n_var = 0, n_conn = 0;
    std::vector<uint8_t> is_factor;
    std::vector<int32_t> ftype;
    std::vector<int64_t> variables, factors;
    std::vector<int32_t> marg_of;  // id -> marginal sid (-1 for factors)
    std::vector<int64_t> adj_off, adj_nbr;
    std::vector<int32_t> adj_conn;  // connection index of each adjacency entry
    std::vector<int64_t> conn_var, conn_fac;
    bool built = false;

    // ---- signals -----------------------------------------------------------------------------------
    std::vector<uint8_t> kind;
    std::vector<int64_t> svar, sfac;
    std::vector<int32_t> r0, r1;
    std::vector<uint8_t> props;       // mirror
    std::vector<int32_t> lis_count, dep_count;

    // ---- dependency log ----------------------------------------------------------------------------
    std::vector<int32_t> e_sig, e_dep;
    std::vector<uint8_t> e_flags;  // CXB_NIB_INTERMEDIATE | CXB_NIB_WEAK | E_LISTEN
    std::vector<uint8_t> e_nib;    // mirror of the dynamic C / F bits

    std::vector<std::pair<int64_t, int32_t>> links;  // (variable id, signal) in link order
    std::vector<int64_t> warnings;

    int32_t n_sig() const { return (int32_t)kind.size(); }

    int32_t new_signal(uint8_t k = CXB_KIND_UNSPECIFIED, int64_t v = -1, int64_t f = -1) {
        kind.push_back(k);
        svar.push_back(v);
        sfac.push_back(f);
        r0.push_back(0);
        r1.push_back(-1);
        props.push_back(0);
        lis_count.push_back(0);
        dep_count.push_back(0);
        return n_sig() - 1;
    }
    int32_t m2v_of_conn(int32_t c) const { return (int32_t)(n_var + 2 * (int64_t)c); }
    int32_t m2f_of_conn(int32_t c) const { return (int32_t)(n_var + 2 * (int64_t)c + 1); }
    // binary search in the sorted adjacency of `a` for neighbour `b`
    int32_t conn_of(int64_t a, int64_t b) const {
        auto lo = adj_nbr.begin() + adj_off[a], hi = adj_nbr.begin() + adj_off[a + 1];
        auto it = std::lower_bound(lo, hi, b);
        if (it == hi || *it != b) return -1;
        return adj_conn[it - adj_nbr.begin()];
    }

    int32_t build(int64_t n, const uint8_t* isf, const int32_t* ft, int64_t n_edges, const int64_t* ev, const int64_t* ef,
                  std::string& err) {
        if (built || n_sig() != 0) {
            err = "graph already built (build the graph before creating free signals)";
            return CXB_ERR_STATE;
        }
        if (n < 0 || n_edges < 0 || (n > 0 && !isf) || (n_edges > 0 && (!ev || !ef))) {
            err = "graph_build: negative size or null array";
            return CXB_ERR_BAD_ARG;
        }
        if ((n + 2 * n_edges) >= (int64_t)2000000000) {
            err = "graph too large for 32-bit signal ids";
            return CXB_ERR_BAD_ARG;
        }
        n_ids = n;
        is_factor.assign(isf, isf + n);
        ftype.assign(n, 0);
        marg_of.assign(n, -1);
        for (int64_t i = 0; i < n; ++i) {
            if (isf[i]) {
                factors.push_back(i);
                ftype[i] = ft ? ft[i] : 0;
            } else {
                variables.push_back(i);
            }
        }
        n_var = (int64_t)variables.size();
        n_conn = n_edges;
        size_t total = (size_t)(n_var + 2 * n_edges);
        kind.reserve(total);
        for (int64_t v : variables) marg_of[v] = new_signal(CXB_KIND_MARGINAL, v, -1);  // set_signals_variants!
        std::vector<int64_t> deg(n + 1, 0);
        for (int64_t c = 0; c < n_edges; ++c) {
            int64_t v = ev[c], f = ef[c];
            if (v < 0 || v >= n || f < 0 || f >= n || isf[v] || !isf[f]) {
                err = "graph_build: edge endpoints must be (variable id, factor id)";
                return CXB_ERR_BAD_ARG;
            }
            new_signal(CXB_KIND_M2V, v, f);
            new_signal(CXB_KIND_M2F, v, f);
            ++deg[v];
            ++deg[f];
        }
        conn_var.assign(ev, ev + n_edges);
        conn_fac.assign(ef, ef + n_edges);
        adj_off.assign(n + 1, 0);
        for (int64_t i = 0; i < n; ++i) adj_off[i + 1] = adj_off[i] + deg[i];
        adj_nbr.resize(2 * n_edges);
        adj_conn.resize(2 * n_edges);
        std::vector<int64_t> cur(adj_off.begin(), adj_off.end() - 1);
        for (int64_t c = 0; c < n_edges; ++c) {
            adj_nbr[cur[ev[c]]] = ef[c];
            adj_conn[cur[ev[c]]++] = (int32_t)c;
            adj_nbr[cur[ef[c]]] = ev[c];
            adj_conn[cur[ef[c]]++] = (int32_t)c;
        }
        std::vector<std::pair<int64_t, int32_t>> tmp;
        for (int64_t i = 0; i < n; ++i) {  // sort each adjacency by neighbour id (skip if already sorted)
            int64_t a = adj_off[i], b = adj_off[i + 1];
            if (std::is_sorted(adj_nbr.begin() + a, adj_nbr.begin() + b)) {
                if (std::adjacent_find(adj_nbr.begin() + a, adj_nbr.begin() + b) != adj_nbr.begin() + b) {
                    err = "graph_build: duplicate edge";
                    return CXB_ERR_BAD_ARG;
                }
                continue;
            }
            tmp.clear();
            for (int64_t k = a; k < b; ++k) tmp.emplace_back(adj_nbr[k], adj_conn[k]);
            std::sort(tmp.begin(), tmp.end());
            for (int64_t k = a; k < b; ++k) {
                adj_nbr[k] = tmp[k - a].first;
                adj_conn[k] = tmp[k - a].second;
                if (k > a && adj_nbr[k] == adj_nbr[k - 1]) {
                    err = "graph_build: duplicate edge";
                    return CXB_ERR_BAD_ARG;
                }
            }
        }
        built = true;
        return CXB_OK;
    }

    // add_dependency!, src/signal.jl:286-337 — appends to the log, maintains the mirrors
    void add_dependency(int32_t s, int32_t d, bool weak, bool listen, bool check_computed, bool intermediate) {
        if (s == d) return;  // :295-297
        uint8_t fl = (weak ? CXB_NIB_WEAK : 0) | (intermediate ? CXB_NIB_INTERMEDIATE : 0) | (listen ? E_LISTEN : 0);
        uint8_t nib = 0;
        bool dc = props[d] & P_COMPUTED, sc = props[s] & P_COMPUTED;
        if (check_computed && dc) {  // :324-331
            nib |= CXB_NIB_COMPUTED;
            if (!sc) nib |= CXB_NIB_FRESH;
            props[s] = (props[s] & P_COMPUTED) | P_PP;
        } else if (check_computed && !dc) {  // :332-334
            props[s] = props[s] & P_COMPUTED;
        }
        e_sig.push_back(s);
        e_dep.push_back(d);
        e_flags.push_back(fl);
        e_nib.push_back(nib);
        ++dep_count[s];
        ++lis_count[d];
    }

    // ---- DefaultDependencyResolver, src/dependencies.jl ------------------------------------------
    void resolve_factor_default(int64_t f) {  // :17-31, nested iteration order of the neighbour list
        int64_t a = adj_off[f], b = adj_off[f + 1];
        for (int64_t i = a; i < b; ++i)
            for (int64_t j = a; j < b; ++j)
                if (i != j) add_dependency(m2v_of_conn(adj_conn[i]), m2f_of_conn(adj_conn[j]), false, true, true, false);
    }
    // :128-173; [lo,hi] inclusive 0-based positions in the variable's neighbour list starting at `base`
    int32_t segment_tree(int64_t v, int64_t base, int64_t lo, int64_t hi) {
        int64_t len = hi - lo + 1;
        if (len == 1) return m2v_of_conn(adj_conn[base + lo]);
        int64_t mid = len / 2;
        int32_t left = segment_tree(v, base, lo, lo + mid - 1);
        int32_t right = segment_tree(v, base, lo + mid, hi);
        for (int64_t k = lo; k < lo + mid; ++k) {
            int32_t mf = m2f_of_conn(adj_conn[base + k]);
            if (lis_count[mf] > 0) add_dependency(mf, right, false, true, true, true);
        }
        for (int64_t k = lo + mid; k <= hi; ++k) {
            int32_t mf = m2f_of_conn(adj_conn[base + k]);
            if (lis_count[mf] > 0) add_dependency(mf, left, false, true, true, true);
        }
        int32_t node = new_signal(CXB_KIND_PRODUCT, v, -1);
        r0[node] = (int32_t)lo;
        r1[node] = (int32_t)hi;
        add_dependency(node, left, false, true, true, true);
        add_dependency(node, right, false, true, true, true);
        return node;
    }
    void resolve_variable_default(int64_t v) {  // :33-126
        int64_t a = adj_off[v], n = adj_off[v + 1] - a;
        int32_t marg = marg_of[v];
        if (n == 0) {
            warnings.push_back(v);
            return;
        }
        if (n < 2) {
            add_dependency(marg, m2v_of_conn(adj_conn[a]), false, true, true, true);
            return;
        }
        if (n <= 5) {
            for (int64_t i = 0; i < n; ++i) {
                add_dependency(marg, m2v_of_conn(adj_conn[a + i]), false, true, true, true);
                int32_t mf = m2f_of_conn(adj_conn[a + i]);
                if (lis_count[mf] > 0)
                    for (int64_t j = 0; j < n; ++j)
                        if (j != i) add_dependency(mf, m2v_of_conn(adj_conn[a + j]), false, true, true, true);
            }
            return;
        }
        int64_t mid = n / 2;
        int32_t left = segment_tree(v, a, 0, mid - 1);
        int32_t right = segment_tree(v, a, mid, n - 1);
        for (int64_t k = 0; k < mid; ++k) {
            int32_t mf = m2f_of_conn(adj_conn[a + k]);
            if (lis_count[mf] > 0) add_dependency(mf, right, false, true, true, true);
        }
        for (int64_t k = mid; k < n; ++k) {
            int32_t mf = m2f_of_conn(adj_conn[a + k]);
            if (lis_count[mf] > 0) add_dependency(mf, left, false, true, true, true);
        }
        add_dependency(marg, left, false, true, true, true);
        add_dependency(marg, right, false, true, true, true);
    }
    // MeanFieldResolver, test/inference_engine_tests.jl:597-621
    void resolve_factor_mean_field(int64_t f) {
        int64_t a = adj_off[f], b = adj_off[f + 1];
        for (int64_t i = a; i < b; ++i)
            for (int64_t j = a; j < b; ++j)
                if (i != j) add_dependency(m2v_of_conn(adj_conn[i]), marg_of[adj_nbr[j]], true, true, true, false);
    }
    void resolve_variable_mean_field(int64_t v) {
        for (int64_t i = adj_off[v]; i < adj_off[v + 1]; ++i)
            add_dependency(marg_of[v], m2v_of_conn(adj_conn[i]), false, true, true, true);
    }
    // one resolve_factor_dependencies! / resolve_variable_dependencies! call of a built-in resolver (user resolvers delegate
    // per id, test/inference_engine_tests.jl:813-815)
    int32_t resolve_one(int32_t resolver, int64_t id, bool factor, std::string& err) {
        if (resolver != CXB_RESOLVER_DEFAULT_BP && resolver != CXB_RESOLVER_MEAN_FIELD) {
            err = "unknown resolver";
            return CXB_ERR_BAD_ARG;
        }
        if (id < 0 || id >= n_ids || (is_factor[id] != 0) != factor) {
            err = factor ? "resolve_factor_dependencies: not a factor id" : "resolve_variable_dependencies: not a variable id";
            return CXB_ERR_BAD_ARG;
        }
        const bool bp = resolver == CXB_RESOLVER_DEFAULT_BP;
        if (factor)
            bp ? resolve_factor_default(id) : resolve_factor_mean_field(id);
        else
            bp ? resolve_variable_default(id) : resolve_variable_mean_field(id);
        return CXB_OK;
    }
    int32_t resolve(int32_t resolver, std::string& err) {  // :5-15 factors first, then variables
        if (resolver == CXB_RESOLVER_NONE) return CXB_OK;
        if (resolver != CXB_RESOLVER_DEFAULT_BP && resolver != CXB_RESOLVER_MEAN_FIELD) {
            err = "unknown resolver";
            return CXB_ERR_BAD_ARG;
        }
        size_t guess = 0;
        for (int64_t f : factors) {
            size_t d = (size_t)(adj_off[f + 1] - adj_off[f]);
            guess += d * (d > 0 ? d - 1 : 0);
        }
        for (int64_t v : variables) {
            size_t d = (size_t)(adj_off[v + 1] - adj_off[v]);
            guess += d <= 5 ? d * d : 2 * d + d * 8;
        }
        e_sig.reserve(e_sig.size() + guess);
        e_dep.reserve(e_dep.size() + guess);
        e_flags.reserve(e_flags.size() + guess);
        e_nib.reserve(e_nib.size() + guess);
        bool bp = resolver == CXB_RESOLVER_DEFAULT_BP;
        for (int64_t f : factors) bp ? resolve_factor_default(f) : resolve_factor_mean_field(f);
        for (int64_t v : variables) bp ? resolve_variable_default(v) : resolve_variable_mean_field(v);
        return CXB_OK;
    }
};

// Flattened structure handed to the device (built from the log by stable counting sorts).
struct Csr {
    std::vector<uint32_t> dep_off, dep_ids;  // [N+1], [E]
    std::vector<uint32_t> nib_off;           // [N+1] in 64-bit chunks (>= 1 chunk per signal, src/signal.jl:40-44)
    std::vector<uint64_t> nib;               // 16 nibbles per chunk, src/signal.jl:522-526
    std::vector<uint32_t> lis_off, lis_ids, lis_slot;  // [N+1], [E], [E]; slot = FIRST matching slot (src/signal.jl:345-353)
    std::vector<uint8_t> lis_listen;
    std::vector<uint32_t> edge_pos;  // log entry -> position in dep_ids (to map nibbles back)
};

inline void build_csr(const HostGraph& g, Csr& c) {
    const size_t N = (size_t)g.n_sig(), E = g.e_sig.size();
    c.dep_off.assign(N + 1, 0);
    c.lis_off.assign(N + 1, 0);
    c.nib_off.assign(N + 1, 0);
    for (size_t i = 0; i < N; ++i) {
        c.dep_off[i + 1] = c.dep_off[i] + (uint32_t)g.dep_count[i];
        c.lis_off[i + 1] = c.lis_off[i] + (uint32_t)g.lis_count[i];
        uint32_t chunks = g.dep_count[i] == 0 ? 1u : (uint32_t)((g.dep_count[i] + 15) / 16);
        c.nib_off[i + 1] = c.nib_off[i] + chunks;
    }
    c.dep_ids.resize(E);
    c.edge_pos.resize(E);
    c.nib.assign(c.nib_off[N], 0);
    c.lis_ids.resize(E);
    c.lis_slot.resize(E);
    c.lis_listen.resize(E);
    std::vector<uint32_t> cur(c.dep_off.begin(), c.dep_off.end() - 1);
    for (size_t e = 0; e < E; ++e) {
        uint32_t s = (uint32_t)g.e_sig[e];
        uint32_t pos = cur[s]++;
        c.dep_ids[pos] = (uint32_t)g.e_dep[e];
        c.edge_pos[e] = pos;
        uint32_t slot = pos - c.dep_off[s];
        uint64_t nibble = (uint64_t)((g.e_flags[e] & 0x3) | (g.e_nib[e] & 0xC));
        c.nib[c.nib_off[s] + (slot >> 4)] |= nibble << ((slot & 15) << 2);
    }
    std::vector<uint32_t> lcur(c.lis_off.begin(), c.lis_off.end() - 1);
    for (size_t e = 0; e < E; ++e) {
        uint32_t d = (uint32_t)g.e_dep[e], s = (uint32_t)g.e_sig[e];
        uint32_t k = lcur[d]++;
        c.lis_ids[k] = s;
        c.lis_listen[k] = (g.e_flags[e] & E_LISTEN) ? 1 : 0;
        uint32_t slot = c.edge_pos[e] - c.dep_off[s];
        // duplicates: every listener entry of (d -> s) refreshes only the FIRST slot holding d
        for (uint32_t q = c.dep_off[s]; q < c.dep_off[s] + slot; ++q)
            if (c.dep_ids[q] == d) {
                slot = q - c.dep_off[s];
                break;
            }
        c.lis_slot[k] = slot;
    }
}

// copy the dynamic C/F bits of the packed chunks back into the log mirror
inline void unpack_nibbles(HostGraph& g, const Csr& c) {
    for (size_t e = 0; e < g.e_sig.size(); ++e) {
        uint32_t s = (uint32_t)g.e_sig[e];
        uint32_t slot = c.edge_pos[e] - c.dep_off[s];
        uint64_t nibble = (c.nib[c.nib_off[s] + (slot >> 4)] >> ((slot & 15) << 2)) & 0xF;
        g.e_nib[e] = (uint8_t)(nibble & 0xC);
    }
}

}  // namespace cxb
