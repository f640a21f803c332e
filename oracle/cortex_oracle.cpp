// cortex_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A literal restatement of the hot path of ReactiveBayes/Cortex.jl v0.3.0 (pure Julia, cannot run
// here: no `julia` in the image) with integer signal ids instead of object pointers:
//     src/signal.jl            (all of it: state machine, nibble props, process_dependencies!)
//     src/inference_engine.jl  :228-247, 294-323, 479-546, 555-632 (request, scan, update_marginals!)
//     src/dependencies.jl      :1-173 (DefaultDependencyResolver incl. the segment tree)
//     src/inference_signal.jl  :16-96 (variants)
//     ext/BipartiteFactorGraphsExt/BipartiteFactorGraphsExt.jl:26-48 (iteration-order contract)
// plus the rule arithmetic of the reference's own test fixtures (test/runtests.jl:40-46,78-88;
// test/inference_engine_tests.jl:256-294, 385-432, 1163-1179) and the rule definitions of
// SURVEY.md Appendix C for the benchmark configs.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (cortex.jl_b200/csrc) never links or calls it.
//
// Parity pinning: the reference's own golden behaviour (tests/test_oracle_*.py port
// test/signal_tests.jl, test/dependencies_tests.jl and test/inference_engine_tests.jl).
// Third-party dependency outside /root/reference: BipartiteFactorGraphs.jl 1.0.x (Project.toml:7,17)
// contributes only id allocation and neighbour order; neighbour order = ascending id here
// ("parity unpinned" for that single convention).
//
// Two schedules are provided for update_marginals!:
//   seq — the sequential in-place schedule of src/inference_engine.jl:559-632, literally;
//   lvl — the level-synchronous equivalent the device runs.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/cortex_b200.h"

namespace {

constexpr uint64_t MASK_I = 0x1, MASK_W = 0x2, MASK_C = 0x4, MASK_F = 0x8;  // src/signal.jl:507-510
constexpr uint64_t ALL_W = 0x2222222222222222ull, ALL_C = 0x4444444444444444ull,  // :512-515
                   ALL_F = 0x8888888888888888ull, PASS = 0x1111111111111111ull;   // :519

typedef int32_t (*rule_cb_t)(void* user, int64_t sid, int32_t kind, int64_t var, int64_t fac, int64_t ndeps,
                             const int64_t* dep_ids, const double* dep_values, double* out);
typedef int32_t (*visit_cb_t)(void* user, int64_t dep);

struct Sig {  // src/signal.jl:82-115
    bool computed = false;  // value !== UndefValue(), :162-164
    bool pp = false, p = false;  // SignalProps :48-51
    int32_t kind = CXB_KIND_UNSPECIFIED;
    int64_t var = -1, fac = -1, r0 = 0, r1 = -1;  // variant payload, src/inference_signal.jl:28-96
    int64_t ndeps = 0;
    std::vector<uint64_t> chunks{0};  // SignalDependenciesProps, :36-45 (>= 1 chunk)
    std::vector<int64_t> deps;
    std::vector<uint8_t> listenmask;
    std::vector<int64_t> listeners;
};

struct Rule {
    int32_t kind = CXB_RULE_NONE;
    std::vector<double> params;
};

struct TraceRec {
    int64_t round, var, sid;
    int64_t ns = 0;  // measured duration of this execution (TracedInferenceExecution.total_time_in_ns, src/inference_engine.jl:650-657)
};
static inline int64_t now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Oracle {
    int dim = 1, family = 0;
    std::vector<Sig> sig;
    std::vector<double> val;
    // graph (BipartiteFactorGraph-like: one id space, adjacency sorted by id)
    int64_t n_ids = 0;
    std::vector<uint8_t> is_factor;
    std::vector<int32_t> ftype;
    std::vector<double> fparam;
    std::vector<uint8_t> fparam_set;
    std::vector<uint8_t> var_family;  // per id, 0xFF = the engine family (cxo_set_variable_families)
    std::vector<std::vector<int64_t>> nbr;
    std::vector<int64_t> variables, factors, marg_of;  // marg_of[id] = sid or -1
    std::unordered_map<uint64_t, int64_t> conn;        // (v,f) -> connection index
    int64_t n_var = 0, n_conn = 0;
    std::vector<std::vector<int64_t>> linked;
    std::map<int32_t, Rule> rules;
    rule_cb_t cb = nullptr;
    void* cb_user = nullptr;
    std::vector<int64_t> warnings;
    std::vector<std::vector<int64_t>> lists, prepared;  // cxo_prepare_signals / cxo_prepare_request
    // request state, src/inference_engine.jl:265-270
    std::vector<int64_t> req_ids, req_marg;
    std::vector<uint8_t> ready;
    // trace of the last update
    std::vector<TraceRec> trace;
    cxb_update_stats stats{};
    int64_t n_is_pending_calls = 0;
    std::string err;

    int64_t new_signal() {
        sig.emplace_back();
        val.resize(sig.size() * (size_t)dim, 0.0);
        return (int64_t)sig.size() - 1;
    }
    double* value(int64_t s) { return &val[(size_t)s * dim]; }
    uint64_t key(int64_t v, int64_t f) const { return (uint64_t)v * (uint64_t)n_ids + (uint64_t)f; }
    int64_t m2v(int64_t v, int64_t f) const {
        auto it = conn.find(key(v, f));
        return it == conn.end() ? -1 : n_var + 2 * it->second;
    }
    int64_t m2f(int64_t v, int64_t f) const {
        auto it = conn.find(key(v, f));
        return it == conn.end() ? -1 : n_var + 2 * it->second + 1;
    }

    // ---- nibble access, src/signal.jl:522-609 -------------------------------------------------
    static bool nib(const Sig& s, int64_t i, uint64_t m) { return (s.chunks[i >> 4] >> ((i & 15) << 2)) & m; }
    static void nib_set(Sig& s, int64_t i, uint64_t m) { s.chunks[i >> 4] |= m << ((i & 15) << 2); }

    // is_meeting_pending_criteria, src/signal.jl:668-730
    static bool criteria(const Sig& s) {
        if (s.ndeps == 0) return false;  // :671-673
        size_t nch = s.chunks.size();
        for (size_t i = 0; i + 1 < nch; ++i) {
            uint64_t c = s.chunks[i];
            uint64_t W = (c & ALL_W) >> 1, C = (c & ALL_C) >> 2, F = (c & ALL_F) >> 3;
            if ((C & (W | F)) != PASS) return false;
        }
        int64_t last = s.ndeps - 1;
        int shift = (int)((last & 15) << 2) + 4;  // :708-716
        uint64_t pad = shift >= 64 ? 0ull : (~0ull << shift);
        uint64_t c = s.chunks[last >> 4] | pad;
        uint64_t W = (c & ALL_W) >> 1, C = (c & ALL_C) >> 2, F = (c & ALL_F) >> 3;
        return (C & (W | F)) == PASS;
    }
    // is_pending, src/signal.jl:141-154
    bool is_pending(int64_t id) {
        ++n_is_pending_calls;
        Sig& s = sig[id];
        if (s.p) return true;
        if (s.pp) {
            bool r = criteria(s);
            s.pp = false;
            s.p = r;
            return r;
        }
        return false;
    }
    // add_dependency!, src/signal.jl:286-337 (+ props growth :529-544)
    void add_dependency(int64_t sid, int64_t did, bool weak, bool listen, bool check_computed, bool intermediate) {
        if (sid == did) return;  // :295-297
        Sig& s = sig[sid];
        Sig& d = sig[did];
        int64_t idx = s.ndeps++;
        if ((size_t)((4 * s.ndeps - 1) / 64 + 1) > s.chunks.size()) s.chunks.push_back(0);
        if (weak) {
            nib_set(s, idx, MASK_W);
            has_weak_dep = true;
        }
        if (intermediate) nib_set(s, idx, MASK_I);
        s.deps.push_back(did);
        d.listenmask.push_back(listen ? 1 : 0);
        d.listeners.push_back(sid);
        if (check_computed && d.computed) {  // :324-331
            nib_set(s, idx, MASK_C);
            if (!s.computed) nib_set(s, idx, MASK_F);
            s.pp = true;
            s.p = false;
        } else if (check_computed && !d.computed) {  // :332-334
            s.pp = false;
            s.p = false;
        }
    }
    // set_value! + notify_listener!, src/signal.jl:232-253, 339-356
    void set_value(int64_t sid, const double* v) {
        std::memcpy(value(sid), v, sizeof(double) * dim);
        Sig& s = sig[sid];
        s.computed = true;
        for (auto& c : s.chunks) c &= ~ALL_F;  // unset_all_dependencies_fresh!, :653-655
        s.pp = false;
        s.p = false;
        for (size_t k = 0; k < s.listeners.size(); ++k) {
            Sig& L = sig[s.listeners[k]];
            if (s.listenmask[k]) {
                L.pp = true;
                L.p = false;
            }
            for (int64_t i = 0; i < L.ndeps; ++i) {  // first matching slot only, :345-353
                if (L.deps[i] == sid) {
                    nib_set(L, i, MASK_F);
                    nib_set(L, i, MASK_C);
                    break;
                }
            }
            // strict rule E: signals that received a NON-listening notification during a level (checked after the level)
            if (!s.listenmask[k] && lvl_request > 0) {
                nl_notified.push_back(s.listeners[k]);
                nl_touched.insert(s.listeners[k]);
            }
        }
    }

    // ---- rules --------------------------------------------------------------------------------
    const Rule* rule_of_factor(int64_t f) const {
        auto it = rules.find(ftype[f]);
        return it == rules.end() ? nullptr : &it->second;
    }
    double factor_param(int64_t f, const Rule& r) const {
        if (fparam_set[f]) return fparam[f];
        return r.params.empty() ? 1.0 : r.params[0];
    }
    static void normalise(double* x, int n) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += x[i];
        for (int i = 0; i < n; ++i) x[i] = x[i] / s;
    }
    // reduce(product, values) left-to-right, test/inference_engine_tests.jl:385-413
    int32_t combine(const Sig& s, double* out) {
        if (s.ndeps == 0) {
            err = "combine: signal has no dependencies";
            return CXB_ERR_NO_RULE;
        }
        std::memcpy(out, value(s.deps[0]), sizeof(double) * dim);
        const int family = family_of(s);  // `product` dispatches on the value type (test/runtests.jl:40-46, 89-99)
        if (family == CXB_FAMILY_POINT) {
            err = "combine: an observed value has no product";
            return CXB_ERR_NO_RULE;
        }
        for (int64_t i = 1; i < s.ndeps; ++i) {
            const double* b = value(s.deps[i]);
            switch (family) {
                case CXB_FAMILY_GAUSS_MP: {  // test/runtests.jl:89-95
                    double xi = out[0] * out[1] + b[0] * b[1];
                    double w = out[1] + b[1];
                    double precision = w;
                    out[0] = (1 / precision) * xi;
                    out[1] = precision;
                    break;
                }
                case CXB_FAMILY_GAMMA: {  // test/runtests.jl:97-99
                    double shape = out[0] + b[0] - 1;
                    double scale = (out[1] * b[1]) / (out[1] + b[1]);
                    out[0] = shape;
                    out[1] = scale;
                    break;
                }
                case CXB_FAMILY_GAUSS_CANON:
                case CXB_FAMILY_SUM:
                    for (int k = 0; k < dim; ++k) out[k] = out[k] + b[k];
                    break;
                case CXB_FAMILY_CATEGORICAL:
                    for (int k = 0; k < dim; ++k) out[k] = out[k] * b[k];
                    break;
                case CXB_FAMILY_GAUSS_MV: {  // test/runtests.jl:40-46
                    double xi = out[0] / out[1] + b[0] / b[1];
                    double w = 1 / out[1] + 1 / b[1];
                    double variance = 1 / w;
                    out[0] = variance * xi;
                    out[1] = variance;
                    break;
                }
                case CXB_FAMILY_BETA:  // test/inference_engine_tests.jl:273-278
                    out[0] = out[0] + b[0] - 1;
                    out[1] = out[1] + b[1] - 1;
                    break;
                default:
                    err = "combine: unknown family";
                    return CXB_ERR_NO_RULE;
            }
        }
        if (family == CXB_FAMILY_CATEGORICAL) normalise(out, dim);
        return CXB_OK;
    }
    int family_of(const Sig& s) const {
        if (s.var >= 0 && s.var < (int64_t)var_family.size() && var_family[s.var] != 0xFF) return var_family[s.var];
        return family;
    }
    // mean / var of a 2-parameter value by family (test/runtests.jl:36-38, 57-59, 68-69); POINT: the value itself, 0
    static double vmp_mean(int fam, const double* v) { return fam == CXB_FAMILY_GAMMA ? v[0] * v[1] : v[0]; }
    static double vmp_var(int fam, const double* v) {
        if (fam == CXB_FAMILY_GAMMA) return v[0] * (v[1] * v[1]);
        if (fam == CXB_FAMILY_GAUSS_MP) return 1 / v[1];
        if (fam == CXB_FAMILY_GAUSS_MV) return v[1];
        return 0.0;
    }
    bool run_program(const Sig& s, const Rule& r, double* out) {
        const std::vector<double>& p = r.params;
        if (p.empty()) return false;
        const int nc = (int)p[0];
        if (nc < 0 || (size_t)nc + 1 > p.size()) return false;
        const double* consts = p.data() + 1;
        const double* code = consts + nc;
        const int n_code = (int)p.size() - 1 - nc;
        const double param = fparam_set[s.fac] ? fparam[s.fac] : (nc > 0 ? consts[0] : 1.0);
        double st[16], tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int sp = 0, pc = 0;
        auto imm = [&](int& v) {
            if (pc >= n_code) return false;
            v = (int)code[pc++];
            return true;
        };
        while (pc < n_code) {
            const int op = (int)code[pc++];
            int a = 0, b = 0;
            switch (op) {
                case CXB_OP_DEP:
                    if (!imm(a) || !imm(b) || a < 0 || a >= s.ndeps || b < 0 || b >= dim || sp >= 16) return false;
                    st[sp++] = value(s.deps[a])[b];
                    break;
                case CXB_OP_CONST:
                    if (!imm(a) || a < 0 || a >= nc || sp >= 16) return false;
                    st[sp++] = consts[a];
                    break;
                case CXB_OP_PARAM:
                    if (sp >= 16) return false;
                    st[sp++] = param;
                    break;
                case CXB_OP_NDEPS:
                    if (sp >= 16) return false;
                    st[sp++] = (double)s.ndeps;
                    break;
                case CXB_OP_ADD: case CXB_OP_SUB: case CXB_OP_MUL: case CXB_OP_DIV: {
                    if (sp < 2) return false;
                    const double y = st[--sp], x = st[--sp];
                    st[sp++] = op == CXB_OP_ADD ? x + y : op == CXB_OP_SUB ? x - y : op == CXB_OP_MUL ? x * y : x / y;
                    break;
                }
                case CXB_OP_NEG: case CXB_OP_EXP: case CXB_OP_LOG: case CXB_OP_SQRT:
                    if (sp < 1) return false;
                    st[sp - 1] = op == CXB_OP_NEG ? -st[sp - 1] : op == CXB_OP_EXP ? std::exp(st[sp - 1]) : op == CXB_OP_LOG ? std::log(st[sp - 1]) : std::sqrt(st[sp - 1]);
                    break;
                case CXB_OP_STORE:
                    if (!imm(a) || a < 0 || a >= dim || sp < 1) return false;
                    out[a] = st[--sp];
                    break;
                case CXB_OP_TSET:
                    if (!imm(a) || a < 0 || a >= 8 || sp < 1) return false;
                    tmp[a] = st[--sp];
                    break;
                case CXB_OP_TGET:
                    if (!imm(a) || a < 0 || a >= 8 || sp >= 16) return false;
                    st[sp++] = tmp[a];
                    break;
                default:
                    return false;
            }
        }
        return true;
    }
    int32_t rule_m2v(const Sig& s, double* out) {
        const Rule* r = rule_of_factor(s.fac);
        if (!r || r->kind == CXB_RULE_NONE) {
            err = "The function `compute_message_to_variable!` is not implemented for factor type " +
                  std::to_string(ftype[s.fac]);
            return CXB_ERR_NO_RULE;
        }
        if (s.ndeps < 1) {
            err = "m2v rule: no dependencies";
            return CXB_ERR_NO_RULE;
        }
        const double* in = value(s.deps[0]);
        switch (r->kind) {
            case CXB_RULE_GAUSS_OBS: {
                double rv = factor_param(s.fac, *r);
                out[0] = 1.0 / rv;
                out[1] = in[0] / rv;
                return CXB_OK;
            }
            case CXB_RULE_GAUSS_RW: {
                double q = factor_param(s.fac, *r);
                double den = 1.0 + q * in[0];
                out[0] = in[0] / den;
                out[1] = in[1] / den;
                return CXB_OK;
            }
            case CXB_RULE_GAUSS_MV_OBS:  // test/inference_engine_tests.jl:425-426
                out[0] = in[0];
                out[1] = factor_param(s.fac, *r);
                return CXB_OK;
            case CXB_RULE_GAUSS_MV_RW:  // :427-428
                out[0] = in[0];
                out[1] = in[1] + factor_param(s.fac, *r);
                return CXB_OK;
            case CXB_RULE_BETA_BERNOULLI:  // :256-258
                out[0] = 1.0 + in[0];
                out[1] = 2.0 - in[0];
                return CXB_OK;
            case CXB_RULE_SCALE2:  // :1163-1166
                for (int k = 0; k < dim; ++k) out[k] = 2 * in[k];
                return CXB_OK;
            case CXB_RULE_NORMAL_MEAN_FIELD: {  // test/inference_engine_tests.jl:652-695
                if (s.ndeps != 2 || dim < 2) {
                    err = "NORMAL_MEAN_FIELD: expects the marginals of the two other variables (value_dim >= 2)";
                    return CXB_ERR_NO_RULE;
                }
                const int fa = family_of(sig[s.deps[0]]), fb = family_of(sig[s.deps[1]]);
                const double *a = value(s.deps[0]), *b = value(s.deps[1]);
                if ((fa == CXB_FAMILY_GAMMA) != (fb == CXB_FAMILY_GAMMA)) {  // :666-676
                    const bool a_is_w = fa == CXB_FAMILY_GAMMA;
                    out[0] = a_is_w ? vmp_mean(fb, b) : vmp_mean(fa, a);
                    out[1] = a_is_w ? vmp_mean(fa, a) : vmp_mean(fb, b);
                    return CXB_OK;
                }
                if (fa != CXB_FAMILY_GAMMA) {  // :678-692
                    double dm = vmp_mean(fa, a) - vmp_mean(fb, b);
                    out[0] = 1.5;
                    out[1] = 2 / (vmp_var(fa, a) + vmp_var(fb, b) + dm * dm);
                    return CXB_OK;
                }
                err = "NORMAL_MEAN_FIELD: two Gamma dependencies";
                return CXB_ERR_NO_RULE;
            }
            case CXB_RULE_NORMAL_STRUCTURED: {  // test/inference_engine_tests.jl:942-973, 1008-1028
                for (int k = 0; k < dim; ++k) out[k] = 0.0;
                if (s.kind == CXB_KIND_JOINT) {  // :942-973
                    if (s.ndeps != 3 || dim < 6) {
                        err = "NORMAL_STRUCTURED: a joint marginal expects (m2f, m2f, marginal) and value_dim >= 6";
                        return CXB_ERR_NO_RULE;
                    }
                    const double *m1 = value(s.deps[0]), *m2 = value(s.deps[1]), *g = value(s.deps[2]);
                    double xi_out = m1[1] * m1[0], W_out = m1[1];
                    double xi_mu = m2[1] * m2[0], W_mu = m2[1];
                    double W_bar = g[0] * g[1];
                    double W11 = W_out + W_bar, W12 = -W_bar, W21 = -W_bar, W22 = W_mu + W_bar;
                    double det = W11 * W22 - W12 * W21;  // mu = inv(W) * [xi_out; xi_mu], closed-form 2x2 inverse
                    out[0] = (W22 * xi_out - W12 * xi_mu) / det;
                    out[1] = (W11 * xi_mu - W21 * xi_out) / det;
                    out[2] = W11;
                    out[3] = W12;
                    out[4] = W21;
                    out[5] = W22;
                    return CXB_OK;
                }
                if (s.ndeps == 2) {  // :1013-1020
                    const Sig &d0 = sig[s.deps[0]], &d1 = sig[s.deps[1]];
                    bool first_is_msg = d0.kind == CXB_KIND_M2F;
                    if (first_is_msg == (d1.kind == CXB_KIND_M2F)) {
                        err = "NORMAL_STRUCTURED: expects one message and one marginal";
                        return CXB_ERR_NO_RULE;
                    }
                    const double* m = value(s.deps[first_is_msg ? 0 : 1]);
                    const double* g = value(s.deps[first_is_msg ? 1 : 0]);
                    out[0] = m[0];
                    out[1] = 1 / (1 / m[1] + 1 / (g[0] * g[1]));
                    return CXB_OK;
                }
                if (s.ndeps == 1 && sig[s.deps[0]].kind == CXB_KIND_JOINT && dim >= 6) {  // :1021-1026
                    const double* j = value(s.deps[0]);
                    double det = j[2] * j[5] - j[3] * j[4];
                    double V11 = j[5] / det, V12 = -j[3] / det, V21 = -j[4] / det, V22 = j[2] / det;
                    double dm = j[0] - j[1];
                    out[0] = 1.5;
                    out[1] = 2 / (V11 - V12 - V21 + V22 + dm * dm);
                    return CXB_OK;
                }
                err = "NORMAL_STRUCTURED: unreachable reached";
                return CXB_ERR_NO_RULE;
            }
            case CXB_RULE_PROGRAM: {  // user-defined stack program (include/cortex_b200.h, CXB_OP_*)
                for (int k = 0; k < dim; ++k) out[k] = 0.0;
                if (!run_program(s, *r, out)) {
                    err = "CXB_RULE_PROGRAM: malformed program (stack, dependency or component out of range)";
                    return CXB_ERR_NO_RULE;
                }
                return CXB_OK;
            }
            case CXB_RULE_CAT_TABLE:
            case CXB_RULE_POTTS: {
                if (s.ndeps != 1) {
                    err = "categorical table rule supports pairwise factors only";
                    return CXB_ERR_NO_RULE;
                }
                int K = dim;
                int64_t u = sig[s.deps[0]].var;
                bool u_is_low = u < s.var;  // table indexed (lower id, higher id)
                for (int a = 0; a < K; ++a) {
                    double acc = 0;
                    for (int b = 0; b < K; ++b) {
                        double psi;
                        if (r->kind == CXB_RULE_POTTS)
                            psi = (a == b) ? std::exp(r->params[0]) : 1.0;
                        else
                            psi = u_is_low ? r->params[(size_t)b * K + a] : r->params[(size_t)a * K + b];
                        acc += psi * in[b];
                    }
                    out[a] = acc;
                }
                normalise(out, K);
                return CXB_OK;
            }
            case CXB_RULE_HMM_EMIT: {
                int K = dim;
                int M = (int)r->params[0];
                int o = (int)in[0];
                if (o < 0 || o >= M) {
                    err = "HMM_EMIT: symbol out of range";
                    return CXB_ERR_BAD_ARG;
                }
                for (int a = 0; a < K; ++a) out[a] = r->params[1 + (size_t)a * M + o];
                normalise(out, K);
                return CXB_OK;
            }
        }
        err = "unknown rule kind";
        return CXB_ERR_NO_RULE;
    }
    // process! dispatch, src/inference_engine.jl:479-509
    // `free_strategy`: a bare compute!(strategy, signal) on an Unspecified signal uses the value
    // family's reduce as strategy; inside update_marginals! (process!) it is an error (:506).
    int32_t eval_rule(int64_t sid, double* out, bool free_strategy = false) {
        const Sig& s = sig[sid];
        if (cb) {
            std::vector<double> dv((size_t)s.ndeps * dim);
            for (int64_t i = 0; i < s.ndeps; ++i) std::memcpy(&dv[(size_t)i * dim], value(s.deps[i]), sizeof(double) * dim);
            int32_t st = cb(cb_user, sid, s.kind, s.var, s.fac, s.ndeps, s.deps.data(), dv.data(), out);
            if (st != CXB_OK) err = "rule callback failed";
            return st;
        }
        switch (s.kind) {
            case CXB_KIND_M2V:
                return rule_m2v(s, out);
            case CXB_KIND_M2F:
            case CXB_KIND_MARGINAL:
            case CXB_KIND_PRODUCT:
                return combine(s, out);
            case CXB_KIND_JOINT: {  // compute_joint_marginal!: the rule registered for the factor computes it
                const Rule* r = s.fac >= 0 && s.fac < n_ids && is_factor[s.fac] ? rule_of_factor(s.fac) : nullptr;
                if (r && r->kind == CXB_RULE_NORMAL_STRUCTURED) return rule_m2v(s, out);
                err = "The function `compute_joint_marginal!` is not implemented";
                return CXB_ERR_NO_RULE;
            }
            default:
                if (free_strategy) return combine(s, out);
                err = "Unprocessed signal variant";  // :506
                return CXB_ERR_NO_RULE;
        }
    }
    // compute!, src/signal.jl:392-410
    int32_t compute(int64_t sid, bool force, bool skip_if_no_listeners, bool free_strategy = false) {
        if (skip_if_no_listeners && sig[sid].listeners.empty()) return CXB_OK;
        if (!force && !is_pending(sid)) {
            err = "Signal is not pending. Cannot compute a non-pending signal. Use `force=true` to force computation.";
            return CXB_ERR_NOT_PENDING;
        }
        std::vector<double> out(dim);
        int32_t st = eval_rule(sid, out.data(), free_strategy);
        if (st != CXB_OK) return st;
        set_value(sid, out.data());
        ++stats.updates;
        ++stats.updates_by_kind[sig[sid].kind];
        return CXB_OK;
    }

    // process_dependencies!, src/signal.jl:466-490 (generic callback version)
    template <class F>
    bool process_dependencies(int64_t sid, bool retry, F&& f) {
        bool any = false;
        for (int64_t i = 0; i < sig[sid].ndeps; ++i) {
            int64_t d = sig[sid].deps[i];
            bool processed = f(d);
            if (!processed) {
                if (nib(sig[sid], i, MASK_I)) {
                    bool ip = process_dependencies(d, retry, f);
                    if (ip && retry) processed = f(d);
                    any = any || ip;
                }
            }
            any = any || processed;
        }
        return any;
    }

    // ---- wiring: src/dependencies.jl ------------------------------------------------------------
    void resolve_factor_default(int64_t f) {  // :17-31
        const auto& vs = nbr[f];
        for (int64_t v1 : vs)
            for (int64_t v2 : vs)
                if (v1 != v2) add_dependency(m2v(v1, f), m2f(v2, f), false, true, true, false);
    }
    int64_t segment_tree(int64_t v, int64_t lo, int64_t hi, const std::vector<int64_t>& fs) {  // :128-173 (0-based, inclusive)
        int64_t len = hi - lo + 1;
        if (len == 1) return m2v(v, fs[lo]);
        int64_t mid = len / 2;
        int64_t l0 = lo, l1 = lo + mid - 1, r0 = lo + mid, r1 = hi;
        int64_t left = segment_tree(v, l0, l1, fs);
        int64_t right = segment_tree(v, r0, r1, fs);
        for (int64_t k = l0; k <= l1; ++k) {
            int64_t mf = m2f(v, fs[k]);
            if (!sig[mf].listeners.empty()) add_dependency(mf, right, false, true, true, true);
        }
        for (int64_t k = r0; k <= r1; ++k) {
            int64_t mf = m2f(v, fs[k]);
            if (!sig[mf].listeners.empty()) add_dependency(mf, left, false, true, true, true);
        }
        int64_t node = new_signal();
        sig[node].kind = CXB_KIND_PRODUCT;
        sig[node].var = v;
        sig[node].r0 = lo;
        sig[node].r1 = hi;
        add_dependency(node, left, false, true, true, true);
        add_dependency(node, right, false, true, true, true);
        return node;
    }
    void resolve_variable_default(int64_t v) {  // :33-126
        const std::vector<int64_t> fs = nbr[v];
        int64_t marg = marg_of[v];
        int64_t n = (int64_t)fs.size();
        if (n == 0) {
            warnings.push_back(v);  // :40-43
            return;
        }
        if (n < 2) {
            add_dependency(marg, m2v(v, fs[0]), false, true, true, true);  // :48-55
            return;
        }
        if (n <= 5) {  // :60-88
            for (int64_t f : fs) {
                add_dependency(marg, m2v(v, f), false, true, true, true);
                int64_t mf = m2f(v, f);
                if (!sig[mf].listeners.empty())
                    for (int64_t g : fs)
                        if (g != f) add_dependency(mf, m2v(v, g), false, true, true, true);
            }
            return;
        }
        int64_t mid = n / 2;  // :90-123
        int64_t left = segment_tree(v, 0, mid - 1, fs);
        int64_t right = segment_tree(v, mid, n - 1, fs);
        for (int64_t k = 0; k < mid; ++k) {
            int64_t mf = m2f(v, fs[k]);
            if (!sig[mf].listeners.empty()) add_dependency(mf, right, false, true, true, true);
        }
        for (int64_t k = mid; k < n; ++k) {
            int64_t mf = m2f(v, fs[k]);
            if (!sig[mf].listeners.empty()) add_dependency(mf, left, false, true, true, true);
        }
        add_dependency(marg, left, false, true, true, true);
        add_dependency(marg, right, false, true, true, true);
    }
    // MeanFieldResolver, test/inference_engine_tests.jl:597-621
    void resolve_factor_mean_field(int64_t f) {
        const auto& vs = nbr[f];
        for (int64_t v1 : vs)
            for (int64_t v2 : vs)
                if (v1 != v2) add_dependency(m2v(v1, f), marg_of[v2], true, true, true, false);
    }
    void resolve_variable_mean_field(int64_t v) {
        for (int64_t f : nbr[v]) add_dependency(marg_of[v], m2v(v, f), false, true, true, true);
    }
    int32_t resolve_one(int32_t resolver, int64_t id, bool factor) {
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
    int32_t resolve(int32_t resolver) {  // src/dependencies.jl:5-15 — factors first, then variables
        if (resolver == CXB_RESOLVER_NONE) return CXB_OK;
        if (resolver != CXB_RESOLVER_DEFAULT_BP && resolver != CXB_RESOLVER_MEAN_FIELD) {
            err = "unknown resolver";
            return CXB_ERR_BAD_ARG;
        }
        for (int64_t f : factors)
            resolver == CXB_RESOLVER_DEFAULT_BP ? resolve_factor_default(f) : resolve_factor_mean_field(f);
        for (int64_t v : variables)
            resolver == CXB_RESOLVER_DEFAULT_BP ? resolve_variable_default(v) : resolve_variable_mean_field(v);
        return CXB_OK;
    }

    // ---- requests: src/inference_engine.jl:298-323 ---------------------------------------------
    int32_t request(int64_t n, const int64_t* ids) {
        req_ids.assign(ids, ids + n);
        req_marg.resize(n);
        for (int64_t i = 0; i < n; ++i) {
            int64_t v = ids[i];
            if (v < 0 || v >= n_ids || is_factor[v] || marg_of[v] < 0) {
                err = "request_inference_for: not a variable id";
                return CXB_ERR_BAD_ARG;
            }
            int64_t m = marg_of[v];
            for (int64_t d : sig[m].deps) {
                sig[d].pp = true;
                sig[d].p = false;
            }
            for (int64_t l : linked[v]) {
                sig[l].pp = true;
                sig[l].p = false;
            }
            req_marg[i] = m;
        }
        ready.assign(n, 0);
        return CXB_OK;
    }
    // scan_inference_request, :540-546 (literal: DFS order, duplicates possible)
    void scan(std::vector<int64_t>& out) {
        for (size_t i = 0; i < req_ids.size(); ++i)
            process_dependencies(req_marg[i], true, [&](int64_t d) {
                if (is_pending(d)) {
                    out.push_back(d);
                    return true;
                }
                return false;
            });
    }

    void reset_stats() {
        stats = cxb_update_stats{};
        trace.clear();
    }

    // On wirings with cycles of listening dependencies the reference's `while should_continue` loop (:575-608) need not end:
    // every round recomputes signals that make each other pending again. The reference would hang; the oracle and the device's
    // sequential executor stop after this many executions and report CXB_ERR_STATE (a terminating request recomputes a signal
    // a handful of times at most).
    int64_t seq_execution_cap() const { return 8 * (int64_t)sig.size() + 4096; }
    // update_marginals!, :559-632 — sequential, in place (the reference schedule)
    int32_t update_seq(int64_t n, const int64_t* ids) {
        reset_stats();
        int32_t st = request(n, ids);
        if (st) return st;
        int32_t fail = CXB_OK;
        bool should_continue = true, is_reverse = false;
        int64_t round = 0;
        while (should_continue) {
            bool cont = false;
            bool round_had_exec = false;
            for (int64_t k = 0; k < n; ++k) {
                int64_t i = is_reverse ? n - 1 - k : k;
                if (ready[i]) continue;
                int64_t var = req_ids[i];
                bool processed = process_dependencies(req_marg[i], true, [&](int64_t d) {  // :512-525
                    if (fail) return false;
                    if (is_pending(d)) {
                        if (stats.updates > seq_execution_cap()) {  // the reference loop itself would never return (see below)
                            err = "update_marginals!: the sequential loop does not terminate on this request (signals keep refreshing "
                                  "each other round after round)";
                            fail = CXB_ERR_STATE;
                            return false;
                        }
                        const int64_t t0 = now_ns();
                        int32_t s2 = compute(d, false, false);
                        if (s2) {
                            fail = s2;
                            return false;
                        }
                        trace.push_back({round, var, d, now_ns() - t0});
                        round_had_exec = true;
                        return true;
                    }
                    return false;
                });
                if (fail) return fail;
                if (is_pending(req_marg[i])) ready[i] = 1;  // :593-595
                cont = cont || processed;
            }
            if (round_had_exec) ++stats.levels;
            is_reverse = !is_reverse;
            should_continue = cont;
            ++round;
        }
        for (int64_t i = 0; i < n; ++i) {  // final phase, :610-628
            int64_t m = req_marg[i];
            if (is_pending(m)) {
                const int64_t t0 = now_ns();
                int32_t s2 = compute(m, false, false);
                if (s2) return s2;
                trace.push_back({-1, req_ids[i], m, now_ns() - t0});
                ++stats.final_marginals;
            }
            for (int64_t l : linked[req_ids[i]]) {
                if (!is_pending(l)) continue;
                const int64_t t0 = now_ns();
                int32_t s2 = compute(l, false, false);
                if (s2) return s2;
                trace.push_back({-1, req_ids[i], l, now_ns() - t0});
                ++stats.final_linked;
            }
        }
        return CXB_OK;
    }

    // snapshot-evaluate then apply a set of signals (one level)
    int32_t run_level(const std::vector<int64_t>& F, int64_t level_tag) {
        // independence: no member is a dependency of another member
        std::vector<int64_t> sorted(F);
        std::sort(sorted.begin(), sorted.end());
        for (int64_t s : F)
            for (int64_t d : sig[s].deps)
                if (std::binary_search(sorted.begin(), sorted.end(), d)) {
                    err = "level-synchronous schedule out of contract: frontier member depends on another member";
                    return CXB_ERR_OUT_OF_CONTRACT;
                }
        std::vector<double> tmp(F.size() * (size_t)dim);
        const int64_t t_level0 = now_ns();
        for (size_t k = 0; k < F.size(); ++k) {
            int32_t st = eval_rule(F[k], &tmp[k * dim]);
            if (st) return st;
        }
        for (size_t k = 0; k < F.size(); ++k) {
            set_value(F[k], &tmp[k * dim]);
            ++stats.updates;
            ++stats.updates_by_kind[sig[F[k]].kind];
            trace.push_back({level_tag, sig[F[k]].var, F[k]});
        }
        if (!F.empty()) {  // a level is one batch: its measured time is shared evenly by its members
            const int64_t each = std::max<int64_t>((now_ns() - t_level0) / (int64_t)F.size(), 1);
            for (size_t k = trace.size() - F.size(); k < trace.size(); ++k) trace[k].ns = each;
        }
        return CXB_OK;
    }

    // Opt-in (CXO_STRICT_FRESHNESS=1) extension of the request-time check below to EVERY signal the first traversal of a
    // request visits: not pending, yet FRESH on a strong, computed (non-input) dependency.
    // Strict refusal rules A / B / D / E of the level schedule: on by default on the oracle and on the
    // device alike; CXO_STRICT_FRESHNESS=0 restores the round-1 rules (kept for tests/fuzz_strict_cost.py).
    bool strict_freshness = !(std::getenv("CXO_STRICT_FRESHNESS") && std::atoi(std::getenv("CXO_STRICT_FRESHNESS")) == 0);
    std::unordered_set<int64_t> linked_requested;  // linked signals of the requested variables (rule G)
    bool strong_beneath_done = false, waits_on_leaf = false;
    // EXPERIMENT, off by default and oracle-only (tests/fuzz_bp_graphs.py documents what it buys): rule G, CXO_RULE_G=2. The device
    // does not need it: AUTO certifies every recorded level schedule against the sequential executor.
    int rule_g = std::getenv("CXO_RULE_G") ? std::atoi(std::getenv("CXO_RULE_G")) : 0;
    int schedule = CXB_SCHEDULE_AUTO;  // cxo_set_schedule: AUTO and SEQUENTIAL = the reference loop, LEVEL = update_lvl
    bool has_weak_dep = false;    // any weak dependency in the graph (the device allocates its probe marks only then)
    int64_t lvl_request = 0;      // > 0 while a strict level-schedule request runs (its serial number)
    int64_t lvl_serial = 0;
    std::vector<int64_t> nl_notified;     // ... during the current level
    std::unordered_set<int64_t> nl_touched;  // ... during the current request
    std::vector<uint32_t> lvl_visits;       // per level: visits that found the signal pending
    std::vector<uint8_t> revisited_via_I;  // per level: found pending again through an intermediate slot
    bool holds_leftover_freshness(const Sig& m) const {
        for (int64_t k = 0; k < m.ndeps; ++k)
            if (nib(m, k, MASK_F) && !nib(m, k, MASK_W) && sig[m.deps[k]].ndeps > 0) return true;
        return false;
    }
    // level-synchronous schedule, SURVEY Appendix A.5
    int32_t update_lvl(int64_t n, const int64_t* ids) {
        reset_stats();
        int32_t st = request(n, ids);
        if (st) return st;
        struct Scope {  // strict rule E is armed for the duration of this request only
            Oracle* o;
            ~Scope() { o->lvl_request = 0; }
        } scope{this};
        lvl_request = strict_freshness ? ++lvl_serial : 0;
        nl_notified.clear();
        nl_touched.clear();
        // Contract check at request time: a requested marginal that is NOT pending must not hold a FRESH bit on a computed
        // (non-input) dependency. Such leftover freshness comes from an earlier request that could not complete (missing
        // evidence); with it the reference finds the marginal pending as soon as its remaining dependencies arrive and uses
        // the stale message, which depends on the order in which the variables are visited (Gauss-Seidel) - the level
        // schedule advances all variables at once and would answer differently. Refused before anything is computed.
        std::vector<uint8_t> pending_at_request(n, 0);
        for (int64_t i = 0; i < n; ++i) {
            const Sig& m = sig[req_marg[i]];
            if (m.p || (m.pp && criteria(m))) {
                pending_at_request[i] = 1;
                continue;
            }
            for (int64_t k = 0; k < m.ndeps; ++k)
                if (nib(m, k, MASK_F) && sig[m.deps[k]].ndeps > 0) {
                    err = "level-synchronous schedule out of contract: a requested marginal holds leftover freshness from an earlier, "
                          "incomplete request (order-dependent in the reference)";
                    return CXB_ERR_OUT_OF_CONTRACT;
                }
        }
        std::vector<uint8_t> done(sig.size(), 0), inF(sig.size(), 0);
        linked_requested.clear();
        strong_beneath_done = false;
        for (int64_t i = 0; i < n; ++i)
            for (int64_t l : linked[req_ids[i]]) linked_requested.insert(l);
        int64_t level = 0;
        bool stale_beneath = false, found_pending = false, work_beneath_pending = false;
        for (;;) {
            std::vector<int64_t> F;
            auto visit = [&](int64_t d) {
                if (done[d]) return false;
                if (is_pending(d)) {
                    found_pending = true;
                    ++lvl_visits[d];
                    if (!inF[d]) {
                        inF[d] = 1;
                        F.push_back(d);
                        if (level > 0 && sig[d].kind != CXB_KIND_PRODUCT) waits_on_leaf = true;  // rule G: a late message, not a cascade node
                    }
                    return true;
                }
                if (strict_freshness && level == 0 && holds_leftover_freshness(sig[d])) stale_beneath = true;
                return false;
            };
            // DFS identical to process_dependencies! except: never descend through `done`. The reference does descend
            // (a later round reaches the signal again), computes whatever is pending underneath and recomputes the done
            // signal on the retry. With strong dependencies nothing can be pending there (the done signal consumed
            // fresh dependencies; a dependency recomputed later is caught by the listener check below); a WEAK dependency
            // never blocks, so it can: `probe` follows the same descent rule through done signals without computing
            // anything, and a pending weak dependency found there makes the request order-dependent -> refused.
            bool weak_beneath_done = false;
            strong_beneath_done = waits_on_leaf = false;
            lvl_visits.assign(sig.size(), 0);
            revisited_via_I.assign(sig.size(), 0);
            std::vector<uint8_t> probed(sig.size(), 0);
            struct Rec {
                Oracle* o;
                std::vector<uint8_t>& done;
                std::vector<uint8_t>& probed;
                bool& weak_beneath_done;
                decltype(visit)& f;
                void probe(int64_t sid) {
                    if (probed[sid]) return;
                    probed[sid] = 1;
                    for (int64_t i = 0; i < o->sig[sid].ndeps; ++i) {
                        int64_t d = o->sig[sid].deps[i];
                        if (!done[d] && o->is_pending(d)) {
                            if (nib(o->sig[sid], i, MASK_W)) weak_beneath_done = true;
                            else if (o->rule_g >= 2) o->strong_beneath_done = true;
                        } else if (nib(o->sig[sid], i, MASK_I)) {
                            probe(d);
                        }
                    }
                }
                bool go(int64_t sid) {
                    bool any = false;
                    for (int64_t i = 0; i < o->sig[sid].ndeps; ++i) {
                        int64_t d = o->sig[sid].deps[i];
                        bool processed = f(d);
                        if (processed && nib(o->sig[sid], i, MASK_I)) o->revisited_via_I[d] = 1;  // found pending through an I slot
                        if (!processed && nib(o->sig[sid], i, MASK_I)) {
                            if (done[d]) {
                                if (o->has_weak_dep || o->rule_g >= 2) probe(d);  // like the device: graphs without weak dependencies are not probed
                            } else {
                                // rule G: the variable WAITS for a message of another variable's making (a dependency that is
                                // neither pending nor computed in this request, whose slot is not satisfied, and that is not a
                                // ProductOfMessages node of an ordinary cascade)
                                if (o->sig[d].kind != CXB_KIND_PRODUCT && o->sig[d].ndeps > 0 &&
                                    !(nib(o->sig[sid], i, MASK_C) && (nib(o->sig[sid], i, MASK_W) || nib(o->sig[sid], i, MASK_F))))
                                    o->waits_on_leaf = true;
                                bool ip = go(d);
                                if (ip) processed = f(d);
                                any = any || ip;
                            }
                        }
                        any = any || processed;
                    }
                    return any;
                }
            } rec{this, done, probed, weak_beneath_done, visit};
            for (int64_t i = 0; i < n; ++i)
                if (!ready[i]) {
                    found_pending = false;
                    rec.go(req_marg[i]);
                    if (strict_freshness && level == 0 && pending_at_request[i] && found_pending) work_beneath_pending = true;
                }
            if (work_beneath_pending) {
                err = "level-synchronous schedule out of contract: a requested marginal was already pending when the request arrived and "
                      "there is pending work beneath it (the reference gives it exactly one traversal: order-dependent)";
                return CXB_ERR_OUT_OF_CONTRACT;
            }
            if (stale_beneath) {
                err = "level-synchronous schedule out of contract: a signal reached by the request is not pending but holds leftover "
                      "freshness from an earlier, incomplete request (order-dependent in the reference)";
                return CXB_ERR_OUT_OF_CONTRACT;
            }
            // Rule G. The reference re-traverses a variable that is not ready yet in every later round - THROUGH the signals it
            // computed earlier in the request: a dependency that has become pending beneath such a signal is computed there and
            // the signal recomputed (Gauss-Seidel), which a level schedule never does. Beneath a done signal a pending dependency
            // is harmless only when no later round comes by, i.e. when every variable completes by its own cascade of
            // ProductOfMessages nodes (protocol B on graphs with segment trees: the linked m2f are pending beneath the done m2v
            // and wait for the final phase). So: a pending strong dependency beneath a done signal is refused when, in the same
            // traversal, some variable depends on a message of another variable's making - it WAITS for one, or one arrives late
            // (a member of a level after the first that is not a ProductOfMessages node) - (`waits_on_leaf`), or when the
            // traversal is the last one (nothing left to compute, yet a variable is not ready).
            if (strong_beneath_done && (waits_on_leaf || F.empty())) {
                err = "level-synchronous schedule out of contract: a pending dependency lies beneath a signal already computed in this "
                      "request (the reference recomputes that signal when a later round reaches it: order-dependent)";
                return CXB_ERR_OUT_OF_CONTRACT;
            }
            if (weak_beneath_done) {
                err = "level-synchronous schedule out of contract: a pending weak dependency lies beneath a signal already computed "
                      "in this request (the reference would recompute that signal: order-dependent)";
                return CXB_ERR_OUT_OF_CONTRACT;
            }
            if (F.empty()) break;
            // contract: inside the loop phase no signal may be computed AFTER one of its listeners was
            // (then the sequential reference order is Gauss-Seidel and values are order-dependent)
            for (int64_t s : F)
                for (int64_t l : sig[s].listeners)
                    if (done[l]) {
                        err = "level-synchronous schedule out of contract: a dependency is recomputed after its listener "
                              "within one request (order-dependent in the reference): signal " +
                              std::to_string(s) + " after its listener " + std::to_string(l);
                        return CXB_ERR_OUT_OF_CONTRACT;
                    }
            // A frontier member reached more than once, at least once through an intermediate slot: the reference computed it at
            // the first visit, finds it not pending at a later one and descends through it if that slot is intermediate,
            // computing whatever is pending beneath it AFTER its listener. (Stated without the visiting order - "more than
            // once, and some visit through an intermediate slot" - so that a breadth-first device traversal can evaluate it.)
            if (strict_freshness)
                for (int64_t s : F)
                    for (int64_t d : sig[s].deps)
                        if (revisited_via_I[s] && lvl_visits[s] > 1 && (sig[d].p || (sig[d].pp && criteria(sig[d])))) {
                            err = "level-synchronous schedule out of contract: a pending signal has a pending dependency "
                                  "(order-dependent in the reference): signal " + std::to_string(s) + ", dependency " + std::to_string(d);
                            return CXB_ERR_OUT_OF_CONTRACT;
                        }
            // strict rule E, second half: a signal that received a non-listening notification EARLIER in this request becomes a
            // frontier member - in the reference the notifications arrive in DFS order, not level by level, and a non-listening
            // one that arrives last leaves the signal not pending
            for (int64_t s : F)
                if (nl_touched.count(s)) {
                    err = "level-synchronous schedule out of contract: signal " + std::to_string(s) + " became pending after a "
                          "non-listening notification of this request (order-dependent in the reference)";
                    return CXB_ERR_OUT_OF_CONTRACT;
                }
            st = run_level(F, level);
            if (st) return st;
            // strict rule E: a NON-listening notification of this level left a signal with complete criteria. Whether the
            // reference finds that signal pending depends on whether a listening notification armed its lazy flag and an
            // is_pending call consumed it in between, i.e. on the order inside what is one level here.
            for (int64_t l : nl_notified)
                if (!done[l] && !inF[l] && criteria(sig[l])) {
                    err = "level-synchronous schedule out of contract: a non-listening dependency completed the pending criteria of "
                          "signal " + std::to_string(l) + " (in the reference its pending state depends on the order of this level)";
                    return CXB_ERR_OUT_OF_CONTRACT;
                }
            nl_notified.clear();
            for (int64_t s : F) {
                done[s] = 1;
                inF[s] = 0;
            }
            for (int64_t i = 0; i < n; ++i)
                if (!ready[i] && is_pending(req_marg[i])) ready[i] = 1;
            ++stats.levels;
            ++level;
        }
        // Final-phase contract (rule F). The reference interleaves per variable: marginal(v_i), then the linked signals of
        // v_i, then v_{i+1} (src/inference_engine.jl:610-628); the level schedule computes every pending marginal, then every
        // pending linked signal. The two differ only when a final-phase candidate depends on another one whose turn comes on
        // the other side of it, so those wirings are refused (statically, from the request and the dependency lists; the
        // pending tests are evaluated without caching):
        //   F1 a linked signal of v_i depends on the marginal of a LATER requested variable that is pending now;
        //   F2 a pending requested marginal depends on a linked signal of an EARLIER requested variable;
        //   F4 a requested marginal depends on another requested marginal that is pending now;
        //   F3 (below, before the linked level) a linked signal depends on a linked signal that is pending then.
        auto pending_now = [&](int64_t s) { return sig[s].p || (sig[s].pp && criteria(sig[s])); };
        {
            std::unordered_map<int64_t, int64_t> last_req;   // marginal sid -> last request position
            std::unordered_map<int64_t, int64_t> first_link;  // linked sid -> first request position that links it
            for (int64_t i = 0; i < n; ++i) last_req[req_marg[i]] = i;
            for (int64_t i = 0; i < n; ++i)
                for (int64_t l : linked[req_ids[i]])
                    if (!first_link.count(l)) first_link[l] = i;
            bool hazard = false;
            for (int64_t i = 0; i < n && !hazard; ++i) {
                const int64_t m = req_marg[i];
                for (int64_t d : sig[m].deps) {
                    auto r = last_req.find(d);
                    if (r != last_req.end() && d != m && pending_now(d)) hazard = true;                      // F4
                    auto l = first_link.find(d);
                    if (l != first_link.end() && l->second < i && pending_now(m)) hazard = true;             // F2
                }
                for (int64_t l : linked[req_ids[i]])
                    for (int64_t d : sig[l].deps) {
                        auto r = last_req.find(d);
                        if (r != last_req.end() && r->second > i && pending_now(d)) hazard = true;           // F1
                    }
            }
            if (hazard) {
                err = "level-synchronous schedule out of contract: a final-phase signal depends on another final-phase signal across "
                      "the reference's per-variable order (marginal, then linked signals, variable by variable)";
                return CXB_ERR_OUT_OF_CONTRACT;
            }
        }
        // final phase: pending marginals, then pending linked signals (both snapshot-style)
        std::vector<int64_t> M;
        for (int64_t i = 0; i < n; ++i)
            if (is_pending(req_marg[i]) && !inF[req_marg[i]]) {
                inF[req_marg[i]] = 1;
                M.push_back(req_marg[i]);
            }
        st = run_level(M, -1);
        if (st) return st;
        for (int64_t s : M) inF[s] = 0;
        stats.final_marginals = (int64_t)M.size();
        {  // F3
            std::unordered_set<int64_t> is_link;
            for (int64_t i = 0; i < n; ++i)
                for (int64_t l : linked[req_ids[i]]) is_link.insert(l);
            for (int64_t l : is_link)
                for (int64_t d : sig[l].deps)
                    if (d != l && is_link.count(d) && pending_now(d)) {
                        err = "level-synchronous schedule out of contract: a linked signal depends on another linked signal that is "
                              "pending in the final phase (the reference computes them one after the other)";
                        return CXB_ERR_OUT_OF_CONTRACT;
                    }
        }
        std::vector<int64_t> L;
        for (int64_t i = 0; i < n; ++i)
            for (int64_t l : linked[req_ids[i]])
                if (is_pending(l) && !inF[l]) {
                    inF[l] = 1;
                    L.push_back(l);
                }
        st = run_level(L, -2);
        if (st) return st;
        stats.final_linked = (int64_t)L.size();
        return CXB_OK;
    }
};

inline Oracle* O(void* h) { return reinterpret_cast<Oracle*>(h); }

}  // namespace

extern "C" {

int32_t cxo_create(int32_t /*device*/, int32_t /*dtype*/, int32_t value_dim, int32_t family, void** out) {
    if (value_dim < 1 || !out) return CXB_ERR_BAD_ARG;
    Oracle* o = new Oracle();
    o->dim = value_dim;
    o->family = family;
    *out = o;
    return CXB_OK;
}
void cxo_destroy(void* h) { delete O(h); }
const char* cxo_last_error(void* h) { return O(h)->err.c_str(); }

int32_t cxo_graph_build(void* h, int64_t n_ids, const uint8_t* is_factor, const int32_t* factor_type, int64_t n_edges,
                        const int64_t* edge_var, const int64_t* edge_fac) {
    Oracle* o = O(h);
    if (o->n_ids != 0 || !o->sig.empty()) {
        o->err = "graph already built (build the graph before creating free signals)";
        return CXB_ERR_STATE;
    }
    if (n_ids < 0 || n_edges < 0 || (n_ids > 0 && !is_factor) || (n_edges > 0 && (!edge_var || !edge_fac))) {
        o->err = "graph_build: negative size or null array";
        return CXB_ERR_BAD_ARG;
    }
    o->n_ids = n_ids;
    o->is_factor.assign(is_factor, is_factor + n_ids);
    o->ftype.assign(n_ids, 0);
    o->fparam.assign(n_ids, 0.0);
    o->fparam_set.assign(n_ids, 0);
    o->nbr.assign(n_ids, {});
    o->marg_of.assign(n_ids, -1);
    o->linked.assign(n_ids, {});
    for (int64_t i = 0; i < n_ids; ++i) {
        if (is_factor[i]) {
            o->factors.push_back(i);
            o->ftype[i] = factor_type ? factor_type[i] : 0;
        } else {
            o->variables.push_back(i);
        }
    }
    o->n_var = (int64_t)o->variables.size();
    for (int64_t v : o->variables) {
        int64_t s = o->new_signal();
        o->marg_of[v] = s;
        o->sig[s].kind = CXB_KIND_MARGINAL;  // set_signals_variants!, src/inference_engine.jl:228-247
        o->sig[s].var = v;
    }
    o->n_conn = n_edges;
    for (int64_t c = 0; c < n_edges; ++c) {
        int64_t v = edge_var[c], f = edge_fac[c];
        if (v < 0 || v >= n_ids || f < 0 || f >= n_ids || is_factor[v] || !is_factor[f] || o->conn.count(o->key(v, f))) {
            o->err = "graph_build: bad or duplicate edge";
            return CXB_ERR_BAD_ARG;
        }
        o->conn[o->key(v, f)] = c;
        int64_t a = o->new_signal(), b = o->new_signal();
        o->sig[a].kind = CXB_KIND_M2V;
        o->sig[a].var = v;
        o->sig[a].fac = f;
        o->sig[b].kind = CXB_KIND_M2F;
        o->sig[b].var = v;
        o->sig[b].fac = f;
        o->nbr[v].push_back(f);
        o->nbr[f].push_back(v);
    }
    for (auto& a : o->nbr) std::sort(a.begin(), a.end());
    return CXB_OK;
}

int32_t cxo_register_rule(void* h, int32_t factor_type, int32_t rule_kind, const double* params, int64_t n_params) {
    Rule r;
    r.kind = rule_kind;
    if (params && n_params > 0) r.params.assign(params, params + n_params);
    O(h)->rules[factor_type] = r;
    return CXB_OK;
}
int32_t cxo_set_variable_families(void* h, int64_t n, const int64_t* variable_ids, const int32_t* families) {
    Oracle* o = O(h);
    for (int64_t i = 0; i < n; ++i) {
        int64_t v = variable_ids[i];
        if (v < 0 || v >= o->n_ids || o->is_factor[v]) {
            o->err = "set_variable_families: not a variable id";
            return CXB_ERR_BAD_ARG;
        }
        if (families[i] < 0 || families[i] > CXB_FAMILY_POINT || families[i] == CXB_FAMILY_CATEGORICAL) {
            o->err = "set_variable_families: family must be one of the fixed-size (non-categorical) families";
            return CXB_ERR_BAD_ARG;
        }
        if ((int64_t)o->var_family.size() < o->n_ids) o->var_family.resize((size_t)o->n_ids, 0xFF);
        o->var_family[v] = (uint8_t)families[i];
    }
    return CXB_OK;
}
int32_t cxo_set_factor_params(void* h, int64_t n, const int64_t* factor_ids, const double* values) {
    Oracle* o = O(h);
    for (int64_t i = 0; i < n; ++i) {
        int64_t f = factor_ids[i];
        if (f < 0 || f >= o->n_ids || !o->is_factor[f]) {
            o->err = "set_factor_params: not a factor id";
            return CXB_ERR_BAD_ARG;
        }
        o->fparam[f] = values[i];
        o->fparam_set[f] = 1;
    }
    return CXB_OK;
}
int32_t cxo_set_rule_callback(void* h, rule_cb_t cb, void* user) {
    O(h)->cb = cb;
    O(h)->cb_user = user;
    return CXB_OK;
}
int64_t cxo_create_signal(void* h) { return O(h)->new_signal(); }
int32_t cxo_add_dependency(void* h, int64_t s, int64_t d, int32_t flags) {
    Oracle* o = O(h);
    int64_t n = (int64_t)o->sig.size();
    if (s < 0 || s >= n || d < 0 || d >= n) {
        o->err = "add_dependency: bad signal id";
        return CXB_ERR_BAD_ARG;
    }
    o->add_dependency(s, d, flags & CXB_DEP_WEAK, !(flags & CXB_DEP_NO_LISTEN), !(flags & CXB_DEP_NO_CHECK_COMPUTED),
                      flags & CXB_DEP_INTERMEDIATE);
    return CXB_OK;
}
int32_t cxo_resolve_dependencies(void* h, int32_t resolver) { return O(h)->resolve(resolver); }
int32_t cxo_resolve_factor_dependencies(void* h, int32_t resolver, int64_t f) { return O(h)->resolve_one(resolver, f, true); }
int32_t cxo_resolve_variable_dependencies(void* h, int32_t resolver, int64_t v) { return O(h)->resolve_one(resolver, v, false); }
int32_t cxo_set_signal_variant(void* h, int64_t s, int32_t kind, int64_t variable_id, int64_t factor_id) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size() || kind < CXB_KIND_UNSPECIFIED || kind > CXB_KIND_JOINT ||
        variable_id >= o->n_ids || factor_id >= o->n_ids || (variable_id >= 0 && o->is_factor[variable_id]) ||
        (factor_id >= 0 && !o->is_factor[factor_id])) {
        o->err = "set_signal_variant: bad argument";
        return CXB_ERR_BAD_ARG;
    }
    o->sig[s].kind = kind;
    o->sig[s].var = variable_id;
    o->sig[s].fac = factor_id;
    return CXB_OK;
}
int32_t cxo_link_signal(void* h, int64_t v, int64_t s) {
    Oracle* o = O(h);
    if (v < 0 || v >= o->n_ids || o->is_factor[v] || s < 0 || s >= (int64_t)o->sig.size()) {
        o->err = "link_signal: bad argument";
        return CXB_ERR_BAD_ARG;
    }
    o->linked[v].push_back(s);
    return CXB_OK;
}
int32_t cxo_link_signals(void* h, int64_t n, const int64_t* vs, const int64_t* ss) {
    for (int64_t i = 0; i < n; ++i) {
        int32_t st = cxo_link_signal(h, vs[i], ss[i]);
        if (st) return st;
    }
    return CXB_OK;
}
int64_t cxo_n_signals(void* h) { return (int64_t)O(h)->sig.size(); }
int64_t cxo_signal_id(void* h, int32_t kind, int64_t v, int64_t f) {
    Oracle* o = O(h);
    if (v < 0 || v >= o->n_ids) return -1;
    if (kind == CXB_KIND_MARGINAL) return o->marg_of[v];
    if (f < 0 || f >= o->n_ids) return -1;
    if (kind == CXB_KIND_M2V) return o->m2v(v, f);
    if (kind == CXB_KIND_M2F) return o->m2f(v, f);
    return -1;
}
int32_t cxo_signal_info(void* h, int64_t s, int64_t out[5]) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size()) return CXB_ERR_BAD_ARG;
    const Sig& g = o->sig[s];
    out[0] = g.kind;
    out[1] = g.var;
    out[2] = g.fac;
    out[3] = g.r0;
    out[4] = g.r1;
    return CXB_OK;
}
int64_t cxo_get_dependencies(void* h, int64_t s, int64_t* out_ids, uint8_t* out_nib, int64_t cap) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size()) return -1;
    const Sig& g = o->sig[s];
    for (int64_t i = 0; i < g.ndeps && i < cap; ++i) {
        if (out_ids) out_ids[i] = g.deps[i];
        if (out_nib) out_nib[i] = (uint8_t)((g.chunks[i >> 4] >> ((i & 15) << 2)) & 0xF);
    }
    return g.ndeps;
}
int64_t cxo_get_listeners(void* h, int64_t s, int64_t* out_ids, uint8_t* out_listen, int64_t cap) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size()) return -1;
    const Sig& g = o->sig[s];
    for (int64_t i = 0; i < (int64_t)g.listeners.size() && i < cap; ++i) {
        if (out_ids) out_ids[i] = g.listeners[i];
        if (out_listen) out_listen[i] = g.listenmask[i];
    }
    return (int64_t)g.listeners.size();
}
int64_t cxo_get_warnings(void* h, int64_t* out, int64_t cap) {
    Oracle* o = O(h);
    for (int64_t i = 0; i < (int64_t)o->warnings.size() && i < cap; ++i) out[i] = o->warnings[i];
    return (int64_t)o->warnings.size();
}
int32_t cxo_set_values(void* h, int64_t n, const int64_t* sids, const double* values, int64_t stride) {
    Oracle* o = O(h);
    for (int64_t i = 0; i < n; ++i) {
        if (sids[i] < 0 || sids[i] >= (int64_t)o->sig.size()) {
            o->err = "set_values: bad signal id";
            return CXB_ERR_BAD_ARG;
        }
        o->set_value(sids[i], values + i * stride);
    }
    return CXB_OK;
}
int32_t cxo_get_values(void* h, int64_t n, const int64_t* sids, double* out, int64_t stride) {
    Oracle* o = O(h);
    for (int64_t i = 0; i < n; ++i) {
        if (sids[i] < 0 || sids[i] >= (int64_t)o->sig.size()) {
            o->err = "get_values: bad signal id";
            return CXB_ERR_BAD_ARG;
        }
        std::memcpy(out + i * stride, o->value(sids[i]), sizeof(double) * o->dim);
    }
    return CXB_OK;
}
int32_t cxo_is_pending(void* h, int64_t s) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size()) return -1;
    return o->is_pending(s) ? 1 : 0;
}
int32_t cxo_is_computed(void* h, int64_t s) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size()) return -1;
    return o->sig[s].computed ? 1 : 0;
}
// raw (pp, p) without the lazy evaluation: bit0 = is_potentially_pending, bit1 = is_pending
int32_t cxo_raw_props(void* h, int64_t s) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size()) return -1;
    return (o->sig[s].pp ? 1 : 0) | (o->sig[s].p ? 2 : 0);
}
int32_t cxo_request_inference(void* h, int64_t n, const int64_t* ids) { return O(h)->request(n, ids); }
// literal scan_inference_request order (DFS, duplicates possible)
int64_t cxo_scan_dfs(void* h, int64_t* out, int64_t cap) {
    std::vector<int64_t> v;
    O(h)->scan(v);
    for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[i];
    return (int64_t)v.size();
}
// same set, ascending signal id, de-duplicated (what the device reports)
int64_t cxo_scan(void* h, int64_t* out, int64_t cap) {
    std::vector<int64_t> v;
    O(h)->scan(v);
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[i];
    return (int64_t)v.size();
}
// prepared signal lists / requests (include/cortex_b200.h): the oracle's engine dtype is float64
int64_t cxo_prepare_signals(void* h, int64_t n, const int64_t* signals) {
    Oracle* o = O(h);
    if (n <= 0 || !signals) return -1;
    std::unordered_set<int64_t> seen;
    for (int64_t i = 0; i < n; ++i) {
        if (signals[i] < 0 || signals[i] >= (int64_t)o->sig.size() || !seen.insert(signals[i]).second) {
            o->err = "prepare_signals: bad or repeated signal id";
            return -1;
        }
    }
    for (int64_t i = 0; i < n; ++i)
        for (int64_t d : o->sig[signals[i]].deps)
            if (seen.count(d)) {
                o->err = "prepare_signals: the signals depend on each other (sequential set_value! semantics need cxb_set_values)";
                return -1;
            }
    o->lists.emplace_back(signals, signals + n);
    return (int64_t)o->lists.size() - 1;
}
int32_t cxo_set_values_prepared(void* h, int64_t list, const void* values, int32_t /*values_on_device*/) {
    Oracle* o = O(h);
    if (list < 0 || list >= (int64_t)o->lists.size() || !values) return CXB_ERR_BAD_ARG;
    const std::vector<int64_t>& L = o->lists[(size_t)list];
    for (size_t i = 0; i < L.size(); ++i) o->set_value(L[i], (const double*)values + i * o->dim);
    return CXB_OK;
}
int32_t cxo_get_values_prepared(void* h, int64_t list, void* out, int32_t /*out_on_device*/) {
    Oracle* o = O(h);
    if (list < 0 || list >= (int64_t)o->lists.size() || !out) return CXB_ERR_BAD_ARG;
    const std::vector<int64_t>& L = o->lists[(size_t)list];
    for (size_t i = 0; i < L.size(); ++i) std::memcpy((double*)out + i * o->dim, o->value(L[i]), sizeof(double) * o->dim);
    return CXB_OK;
}
int64_t cxo_prepare_request(void* h, int64_t n, const int64_t* ids) {
    Oracle* o = O(h);
    if (n < 0 || (n > 0 && !ids)) return -1;
    o->prepared.emplace_back(ids, ids + n);
    return (int64_t)o->prepared.size() - 1;
}
// cxb_set_schedule: AUTO / SEQUENTIAL = the reference's own loop (seq), LEVEL = the level-synchronous schedule
int32_t cxo_set_schedule(void* h, int32_t schedule) {
    if (schedule < CXB_SCHEDULE_AUTO || schedule > CXB_SCHEDULE_SEQUENTIAL) return CXB_ERR_BAD_ARG;
    O(h)->schedule = schedule;
    return CXB_OK;
}
int32_t cxo_last_schedule(void* h) { return O(h)->schedule == CXB_SCHEDULE_LEVEL ? CXB_SCHEDULE_LEVEL : CXB_SCHEDULE_SEQUENTIAL; }
int32_t cxo_update_marginals(void* h, int64_t n, const int64_t* ids, cxb_update_stats* stats) {
    int32_t st = O(h)->schedule == CXB_SCHEDULE_LEVEL ? O(h)->update_lvl(n, ids) : O(h)->update_seq(n, ids);
    if (stats) *stats = O(h)->stats;
    return st;
}
int32_t cxo_update_marginals_prepared(void* h, int64_t request, cxb_update_stats* stats) {
    Oracle* o = O(h);
    if (request < 0 || request >= (int64_t)o->prepared.size()) return CXB_ERR_BAD_ARG;
    const std::vector<int64_t> ids = o->prepared[(size_t)request];
    return cxo_update_marginals(h, (int64_t)ids.size(), ids.data(), stats);
}
int32_t cxo_update_marginals_seq(void* h, int64_t n, const int64_t* ids, cxb_update_stats* stats) {
    int32_t st = O(h)->update_seq(n, ids);
    if (stats) *stats = O(h)->stats;
    return st;
}
int32_t cxo_trace_enable(void*, int32_t) { return CXB_OK; }
// level trace; for the seq schedule out_level holds the loop round (final phase = -1)
int64_t cxo_trace_get(void* h, int64_t* out_level, int64_t* out_sid, int64_t cap) {
    Oracle* o = O(h);
    for (int64_t i = 0; i < (int64_t)o->trace.size() && i < cap; ++i) {
        if (out_level) out_level[i] = o->trace[i].round;
        if (out_sid) out_sid[i] = o->trace[i].sid;
    }
    return (int64_t)o->trace.size();
}
int64_t cxo_trace_get_times(void* h, int64_t* out_ns, int64_t cap) {
    Oracle* o = O(h);
    for (int64_t i = 0; i < (int64_t)o->trace.size() && i < cap; ++i) out_ns[i] = o->trace[i].ns;
    return (int64_t)o->trace.size();
}
int64_t cxo_trace_get_variables(void* h, int64_t* out_var, int64_t cap) {
    Oracle* o = O(h);
    for (int64_t i = 0; i < (int64_t)o->trace.size() && i < cap; ++i) out_var[i] = o->trace[i].var;
    return (int64_t)o->trace.size();
}
int64_t cxo_count_is_pending_calls(void* h) { return O(h)->n_is_pending_calls; }

// compute!(strategy, signal; force, skip_if_no_listeners), src/signal.jl:392-410
int32_t cxo_compute(void* h, int64_t s, int32_t force, int32_t skip_if_no_listeners) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size()) return CXB_ERR_BAD_ARG;
    return o->compute(s, force != 0, skip_if_no_listeners != 0, true);
}
// process_dependencies!(f, signal; retry) with the callback given as a table: f(dep) = answers[dep] (answers == NULL:
// f = is_pending, the scanner's callback). Writes the visit sequence; returns the number of visits (may exceed cap).
int64_t cxo_process_dependencies_table(void* h, int64_t s, int32_t retry, const uint8_t* answers, int64_t* out_visited, int64_t cap,
                                       int32_t* processed_out) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size()) return -1;
    std::vector<int64_t> seen;
    bool r = o->process_dependencies(s, retry != 0, [&](int64_t d) {
        seen.push_back(d);
        return answers ? answers[d] != 0 : o->is_pending(d);
    });
    if (processed_out) *processed_out = r ? 1 : 0;
    for (int64_t i = 0; i < (int64_t)seen.size() && i < cap; ++i) out_visited[i] = seen[i];
    return (int64_t)seen.size();
}
// process_dependencies!(f, signal; retry), src/signal.jl:466-490
int32_t cxo_process_dependencies(void* h, int64_t s, int32_t retry, visit_cb_t f, void* user) {
    Oracle* o = O(h);
    if (s < 0 || s >= (int64_t)o->sig.size()) return -1;
    return o->process_dependencies(s, retry != 0, [&](int64_t d) { return f(user, d) != 0; }) ? 1 : 0;
}

// ---- dense CPU kernels for the structured configs (cpu_baseline of bench.py) -----------------
// One linear-Gaussian chain batch, the same six message classes as cxb_chains_* in fp64, computed with
// the canonical-form rules above in dependency order. y[T][B], out[6][T][B][2].
int32_t cxo_chains_reference(int64_t B, int64_t T, const double* q, const double* r, const double* y, double* out) {
    auto at = [&](int m, int64_t t, int64_t b) { return out + (((size_t)m * T + t) * B + b) * 2; };
    for (int64_t b = 0; b < B; ++b) {
        for (int64_t t = 0; t < T; ++t) {  // forward round
            double* obs = at(0, t, b);
            obs[0] = 1.0 / r[b];
            obs[1] = y[t * B + b] / r[b];
            double* mf = at(2, t, b);
            if (t == 0) {
                mf[0] = obs[0];
                mf[1] = obs[1];
                at(1, t, b)[0] = 0;
                at(1, t, b)[1] = 0;
            } else {
                const double* pm = at(2, t - 1, b);
                double den = 1.0 + q[b] * pm[0];
                double* pr = at(1, t, b);
                pr[0] = pm[0] / den;
                pr[1] = pm[1] / den;
                mf[0] = obs[0] + pr[0];
                mf[1] = obs[1] + pr[1];
            }
        }
        for (int64_t t = T - 1; t >= 0; --t) {  // reverse round
            const double* obs = at(0, t, b);
            double* bw = at(3, t, b);   // m2v(x_t, tr_t)
            double* mb = at(4, t, b);   // m2f(x_t, tr_{t-1})
            if (t == T - 1) {
                bw[0] = 0;
                bw[1] = 0;
                mb[0] = obs[0];
                mb[1] = obs[1];
            } else {
                const double* nm = at(4, t + 1, b);
                double den = 1.0 + q[b] * nm[0];
                bw[0] = nm[0] / den;
                bw[1] = nm[1] / den;
                mb[0] = obs[0] + bw[0];
                mb[1] = obs[1] + bw[1];
            }
            double* mg = at(5, t, b);  // marginal: lik, tr_{t-1}, tr_t left-to-right
            double a0 = obs[0], a1 = obs[1];
            if (t > 0) {
                a0 = a0 + at(1, t, b)[0];
                a1 = a1 + at(1, t, b)[1];
            }
            if (t < T - 1) {
                a0 = a0 + bw[0];
                a1 = a1 + bw[1];
            }
            mg[0] = a0;
            mg[1] = a1;
        }
    }
    return CXB_OK;
}

}  // extern "C"
