"""Model builders shared by the tests, bench.py's CPU leg and smoke(): the reference's own test models
(test/inference_engine_tests.jl) and the benchmark graph families of BASELINE.json at reduced size."""
from __future__ import annotations

import numpy as np

from tests._pkg import pkg

C = pkg
cap = pkg.capi


def make_ssm_model(n, api, *, form="canon", dtype=cap.F64, q=1.0, r=1.0, resolver=None, processor=None, trace=False):
    """test/inference_engine_tests.jl:436-462: x_i, y_i, likelihood_i (y_i, x_i), transition_i (x_i, x_{i+1})."""
    g = C.BipartiteFactorGraph()
    x = [g.add_variable(C.Variable(name="x", index=(i,))) for i in range(n)]
    y = [g.add_variable(C.Variable(name="y", index=(i,))) for i in range(n)]
    lik = [g.add_factor(C.Factor(functional_form="likelihood")) for _ in range(n)]
    tr = [g.add_factor(C.Factor(functional_form="transition")) for _ in range(n - 1)]
    for i in range(n):
        g.add_edge(y[i], lik[i], C.Connection(label="out"))
        g.add_edge(x[i], lik[i], C.Connection(label="out"))
    for i in range(n - 1):
        g.add_edge(x[i], tr[i], C.Connection(label="out"))
        g.add_edge(x[i + 1], tr[i], C.Connection(label="in"))
    if processor is None:
        if form == "canon":
            processor = C.RuleProcessor({"likelihood": (cap.RULE_GAUSS_OBS, [r]), "transition": (cap.RULE_GAUSS_RW, [q])},
                                        family=cap.FAMILY_GAUSS_CANON, value_dim=2)
        else:  # the reference fixture's own moment form
            processor = C.RuleProcessor({"likelihood": (cap.RULE_GAUSS_MV_OBS, [r]), "transition": (cap.RULE_GAUSS_MV_RW, [q])},
                                        family=cap.FAMILY_GAUSS_MV, value_dim=2)
    engine = C.InferenceEngine(model_engine=g, dependency_resolver=resolver or C.DefaultDependencyResolver(),
                               inference_request_processor=processor, dtype=dtype, api=api, trace=trace)
    return engine, x, y, lik, tr


def ssm_set_data(engine, y, lik, data):
    sigs = [C.get_connection_message_to_factor(engine, y[i], lik[i]) for i in range(len(y))]
    C.set_values(sigs, np.asarray(data, dtype=np.float64).reshape(-1, 1))


def make_beta_bernoulli_model(n, api, dtype=cap.F64):
    """test/inference_engine_tests.jl:316-341."""
    g = C.BipartiteFactorGraph()
    p = g.add_variable(C.Variable(name="p"))
    o, f = [], []
    for i in range(n):
        oi = g.add_variable(C.Variable(name="o", index=(i,)))
        fi = g.add_factor(C.Factor(functional_form="bernoulli"))
        o.append(oi)
        f.append(fi)
        g.add_edge(p, fi, C.Connection(label="out"))
        g.add_edge(oi, fi, C.Connection(label="out"))
    proc = C.RuleProcessor({"bernoulli": (cap.RULE_BETA_BERNOULLI, [])}, family=cap.FAMILY_BETA, value_dim=2)
    engine = C.InferenceEngine(model_engine=g, dependency_resolver=C.DefaultDependencyResolver(),
                               inference_request_processor=proc, dtype=dtype, api=api)
    return engine, p, o, f


def potts_table(K, beta):
    return np.exp(beta * np.eye(K))


def make_grid_model(H, W, K, beta, api, dtype=cap.F64, rule="potts", link=True, ghost_top=False, ghost_bottom=False):
    """BASELINE config 4 at reduced size: pixels row-major, one unary leaf factor per pixel, pairwise factors
    on the 4-neighbourhood (horizontal edges then vertical, ascending ids), protocol B linked signals.

    ghost_top / ghost_bottom (row sharding): an extra row of *ghost* pixels stands for the neighbour
    shard's boundary row. A ghost has exactly one factor (the cut edge), so its m2f is a plain data signal that the
    halo exchange sets every sweep; `pix` then has H + ghosts rows and only rows `real` are requested."""
    g = C.BipartiteFactorGraph()
    R = H + int(ghost_top) + int(ghost_bottom)
    is_ghost = [ghost_top and i == 0 or ghost_bottom and i == R - 1 for i in range(R)]
    pix = [[g.add_variable(C.Variable(name="s", index=(i, j))) for j in range(W)] for i in range(R)]
    un = [[None if is_ghost[i] else g.add_factor(C.Factor(functional_form="unary")) for j in range(W)] for i in range(R)]
    for i in range(R):
        for j in range(W):
            if not is_ghost[i]:
                g.add_edge(pix[i][j], un[i][j], C.Connection(label="out"))
    pair = []
    for i in range(R):
        for j in range(W):
            if j + 1 < W and not is_ghost[i]:
                f = g.add_factor(C.Factor(functional_form="pair"))
                g.add_edge(pix[i][j], f, C.Connection(label="a"))
                g.add_edge(pix[i][j + 1], f, C.Connection(label="b"))
                pair.append((f, pix[i][j], pix[i][j + 1]))
            if i + 1 < R:
                f = g.add_factor(C.Factor(functional_form="pair"))
                g.add_edge(pix[i][j], f, C.Connection(label="a"))
                g.add_edge(pix[i + 1][j], f, C.Connection(label="b"))
                pair.append((f, pix[i][j], pix[i + 1][j]))
    if rule == "potts":
        rules = {"pair": (cap.RULE_POTTS, [beta])}
    else:
        rules = {"pair": (cap.RULE_CAT_TABLE, potts_table(K, beta).ravel())}
    proc = C.RuleProcessor(rules, family=cap.FAMILY_CATEGORICAL, value_dim=K)
    engine = C.InferenceEngine(model_engine=g, dependency_resolver=C.DefaultDependencyResolver(),
                               inference_request_processor=proc, dtype=dtype, api=api)
    if link:
        protocol_b_link(engine, [v for i, row in enumerate(pix) if not is_ghost[i] for v in row])
    return engine, pix, un, pair


def protocol_b_link(engine, variable_ids):
    """link every non-leaf m2f(v,f) to its variable."""
    for v in variable_ids:
        var = C.get_variable(engine, v)
        for f in C.get_connected_factor_ids(engine, v):
            if len(C.get_connected_variable_ids(engine, f)) > 1:
                C.link_signal_to_variable(var, C.get_connection_message_to_factor(engine, v, f))


def protocol_b_init(engine, variable_ids, K):
    sigs = []
    for v in variable_ids:
        for f in C.get_connected_factor_ids(engine, v):
            if len(C.get_connected_variable_ids(engine, f)) > 1:
                sigs.append(C.get_connection_message_to_factor(engine, v, f))
    C.set_values(sigs, np.full((len(sigs), K), 1.0 / K))


def protocol_b_sweep(engine, variable_ids, unary_signals, unary_values, schedule="lvl"):
    """Re-assert the evidence, then one update_marginals!(engine, all) (Appendix B step 2)."""
    C.set_values(unary_signals, unary_values)
    return C.update_marginals(engine, variable_ids, schedule=schedule)


def chung_lu_edges(n, m, alpha=2.5, seed=1234):
    """BASELINE config 5 generator: weights w_i ∝ (i+10)^(-1/(alpha-1)); m distinct pairs, no self loops,
    returned sorted so that factor ids ascend with (u, v)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = (np.arange(n) + 10.0) ** (-1.0 / (alpha - 1.0))
    p = w / w.sum()
    pairs = set()
    while len(pairs) < m:
        need = m - len(pairs)
        a = rng.choice(n, size=2 * need + 16, p=p)
        b = rng.choice(n, size=2 * need + 16, p=p)
        for u, v in zip(a, b):
            if u != v:
                pairs.add((min(u, v), max(u, v)))
                if len(pairs) == m:
                    break
    return sorted((int(u), int(v)) for u, v in pairs)


def make_powerlaw_model(n, m, K, api, dtype=cap.F64, n_tables=16, seed=1234, link=True):
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    edges = chung_lu_edges(n, m, seed=seed)
    tables = np.exp(rng.standard_normal((n_tables, K, K)))
    ttype = rng.integers(0, n_tables, size=len(edges))
    g = C.BipartiteFactorGraph()
    vs = [g.add_variable(C.Variable(name="v", index=(i,))) for i in range(n)]
    un = [g.add_factor(C.Factor(functional_form="unary")) for _ in range(n)]
    for i in range(n):
        g.add_edge(vs[i], un[i], C.Connection(label="out"))
    pair = []
    for (u, v), t in zip(edges, ttype):
        f = g.add_factor(C.Factor(functional_form=f"pair{int(t)}"))
        g.add_edge(vs[u], f, C.Connection(label="a"))
        g.add_edge(vs[v], f, C.Connection(label="b"))
        pair.append((f, vs[u], vs[v]))
    rules = {f"pair{t}": (cap.RULE_CAT_TABLE, tables[t].ravel()) for t in range(n_tables)}
    proc = C.RuleProcessor(rules, family=cap.FAMILY_CATEGORICAL, value_dim=K)
    engine = C.InferenceEngine(model_engine=g, dependency_resolver=C.DefaultDependencyResolver(),
                               inference_request_processor=proc, dtype=dtype, api=api)
    if link:
        protocol_b_link(engine, vs)
    unary = rng.dirichlet(np.ones(K), size=n)
    return engine, vs, un, pair, unary, tables, ttype


def engine_state(engine):
    """Observable state of every signal: (is_computed, is_pending, nibbles, value) — for seq-vs-lvl and
    oracle-vs-device comparisons. NOTE: evaluates is_pending (lazy cache) on every signal."""
    st = engine.store
    out = []
    for s in range(st.n_signals()):
        sig = C.Signal(st, s)
        out.append((C.is_computed(sig), C.is_pending(sig), tuple(C.get_dependency_props(sig))))
    ids = [C.Signal(st, s) for s in range(st.n_signals())]
    vals = C.get_values(ids) if ids else np.zeros((0, st.value_dim))
    return out, vals


TOL = {cap.F64: 1e-12, cap.F32: 1e-5}  # north_star: rel 1e-5 (fp32) / 1e-12 (fp64) on marginals


def assert_values_close(got, want, dtype, kind="prob", err_msg=""):
    """Element-wise comparison at the north-star tolerance.

    kind="prob": normalised categorical vectors - |got - want| <= tol * |want| + tol * 1e-3 (the absolute term is the floor of
    a vector that sums to one: 1e-8 in fp32, 1e-15 in fp64; a state of probability 1e-6 must still be right to 1 %).
    kind="canon": Gaussian messages in canonical form (precision, precision * mean), last axis 2: the precision element-wise
    (rel tol), and as (mean, variance) - SURVEY Appendix C "marginals reported as (mean, variance)" - the variance rel tol,
    the mean to tol * (|mean| + standard deviation): a mean that crosses zero has no relative error of its own, its
    natural scale is the posterior standard deviation. The vacuous message (0, 0) must be reproduced exactly.
    kind="plain": element-wise rel tol with the 1e-3 * tol floor."""
    tol = TOL[dtype]
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape, err_msg)
    if kind in ("prob", "plain"):
        np.testing.assert_allclose(got, want, rtol=tol, atol=tol * 1e-3, err_msg=err_msg)
        return
    if kind == "mp":  # NormalMeanPrecision (mean, precision): the precision rel tol, the mean to tol * (|mean| + standard deviation)
        np.testing.assert_allclose(got[..., 1], want[..., 1], rtol=tol, atol=0, err_msg=err_msg + " (precision)")
        bound = tol * (np.abs(want[..., 0]) + 1.0 / np.sqrt(want[..., 1]))
        bad = np.abs(got[..., 0] - want[..., 0]) > bound
        assert not bad.any(), f"{err_msg} (mean): {int(bad.sum())} of {bad.size} differ"
        return
    assert kind == "canon" and want.shape[-1] == 2
    lam_w, eta_w, lam_g, eta_g = want[..., 0], want[..., 1], got[..., 0], got[..., 1]
    vac = lam_w == 0
    assert np.array_equal(lam_g[vac], lam_w[vac]) and np.array_equal(eta_g[vac], eta_w[vac]), err_msg
    other = lam_w < 0  # not a Gaussian message (an engine stores the observations y_t in the same array): plain comparison
    np.testing.assert_allclose(got[other], want[other], rtol=tol, atol=tol * 1e-3, err_msg=err_msg)
    ok = lam_w > 0
    lw, ew, lg, eg = lam_w[ok], eta_w[ok], lam_g[ok], eta_g[ok]
    np.testing.assert_allclose(lg, lw, rtol=tol, atol=0, err_msg=err_msg + " (precision)")
    np.testing.assert_allclose(1.0 / lg, 1.0 / lw, rtol=tol, atol=0, err_msg=err_msg + " (variance)")
    mean_w, mean_g = ew / lw, eg / lg
    bound = tol * (np.abs(mean_w) + np.sqrt(1.0 / lw))
    bad = np.abs(mean_g - mean_w) > bound
    assert not bad.any(), f"{err_msg} (mean): {int(bad.sum())} of {bad.size} differ, worst ratio {np.max(np.abs(mean_g - mean_w) / bound):.3g}"


def canon_to_mv(v):
    v = np.asarray(v, dtype=np.float64)
    return np.stack([v[..., 1] / v[..., 0], 1.0 / v[..., 0]], axis=-1)


def rts_smoother(y, q, r):
    """Independent numpy Kalman filter + RTS smoother for the random-walk model with a flat prior on x_1
    (returns means, variances) — cross-check that does not share code with the oracle."""
    T = len(y)
    mf, Pf = np.zeros(T), np.zeros(T)
    mf[0], Pf[0] = y[0], r
    for t in range(1, T):
        mp, Pp = mf[t - 1], Pf[t - 1] + q
        k = Pp / (Pp + r)
        mf[t] = mp + k * (y[t] - mp)
        Pf[t] = (1 - k) * Pp
    ms, Ps = mf.copy(), Pf.copy()
    for t in range(T - 2, -1, -1):
        Pp = Pf[t] + q
        g = Pf[t] / Pp
        ms[t] = mf[t] + g * (ms[t + 1] - mf[t])
        Ps[t] = Pf[t] + g * g * (Ps[t + 1] - Pp)
    return ms, Ps


def make_hmm_model(T, K, M, A, E, api, dtype=cap.F64):
    """SURVEY Appendix C HMM graph for one chain: z_t, y_t, prior leaf on z_1, emission (y_t,z_t), transition (z_t,z_{t+1})."""
    g = C.BipartiteFactorGraph()
    z = [g.add_variable(C.Variable(name="z", index=(t,))) for t in range(T)]
    y = [g.add_variable(C.Variable(name="y", index=(t,))) for t in range(T)]
    prior = g.add_factor(C.Factor(functional_form="prior"))
    em = [g.add_factor(C.Factor(functional_form="emission")) for _ in range(T)]
    tr = [g.add_factor(C.Factor(functional_form="transition")) for _ in range(T - 1)]
    g.add_edge(z[0], prior, C.Connection(label="out"))
    for t in range(T):
        g.add_edge(y[t], em[t], C.Connection(label="out"))
        g.add_edge(z[t], em[t], C.Connection(label="in"))
    for t in range(T - 1):
        g.add_edge(z[t], tr[t], C.Connection(label="in"))
        g.add_edge(z[t + 1], tr[t], C.Connection(label="out"))
    rules = {"emission": (cap.RULE_HMM_EMIT, np.concatenate([[float(M)], np.asarray(E, dtype=np.float64).ravel()])),
             "transition": (cap.RULE_CAT_TABLE, np.asarray(A, dtype=np.float64).ravel())}
    proc = C.RuleProcessor(rules, family=cap.FAMILY_CATEGORICAL, value_dim=K)
    engine = C.InferenceEngine(model_engine=g, dependency_resolver=C.DefaultDependencyResolver(),
                               inference_request_processor=proc, dtype=dtype, api=api)
    return engine, z, y, prior, em, tr


def hmm_set_data(engine, z, y, prior, em, obs, K):
    sig = [C.get_connection_message_to_factor(engine, y[t], em[t]) for t in range(len(y))]
    vals = np.zeros((len(y), K))
    vals[:, 0] = obs
    C.set_values(sig, vals)
    C.set_value(C.get_connection_message_to_variable(engine, z[0], prior), np.full(K, 1.0 / K))


def level_trace(engine):
    """{level: sorted signal ids} of the last update_marginals (levels >= 0 loop phase, -1 marginals, -2 linked)."""
    n = engine.api.trace_get(engine.store.h, None, None, 0)
    lv = np.zeros(max(n, 1), dtype=np.int64)
    sg = np.zeros(max(n, 1), dtype=np.int64)
    engine.api.trace_get(engine.store.h, lv.ctypes.data_as(cap.i64p), sg.ctypes.data_as(cap.i64p), n)
    out = {}
    for l, s in zip(lv[:n].tolist(), sg[:n].tolist()):
        out.setdefault(l, []).append(s)
    return {l: sorted(v) for l, v in out.items()}


def chung_lu_edges_fast(n, m, alpha=2.5, seed=1234):
    """Vectorised Chung-Lu generator for benchmark-sized graphs (same weights as chung_lu_edges): m distinct pairs
    (u < v), sorted, as an (m, 2) int64 array."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = (np.arange(n) + 10.0) ** (-1.0 / (alpha - 1.0))
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    have = np.zeros(0, dtype=np.int64)
    while have.size < m:
        need = int((m - have.size) * 1.15) + 1024
        a = np.searchsorted(cdf, rng.random(need))
        b = np.searchsorted(cdf, rng.random(need))
        keep = a != b
        lo, hi = np.minimum(a[keep], b[keep]), np.maximum(a[keep], b[keep])
        have = np.unique(np.concatenate([have, lo.astype(np.int64) * n + hi]))
    if have.size > m:
        have = np.sort(rng.choice(have, size=m, replace=False))
    return np.stack([have // n, have % n], axis=1)


# ---- variational message passing: the mean-field SSM of test/inference_engine_tests.jl:593-809 -------------------------
def make_ssm_mean_field_model(n, api, *, dtype=cap.F64, resolver=None, processor=None):
    """test/inference_engine_tests.jl:698-741 — ssnoise, obsnoise (Gamma), x_i (NormalMeanPrecision), y_i (observed);
    likelihood_i (y_i, x_i, obsnoise), transition_i (x_i, x_{i+1}, ssnoise); same creation and edge order."""
    g = C.BipartiteFactorGraph()
    ssnoise = g.add_variable(C.Variable(name="ssnoise"))
    obsnoise = g.add_variable(C.Variable(name="obsnoise"))
    x = [g.add_variable(C.Variable(name="x", index=(i,))) for i in range(n)]
    y = [g.add_variable(C.Variable(name="y", index=(i,))) for i in range(n)]
    lik = [g.add_factor(C.Factor(functional_form="likelihood")) for _ in range(n)]
    tr = [g.add_factor(C.Factor(functional_form="transition")) for _ in range(n - 1)]
    for i in range(n):
        g.add_edge(y[i], lik[i], C.Connection(label="out"))
        g.add_edge(x[i], lik[i], C.Connection(label="out"))
        g.add_edge(obsnoise, lik[i], C.Connection(label="out"))
    for i in range(n - 1):
        g.add_edge(x[i], tr[i], C.Connection(label="out"))
        g.add_edge(x[i + 1], tr[i], C.Connection(label="in"))
        g.add_edge(ssnoise, tr[i], C.Connection(label="out"))
    builtin = processor is None
    if builtin:
        processor = C.RuleProcessor({"likelihood": (cap.RULE_NORMAL_MEAN_FIELD, []), "transition": (cap.RULE_NORMAL_MEAN_FIELD, [])},
                                    family=cap.FAMILY_GAUSS_MP, value_dim=2)
    engine = C.InferenceEngine(model_engine=g, dependency_resolver=resolver or C.MeanFieldResolver(),
                               inference_request_processor=processor, dtype=dtype, api=api)
    if builtin:  # the value types the reference carries in the Julia values themselves
        C.set_variable_families(engine, [ssnoise, obsnoise], cap.FAMILY_GAMMA)
        C.set_variable_families(engine, y, cap.FAMILY_POINT)
    # initial marginals, :728-738
    C.set_value(C.get_variable_marginal(C.get_variable(engine, ssnoise)), [1.0, 1.0])   # Gamma(1, 1)
    C.set_value(C.get_variable_marginal(C.get_variable(engine, obsnoise)), [1.0, 1.0])
    C.set_values([C.get_variable_marginal(C.get_variable(engine, v)) for v in x], np.tile([0.0, 1.0], (n, 1)))  # N(0, 1)
    return engine, x, y, obsnoise, ssnoise, lik, tr


def ssm_mean_field_experiment(engine, x, y, obsnoise, ssnoise, dataset, vmp_iterations, schedule="lvl"):
    """`experiment` of test/inference_engine_tests.jl:743-781, call for call (including the repeated and merged updates)."""
    n = len(dataset)
    C.set_values([C.get_variable_marginal(C.get_variable(engine, v)) for v in y],
                 np.stack([np.asarray(dataset, dtype=np.float64), np.zeros(n)], axis=1))
    up = lambda ids: C.update_marginals(engine, ids, schedule=schedule)  # noqa: E731
    for iteration in range(1, vmp_iterations + 1):
        if iteration // 2 == 0:  # div(iteration, 2) == 0, :753
            up(x), up(ssnoise), up(obsnoise)
        else:
            up(obsnoise), up(ssnoise), up(x)
        for _ in range(3):
            up(obsnoise)
        for _ in range(3):
            up(ssnoise)
        up([ssnoise, obsnoise])
    mg = lambda v: C.get_value(C.get_variable_marginal(C.get_variable(engine, v)))  # noqa: E731
    return {"x": C.get_values([C.get_variable_marginal(C.get_variable(engine, v)) for v in x]),
            "ssnoise": mg(ssnoise), "obsnoise": mg(obsnoise)}


def ssm_mean_field_dataset(n, seed=1234, ssnoise_real=100.0, obsnoise_real=100.0):
    """:785-796 with numpy's PCG64 instead of StableRNG (the assertions do not depend on the draws)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    walk = np.cumsum(np.concatenate([[0.0], rng.standard_normal(n - 1) / np.sqrt(ssnoise_real)]))
    return walk + rng.standard_normal(n) / np.sqrt(obsnoise_real)


# ---- structured VMP: test/inference_engine_tests.jl:811-1147 ------------------------------------------------------------
class StructuredResolver(C.AbstractDependencyResolver):
    """:811-907, literally: mean-field (weak) dependencies around the likelihood factors; around a transition factor the
    two states form one cluster with a JointMarginal signal (linked to both states, local marginal of the factor)."""

    def resolve_variable_dependencies(self, engine, variable_id):  # :813-815
        return C.DefaultDependencyResolver().resolve_variable_dependencies(engine, variable_id)

    def resolve_factor_dependencies(self, engine, factor_id):
        m2v = lambda v: C.get_connection_message_to_variable(engine, v, factor_id)  # noqa: E731
        m2f = lambda v: C.get_connection_message_to_factor(engine, v, factor_id)  # noqa: E731
        marg = lambda v: C.get_variable_marginal(C.get_variable(engine, v))  # noqa: E731
        connected = list(C.get_connected_variable_ids(engine, factor_id))
        if C.get_factor_functional_form(C.get_factor(engine, factor_id)) == "likelihood":
            for v1 in connected:
                for v2 in connected:
                    if v1 != v2:
                        C.add_dependency(m2v(v1), marg(v2), weak=True)
            return
        clusters = {}  # by variable name (:837-846); iteration = insertion order
        for v in connected:
            clusters.setdefault(C.get_variable_name(C.get_variable(engine, v)), []).append(v)
        deps = []
        for cluster in clusters.values():
            if len(cluster) == 1:
                deps.append(marg(cluster[0]))
                continue
            joint = C.create_inference_signal(engine)
            C.set_variant(joint, C.JointMarginal(factor_id, tuple(cluster)))
            for v in cluster:
                C.link_signal_to_variable(C.get_variable(engine, v), joint)
                C.add_local_marginal_to_factor(C.get_factor(engine, factor_id), joint)
                C.add_dependency(joint, m2f(v), weak=True)
            deps.append(joint)
        for d1 in deps:
            for d2 in deps:
                if C.isa_variant(d1, C.JointMarginal) and d1 != d2:
                    C.add_dependency(d1, d2, weak=True)
        for index, cluster in enumerate(clusters.values()):
            for m1 in cluster:
                for m2 in cluster:
                    if m1 != m2:
                        C.add_dependency(m2v(m1), m2f(m2))
            for m1 in cluster:
                for another_index, other in enumerate(deps):
                    if index != another_index:
                        C.add_dependency(m2v(m1), other, weak=True)


def make_ssm_structured_model(n, api, *, dtype=cap.F64, processor=None):
    """:1031-1076. Values: NormalMeanPrecision (mean, precision), Gamma (shape, scale), observation (y), and
    MvNormalMeanPrecision (mu1, mu2, W11, W12, W21, W22) for the joint marginals: value_dim = 6."""
    g = C.BipartiteFactorGraph()
    ssnoise = g.add_variable(C.Variable(name="ssnoise"))
    obsnoise = g.add_variable(C.Variable(name="obsnoise"))
    x = [g.add_variable(C.Variable(name="x", index=(i,))) for i in range(n)]
    y = [g.add_variable(C.Variable(name="y", index=(i,))) for i in range(n)]
    lik = [g.add_factor(C.Factor(functional_form="likelihood")) for _ in range(n)]
    tr = [g.add_factor(C.Factor(functional_form="transition")) for _ in range(n - 1)]
    for i in range(n):
        g.add_edge(y[i], lik[i], C.Connection(label="out"))
        g.add_edge(x[i], lik[i], C.Connection(label="out"))
        g.add_edge(obsnoise, lik[i], C.Connection(label="out"))
    for i in range(n - 1):
        g.add_edge(x[i], tr[i], C.Connection(label="out"))
        g.add_edge(x[i + 1], tr[i], C.Connection(label="in"))
        g.add_edge(ssnoise, tr[i], C.Connection(label="out"))
    builtin = processor is None
    if builtin:
        processor = C.RuleProcessor({"likelihood": (cap.RULE_NORMAL_MEAN_FIELD, []), "transition": (cap.RULE_NORMAL_STRUCTURED, [])},
                                    family=cap.FAMILY_GAUSS_MP, value_dim=6)
    engine = C.InferenceEngine(model_engine=g, dependency_resolver=StructuredResolver(),
                               inference_request_processor=processor, dtype=dtype, api=api)
    if builtin:
        C.set_variable_families(engine, [ssnoise, obsnoise], cap.FAMILY_GAMMA)
        C.set_variable_families(engine, y, cap.FAMILY_POINT)
    pad = lambda a, b: [a, b, 0.0, 0.0, 0.0, 0.0]  # noqa: E731
    C.set_value(C.get_variable_marginal(C.get_variable(engine, ssnoise)), pad(1.0, 1.0))
    C.set_value(C.get_variable_marginal(C.get_variable(engine, obsnoise)), pad(1.0, 1.0))
    C.set_values([C.get_variable_marginal(C.get_variable(engine, v)) for v in x], np.tile(pad(0.0, 1.0), (n, 1)))
    return engine, x, y, obsnoise, ssnoise, lik, tr


def ssm_structured_experiment(engine, x, y, obsnoise, ssnoise, dataset, vmp_iterations, schedule="lvl", merged_all=True):
    """`experiment` of :1078-1120, call for call. `merged_all=False` leaves out the last call of an iteration (all
    variables in one request): with more than 5 transition factors ssnoise goes through the segment tree of
    src/dependencies.jl:90-173, its request then recomputes the joint marginals BEFORE the state messages they listen
    to, i.e. the result depends on the order of the ids — the level-synchronous schedule refuses that request
    (CXB_ERR_OUT_OF_CONTRACT) instead of returning something order-dependent."""
    n = len(dataset)
    vals = np.zeros((n, 6))
    vals[:, 0] = np.asarray(dataset, dtype=np.float64)
    C.set_values([C.get_variable_marginal(C.get_variable(engine, v)) for v in y], vals)
    up = lambda ids: C.update_marginals(engine, ids, schedule=schedule)  # noqa: E731
    for iteration in range(1, vmp_iterations + 1):
        if iteration // 2 == 1:
            up(x), up(ssnoise), up(obsnoise)
        else:
            up(obsnoise), up(ssnoise), up(x)
        for _ in range(3):
            up(ssnoise)
        for _ in range(2):
            up(x)
        for _ in range(3):
            up(obsnoise)
        up([ssnoise, obsnoise])
        if merged_all:
            up([ssnoise, obsnoise] + list(x))
    mg = lambda v: C.get_value(C.get_variable_marginal(C.get_variable(engine, v)))  # noqa: E731
    return {"x": C.get_values([C.get_variable_marginal(C.get_variable(engine, v)) for v in x]),
            "ssnoise": mg(ssnoise), "obsnoise": mg(obsnoise)}


def make_ssm_batch_model(lengths, api, *, dtype=cap.F64, q=None, r=None, interleave=False):
    """A batch of independent random-walk chains (test/inference_engine_tests.jl:436-462 per chain) in ONE graph, built
    through the generic frontend. Returns (engine, xs, ys, liks, trs) with one list per chain. Per-chain noise variances
    go in as per-factor parameters (cxb_set_factor_params). `interleave` creates the variables chain-interleaved (time-
    major ids) instead of chain after chain."""
    g = C.BipartiteFactorGraph()
    B = len(lengths)
    xs = [[None] * T for T in lengths]
    ys = [[None] * T for T in lengths]
    order = [(b, t) for t in range(max(lengths)) for b in range(B) if t < lengths[b]] if interleave else [(b, t) for b in range(B) for t in range(lengths[b])]
    for b, t in order:
        xs[b][t] = g.add_variable(C.Variable(name="x", index=(b, t)))
    for b, t in order:
        ys[b][t] = g.add_variable(C.Variable(name="y", index=(b, t)))
    liks = [[None] * T for T in lengths]
    trs = [[None] * max(T - 1, 0) for T in lengths]
    for b, t in order:
        liks[b][t] = g.add_factor(C.Factor(functional_form="likelihood"))
    for b, t in order:
        if t + 1 < lengths[b]:
            trs[b][t] = g.add_factor(C.Factor(functional_form="transition"))
    for b, t in order:
        g.add_edge(ys[b][t], liks[b][t], C.Connection(label="out"))
        g.add_edge(xs[b][t], liks[b][t], C.Connection(label="out"))
    for b, t in order:
        if t + 1 < lengths[b]:
            g.add_edge(xs[b][t], trs[b][t], C.Connection(label="out"))
            g.add_edge(xs[b][t + 1], trs[b][t], C.Connection(label="in"))
    processor = C.RuleProcessor({"likelihood": (cap.RULE_GAUSS_OBS, [1.0]), "transition": (cap.RULE_GAUSS_RW, [1.0])},
                                family=cap.FAMILY_GAUSS_CANON, value_dim=2)
    engine = C.InferenceEngine(model_engine=g, dependency_resolver=C.DefaultDependencyResolver(), inference_request_processor=processor,
                               dtype=dtype, api=api)
    if q is not None or r is not None:
        fids, vals = [], []
        for b in range(B):
            if r is not None:
                fids += liks[b]
                vals += [float(r[b])] * lengths[b]
            if q is not None:
                fids += trs[b]
                vals += [float(q[b])] * len(trs[b])
        fa = np.ascontiguousarray(fids, dtype=np.int64)
        va = np.ascontiguousarray(vals, dtype=np.float64)
        engine.store.check(api.set_factor_params(engine.store.h, len(fa), fa.ctypes.data_as(cap.i64p), va.ctypes.data_as(cap.f64p)))
    return engine, xs, ys, liks, trs
