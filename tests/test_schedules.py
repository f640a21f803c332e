"""The level-synchronous schedule (what the device runs) must equal the reference's
sequential in-place schedule (src/inference_engine.jl:559-632) on every benchmark graph family:
same executed multiset, bit-identical values, same observable end state."""
from collections import Counter

import numpy as np
import pytest

from tests._pkg import pkg
from tests import models

C = pkg
cap = pkg.capi


def _trace(engine):
    n = engine.api.trace_get(engine.store.h, None, None, 0)
    lv = np.zeros(max(n, 1), dtype=np.int64)
    sg = np.zeros(max(n, 1), dtype=np.int64)
    engine.api.trace_get(engine.store.h, lv.ctypes.data_as(cap.i64p), sg.ctypes.data_as(cap.i64p), n)
    return lv[:n], sg[:n]


def _same_state(e1, e2):
    s1, v1 = models.engine_state(e1)
    s2, v2 = models.engine_state(e2)
    assert s1 == s2
    assert np.array_equal(v1, v2)


def test_chain_seq_equals_lvl(oracle_api):
    T = 200
    data = np.cumsum(np.random.Generator(np.random.PCG64(1234)).standard_normal(T))
    engines = []
    for sched in ("seq", "lvl"):
        e, x, y, lik, tr = models.make_ssm_model(T, oracle_api, form="canon", q=0.7, r=1.3)
        models.ssm_set_data(e, y, lik, data)
        st = C.update_marginals(e, x, schedule=sched)
        assert st.updates == 6 * T - 4
        engines.append((e, st, _trace(e)))
    (e1, st1, (lv1, sg1)), (e2, st2, (lv2, sg2)) = engines
    assert Counter(sg1.tolist()) == Counter(sg2.tolist())
    assert st1.levels == 2  # two productive rounds
    assert st2.levels == 2 * T - 1
    widths = Counter(lv2[lv2 >= 0].tolist())
    assert widths[0] == T and all(widths[k] == 2 for k in range(1, 2 * T - 1))
    _same_state(e1, e2)
    # a second request on the already-updated model does nothing in either schedule
    for e, *_ in engines:
        pass
    assert C.update_marginals(e1, list(range(T)), schedule="seq").updates == 0
    assert C.update_marginals(e2, list(range(T)), schedule="lvl").updates == 0


def _sync_bp_reference(H, W, K, unary, psi, sweeps):
    """Independent dense numpy flooding BP on a grid (no shared code with the oracle)."""
    # messages m[(i,j) -> (k,l)] initialised uniform
    nbrs = lambda i, j: [(a, b) for a, b in ((i - 1, j), (i + 1, j), (i, j - 1), (i, j + 1)) if 0 <= a < H and 0 <= b < W]
    v2f = {((i, j), n): np.full(K, 1.0 / K) for i in range(H) for j in range(W) for n in nbrs(i, j)}
    marg = np.zeros((H, W, K))
    for _ in range(sweeps):
        f2v = {}
        for (src, dst), m in v2f.items():  # factor (src,dst) sends to dst
            out = psi.T @ m if src < dst else psi @ m
            f2v[(src, dst)] = out / out.sum()
        new = {}
        for i in range(H):
            for j in range(W):
                inc = {n: f2v[(n, (i, j))] for n in nbrs(i, j)}
                b = unary[i, j].copy()
                for n in sorted(inc):
                    b = b * inc[n]
                marg[i, j] = b / b.sum()
                for n in inc:
                    o = unary[i, j].copy()
                    for n2 in sorted(inc):
                        if n2 != n:
                            o = o * inc[n2]
                    new[((i, j), n)] = o / o.sum()
        v2f = new
    return marg


@pytest.mark.parametrize("rule", ["potts", "table"])
def test_grid_protocol_b_seq_equals_lvl_equals_flooding(oracle_api, rule):
    H, W, K, beta, sweeps = 4, 5, 3, 0.7, 5
    rng = np.random.Generator(np.random.PCG64(1234))
    unary = rng.dirichlet(np.ones(K), size=(H, W))
    res = []
    for sched in ("seq", "lvl"):
        e, pix, un, pair = models.make_grid_model(H, W, K, beta, oracle_api, rule=rule)
        vids = [v for row in pix for v in row]
        models.protocol_b_init(e, vids, K)
        usig = [C.get_connection_message_to_variable(e, pix[i][j], un[i][j]) for i in range(H) for j in range(W)]
        for s in range(sweeps):
            st = models.protocol_b_sweep(e, vids, usig, unary.reshape(-1, K), schedule=sched)
            n_pair_conn = 2 * len(pair)
            assert st.updates == 2 * n_pair_conn + H * W  # m2v + m2f + marginals, every sweep (Appendix B)
            assert st.final_marginals == H * W and st.final_linked == n_pair_conn
            if sched == "lvl":
                assert st.levels == 1
        res.append(e)
    _same_state(res[0], res[1])
    got = C.get_values([C.get_variable_marginal(C.get_variable(res[1], v)) for v in vids]).reshape(H, W, K)
    want = _sync_bp_reference(H, W, K, unary, models.potts_table(K, beta), sweeps)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)


def test_powerlaw_with_segment_trees_seq_equals_lvl(oracle_api):
    n, m, K, sweeps = 60, 150, 4, 4
    res, stats = [], []
    for sched in ("seq", "lvl"):
        e, vs, un, pair, unary, tables, ttype = models.make_powerlaw_model(n, m, K, oracle_api)
        models.protocol_b_init(e, vs, K)
        usig = [C.get_connection_message_to_variable(e, vs[i], un[i]) for i in range(n)]
        for s in range(sweeps):
            st = models.protocol_b_sweep(e, vs, usig, unary, schedule=sched)
        res.append(e)
        stats.append((st, _trace(e)))
    degs = [len(C.get_connected_factor_ids(res[0], v)) for v in vs]
    assert max(degs) > 5  # hubs use the segment tree
    n_prod = sum(d - 2 for d in degs if d > 5)
    (st1, (lv1, sg1)), (st2, (lv2, sg2)) = stats
    assert st1.updates == st2.updates == 4 * m + n + n_prod
    assert Counter(sg1.tolist()) == Counter(sg2.tolist())
    assert max(Counter(sg2.tolist()).values()) == 1  # every signal executed exactly once per sweep
    assert st2.levels > 1  # one level per product-tree height
    _same_state(res[0], res[1])


def test_naive_loopy_protocol_is_out_of_contract_or_equal(oracle_api):
    """Appendix B 'naive protocol': call 2 is Gauss-Seidel in the reference. The level-synchronous
    schedule must either reproduce it exactly or refuse (CXB_ERR_OUT_OF_CONTRACT) — never silently differ."""
    H, W, K = 3, 3, 3
    rng = np.random.Generator(np.random.PCG64(5))
    unary = rng.dirichlet(np.ones(K), size=(H, W))
    engines = []
    for sched in ("seq", "lvl"):
        e, pix, un, pair = models.make_grid_model(H, W, K, 0.5, oracle_api, link=False)
        vids = [v for row in pix for v in row]
        models.protocol_b_init(e, vids, K)
        usig = [C.get_connection_message_to_variable(e, pix[i][j], un[i][j]) for i in range(H) for j in range(W)]
        C.set_values(usig, unary.reshape(-1, K))
        C.update_marginals(e, vids, schedule=sched)
        engines.append((e, vids))
    _same_state(engines[0][0], engines[1][0])  # call 1 is synchronous in both
    e_seq, vids = engines[0]
    e_lvl, _ = engines[1]
    C.update_marginals(e_seq, vids, schedule="seq")
    try:
        C.update_marginals(e_lvl, vids, schedule="lvl")
    except C.OutOfContractError:
        return
    _same_state(e_seq, e_lvl)


# ---- random scripts: what is guaranteed, and the one known gap ----------------------------------------------------------------
def _chain_state(e):
    st, vals = models.engine_state(e)
    return st, [None if not c else tuple(np.round(v, 12)) for (c, _, _), v in zip(st, vals)]


@pytest.mark.parametrize("seed", range(25))
def test_random_scripts_of_full_updates_and_full_requests(oracle_api, seed):
    """Every step either sets ALL observations (new values) or requests ALL variables, in random order and number: the
    level-synchronous schedule leaves exactly the sequential schedule's state (values, pending flags, nibbles)."""
    rng = np.random.Generator(np.random.PCG64(777 + seed))
    T = int(rng.integers(3, 10))
    eng = [models.make_ssm_model(T, oracle_api, form="canon") for _ in range(2)]
    for _ in range(12):
        if rng.random() < 0.5:
            vals = np.stack([rng.standard_normal(T), np.zeros(T)], axis=1)
            for (e, x, y, lik, tr) in eng:
                C.set_values([C.get_connection_message_to_factor(e, y[i], lik[i]) for i in range(T)], vals)
        else:
            for (e, x, y, lik, tr), schedule in zip(eng, ("lvl", "seq")):
                C.update_marginals(e, x, schedule=schedule)
            assert _chain_state(eng[0][0]) == _chain_state(eng[1][0])


def test_incremental_evidence_with_leftover_freshness_is_refused(backend):
    """The smallest script on which a level schedule WITHOUT the request-time check answered differently from the
    reference (found by tests/fuzz_schedules.py): the first request runs while y_0 has no value, so marg(x_0) and
    marg(x_1) cannot be computed and keep FRESH bits on the backward messages; after all observations are set, the
    reference finds those marginals pending as soon as their other dependencies arrive and computes them from the STALE
    backward messages (marg(x_1) = (2.0, 5.5)), whereas advancing all variables at once recomputes the backward messages
    first ((2.0, 6.0)). The answer depends on the visiting order, so the request is refused - by the oracle's level
    schedule and by the device alike - before anything is computed; the sequential oracle shows the reference's answer."""
    T = 3
    e, x, y, lik, tr = models.make_ssm_model(T, backend, form="canon")
    sig = [C.get_connection_message_to_factor(e, y[i], lik[i]) for i in range(T)]
    C.set_values(sig[1:], np.array([[2.0, 0.0], [2.0, 0.0]]))
    C.update_marginals(e, x, schedule="lvl")  # incomplete: y_0 is missing
    C.set_values(sig, np.array([[3.0, 0.0]] * 3))
    before = models.engine_state(e)
    with pytest.raises(C.OutOfContractError, match="leftover freshness"):
        C.update_marginals(e, x, schedule="lvl")
    after = models.engine_state(e)
    assert before[0] == after[0] and np.array_equal(before[1], after[1], equal_nan=True)  # a refusal changes nothing
    # the default schedule answers as the reference does: on the device the refused request is rolled back and run by the
    # sequential executor; the stale backward message shows in x_1 (canonical form: precision, precision * mean)
    C.update_marginals(e, x)
    assert C.last_schedule(e) == C.capi.SCHEDULE_SEQUENTIAL
    got = C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in x])
    np.testing.assert_allclose(got[1], [2.0, 5.5])
    e2, x2, y2, lik2, _ = models.make_ssm_model(T, backend, form="canon")
    sig2 = [C.get_connection_message_to_factor(e2, y2[i], lik2[i]) for i in range(T)]
    C.set_values(sig2[1:], np.array([[2.0, 0.0], [2.0, 0.0]]))
    C.update_marginals(e2, x2, schedule="seq")
    C.set_values(sig2, np.array([[3.0, 0.0]] * 3))
    C.update_marginals(e2, x2, schedule="seq")
    got2 = C.get_values([C.get_variable_marginal(C.get_variable(e2, v)) for v in x2])
    np.testing.assert_allclose(got2[1], [2.0, 5.5])
    a, b = models.engine_state(e), models.engine_state(e2)
    assert a[0] == b[0] and np.array_equal(a[1], b[1], equal_nan=True)


@pytest.mark.parametrize("seed", range(20))
def test_random_scripts_of_incremental_evidence_are_sequential_or_refused(oracle_api, seed):
    """Random scripts that set SOME observations and request SOME variables: every request the level schedule accepts leaves
    the sequential schedule's state; the others are refused."""
    rng = np.random.Generator(np.random.PCG64(777 + seed))
    T = int(rng.integers(3, 10))
    eng = [models.make_ssm_model(T, oracle_api, form="canon") for _ in range(2)]
    for _ in range(12):
        ids = [int(i) for i in rng.choice(T, size=int(rng.integers(1, T + 1)), replace=False)]
        if rng.random() < 0.5:
            vals = np.stack([rng.standard_normal(len(ids)), np.zeros(len(ids))], axis=1)
            for (e, x, y, lik, tr) in eng:
                C.set_values([C.get_connection_message_to_factor(e, y[i], lik[i]) for i in ids], vals)
        else:
            try:
                C.update_marginals(eng[0][0], [eng[0][1][i] for i in ids], schedule="lvl")
            except C.OutOfContractError:
                return
            C.update_marginals(eng[1][0], [eng[1][1][i] for i in ids], schedule="seq")
            assert _chain_state(eng[0][0]) == _chain_state(eng[1][0])


# ---- hand-wired signal DAGs (outside the default BP wiring): the oracle's STRICT level schedule --------------------
# On by default on the oracle and on the device (CXO_STRICT_FRESHNESS=0 / CXB_STRICT=0, read when an engine is created, give
# the round-1 rules). They add to the level schedule:
#   (A) a signal the first traversal visits that is not pending but FRESH on a strong, computed, non-input dependency,
#   (B) a requested marginal that was already pending when the request arrived and has pending work beneath it,
#   (D) a frontier member found pending more than once, some visit through an intermediate slot, while one of its
#       dependencies is pending,
#   (E) a NON-listening notification that leaves a signal with complete criteria at the end of a level, or a signal that
#       becomes a frontier member after a non-listening notification of the same request
# -> refused. Scripts: tests/fuzz_schedules.py (random DAGs, strong listening dependencies).
_STRICT_SEEDS_THAT_DIFFERED = [13, 147, 171, 180, 188]  # of the first 300 fuzzer seeds, before the strict rules


def _fuzz_script(api, seed, n_ops=20, p_weak=0.0, p_listen=1.0):
    from tests import fuzz_schedules as fz

    rng = np.random.Generator(np.random.PCG64(9000 + seed))
    n_var, n_fac = int(rng.integers(2, 8)), int(rng.integers(1, 8))
    dep_p = float(rng.uniform(0.3, 0.9))
    build_seed = int(rng.integers(1 << 30))
    eng = []
    for _ in range(2):
        e, vs, inputs = fz._build(api, np.random.Generator(np.random.PCG64(build_seed)), n_var, n_fac, dep_p, p_weak=p_weak,
                                  p_listen=p_listen)
        eng.append((e, vs))
    outcome = "equal"
    for op in fz._script(rng, n_var, inputs, n_ops):
        a = fz._run(eng[0][0], eng[0][1], op, "lvl")
        if a != "ok":
            outcome = a
            break
        b = fz._run(eng[1][0], eng[1][1], op, "seq")
        if b != "ok" or fz._state(eng[0][0]) != fz._state(eng[1][0]):
            outcome = "differs"
            break
    return outcome


@pytest.mark.parametrize("seed", _STRICT_SEEDS_THAT_DIFFERED)
def test_strict_level_schedule_refuses_the_scripts_that_differed(oracle_api, monkeypatch, seed):
    monkeypatch.setenv("CXO_STRICT_FRESHNESS", "0")
    assert _fuzz_script(oracle_api, seed) == "differs"  # the default level schedule accepts them and answers differently
    monkeypatch.setenv("CXO_STRICT_FRESHNESS", "1")
    assert _fuzz_script(oracle_api, seed) == "refused"


@pytest.mark.parametrize("seed", range(40))
def test_strict_level_schedule_on_random_dags_is_sequential_or_refused(oracle_api, monkeypatch, seed):
    monkeypatch.setenv("CXO_STRICT_FRESHNESS", "1")
    assert _fuzz_script(oracle_api, seed) in ("equal", "refused", "norule")


_NONLISTENING_SEEDS_THAT_DIFFERED = [21, 37, 46, 56, 150, 159, 185, 216, 229, 252, 254, 281]  # 10 % non-listening dependencies


@pytest.mark.parametrize("seed", _NONLISTENING_SEEDS_THAT_DIFFERED)
def test_strict_level_schedule_refuses_the_non_listening_scripts_that_differed(oracle_api, monkeypatch, seed):
    monkeypatch.setenv("CXO_STRICT_FRESHNESS", "0")
    assert _fuzz_script(oracle_api, seed, p_listen=0.9) == "differs"
    monkeypatch.setenv("CXO_STRICT_FRESHNESS", "1")
    assert _fuzz_script(oracle_api, seed, p_listen=0.9) == "refused"


@pytest.mark.parametrize("seed", range(300, 340))
def test_strict_level_schedule_with_non_listening_dependencies_is_sequential_or_refused(oracle_api, monkeypatch, seed):
    monkeypatch.setenv("CXO_STRICT_FRESHNESS", "1")
    assert _fuzz_script(oracle_api, seed, p_listen=0.9) in ("equal", "refused", "norule")


def test_strict_level_schedule_accepts_the_benchmark_protocols(oracle_api, monkeypatch):
    """The strict rules refuse nothing the benchmark families need: protocol-B sweeps on a power-law graph with segment
    trees and on a Potts grid, and the chain, run unchanged (and equal to the sequential schedule)."""
    monkeypatch.setenv("CXO_STRICT_FRESHNESS", "1")
    test_powerlaw_with_segment_trees_seq_equals_lvl(oracle_api)
    for sched_pair in (("seq", "lvl"),):
        engines = []
        for sched in sched_pair:
            e, x, y, lik, tr = models.make_ssm_model(12, oracle_api, form="canon")
            sig = [C.get_connection_message_to_factor(e, y[i], lik[i]) for i in range(12)]
            for k in range(3):
                C.set_values(sig, np.array([[1.0 + k, 0.5 * i] for i in range(12)]))
                C.update_marginals(e, x, schedule=sched)
            engines.append(e)
        _same_state(engines[0], engines[1])
