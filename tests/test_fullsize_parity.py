"""Parity at the BASELINE sizes ("the benchmarked shapes are only known not to crash").

  * configs[1]: ALL 65,536 x 1,024 chains, fp32 and fp64, every marginal against the oracle's dense chain routine
    (cxo_chains_reference: the rule arithmetic of the explicit-graph oracle in dependency order), all six message classes on
    a sample of chains;
  * configs[3]: the 8192^2, K=16 grid - one shard == two row shards bit for bit at full size, and the oracle (explicit Signal
    graph, sequential update_marginals!) on corner / edge / interior 32 x 32 windows after 2 sweeps (a pixel's messages after
    s sweeps depend on evidence within distance s, so the window padded by 2 is an exact embedded sub-problem);
  * configs[4]: the 10M-variable Chung-Lu graph - after one sweep everything at a variable depends on its own star only, so the
    oracle runs on the stars of the three largest hubs, a few mid-degree variables and leaves (segment trees included);
  * configs[2]: K=64 HMMs at T=1e5 - sampled chains against a dense fp64 forward-backward at the start, the middle and the END
    of the 1e5 steps; the K=64 and the tensor-core K=512 kernels against the ORACLE (explicit graph) at T=2,000 / T=24.
CXB_SKIP_FULLSIZE=1 skips the file (the tests need ~130 GB of device memory and a few minutes)."""
import os

import numpy as np
import pytest

from tests import models
from tests._pkg import pkg as C
from tests.test_device_parity import _hmm_numpy

cap = C.capi
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("CXB_SKIP_FULLSIZE") == "1", reason="CXB_SKIP_FULLSIZE=1")]


@pytest.mark.parametrize("dtype", [cap.F32, cap.F64])
def test_all_65536_chains_of_config_2_against_the_oracle(oracle_api, dtype):
    B, T = 65536, 1024
    npdt = np.float32 if dtype == cap.F32 else np.float64
    rng = np.random.Generator(np.random.PCG64(1234))
    q, r = rng.uniform(0.5, 2.0, B), rng.uniform(0.5, 2.0, B)
    y = np.empty((T, B), dtype=npdt)
    x = np.zeros(B)
    for t in range(T):  # fp64 master data cast to the engine dtype
        x = x + rng.standard_normal(B) * np.sqrt(q)
        y[t] = (x + rng.standard_normal(B) * np.sqrt(r)).astype(npdt)
    ch = C.GaussianChainBatch(B, T, dtype=dtype)
    ch.set_noise(q, r)
    ch.set_observations(y)
    assert ch.update_marginals() == B * (6 * T - 4)
    got_marg = ch.get_marginals().astype(np.float64)
    q_used = q.astype(npdt).astype(np.float64)  # the kernel holds the variances in the engine dtype
    r_used = r.astype(npdt).astype(np.float64)
    chunk = 8192
    for b0 in range(0, B, chunk):
        yc = np.ascontiguousarray(y[:, b0:b0 + chunk].astype(np.float64))
        ref = np.zeros((6, T, chunk, 2))
        qc, rc = np.ascontiguousarray(q_used[b0:b0 + chunk]), np.ascontiguousarray(r_used[b0:b0 + chunk])
        oracle_api.chains_reference(chunk, T, qc.ctypes.data_as(cap.f64p), rc.ctypes.data_as(cap.f64p), yc.ctypes.data_as(cap.f64p),
                                    ref.ctypes.data_as(cap.f64p))
        models.assert_values_close(got_marg[:, b0:b0 + chunk], ref[5], dtype, kind="canon", err_msg=f"marginals of chains {b0}..{b0 + chunk}")
        if b0 == 0:  # all six message classes on the first 8,192 chains
            for m in range(5):
                got = ch.get_messages(m)[:, :chunk].astype(np.float64)
                models.assert_values_close(got, ref[m], dtype, kind="canon", err_msg=C.GaussianChainBatch.MESSAGE_CLASSES[m])


def _oracle_window(oracle_api, unary, r0, r1, c0, c1, N, K, beta, sweeps, pad):
    """Marginals of rows [r0, r1) x cols [c0, c1) of the N x N grid after `sweeps` protocol-B sweeps, computed by the oracle on
    the window padded by `pad` pixels on every side that is not the grid's own border."""
    R0, R1, C0, C1 = max(0, r0 - pad), min(N, r1 + pad), max(0, c0 - pad), min(N, c1 + pad)
    H, W = R1 - R0, C1 - C0
    e, pix, un, pair = models.make_grid_model(H, W, K, beta, oracle_api, rule="potts", link=True)
    vs = [v for row in pix for v in row]
    models.protocol_b_init(e, vs, K)
    usig = [C.get_connection_message_to_variable(e, pix[i][j], un[i][j]) for i in range(H) for j in range(W)]
    u = unary[R0:R1, C0:C1].astype(np.float64).reshape(-1, K)
    for _ in range(sweeps):
        models.protocol_b_sweep(e, vs, usig, u, schedule="seq")
    marg = C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in vs]).reshape(H, W, K)
    return marg[r0 - R0:r1 - R0, c0 - C0:c1 - C0]


def test_potts_grid_of_config_4_at_full_size(oracle_api):
    N, K, beta, sweeps = 8192, 16, 0.7, 2
    rng = np.random.Generator(np.random.PCG64(1234))
    unary = np.empty((N, N, K), dtype=np.float32)
    for i in range(0, N, 256):  # Dirichlet(1) rows = normalised exponentials
        e = rng.standard_exponential((256, N, K), dtype=np.float32)
        unary[i:i + 256] = e / e.sum(axis=-1, keepdims=True)
    full = C.PottsGrid(N, N, K, beta, dtype=cap.F32)
    full.set_unary(unary)
    full.reset_messages()
    n_upd = 0
    for _ in range(sweeps):
        n_upd = full.sweep()
    assert n_upd == 603914240  # 268,402,688 m2v + 268,402,688 m2f + 67,108,864 marginals
    marg_full = full.get_marginals()
    del full
    # (1) probabilities: every marginal is a normalised vector (a checksum over all 67M pixels)
    sums = marg_full.sum(axis=-1, dtype=np.float64)
    assert np.all(np.abs(sums - 1.0) < 1e-5) and np.all(marg_full >= 0)
    # (2) the oracle on corner / edge / interior windows
    for (r0, c0) in ((0, 0), (0, 4000), (N - 32, N - 32), (4096 - 16, 0), (3000, 5000), (4096 - 16, 4096 - 16)):
        want = _oracle_window(oracle_api, unary, r0, r0 + 32, c0, c0 + 32, N, K, beta, sweeps, pad=2)
        models.assert_values_close(marg_full[r0:r0 + 32, c0:c0 + 32], want, cap.F32, kind="prob", err_msg=f"window at ({r0}, {c0})")
    # (3) shard-count invariance at full size: two row shards with the fused peer halo == one shard, bit for bit
    cut = N // 2
    shards = [C.PottsGrid(cut, N, K, beta, dtype=cap.F32, has_upper=i > 0, has_lower=i < 1) for i in range(2)]
    for i, sh in enumerate(shards):
        sh.set_unary(unary[i * cut:(i + 1) * cut])
        sh.reset_messages()
    shards[0].p2p_connect_local(1, shards[1])
    shards[1].p2p_connect_local(0, shards[0])
    for _ in range(sweeps):
        for sh in shards:
            sh.sweep()
    for i, sh in enumerate(shards):
        assert np.array_equal(sh.get_marginals(), marg_full[i * cut:(i + 1) * cut]), f"shard {i}"


def _oracle_star(oracle_api, v, nbr_u, nbr_t, K, tables, unary):
    """One protocol-B sweep of the oracle on the star of variable v: v, its neighbours (in id order), one unary factor per
    variable, the pairwise factors of v in their original order. Returns marginal(v), m2v(v, f_j), m2f(v, f_j)."""
    import ctypes

    d = len(nbr_u)
    ids = sorted(set([v] + [int(u) for u in nbr_u]))
    assert len(ids) == d + 1
    new = {o: i for i, o in enumerate(ids)}
    nv = d + 1
    n_ids = 2 * nv + d
    is_factor = np.zeros(n_ids, dtype=np.uint8)
    is_factor[nv:] = 1
    n_tables = tables.shape[0]
    ftype = np.zeros(n_ids, dtype=np.int32)
    ftype[nv:2 * nv] = n_tables
    ftype[2 * nv:] = nbr_t
    lo = np.array([new[min(v, int(u))] for u in nbr_u], dtype=np.int64)
    hi = np.array([new[max(v, int(u))] for u in nbr_u], dtype=np.int64)
    ev = np.concatenate([np.arange(nv), np.stack([lo, hi], axis=1).ravel()]).astype(np.int64)
    ef = np.concatenate([nv + np.arange(nv), np.repeat(2 * nv + np.arange(d), 2)]).astype(np.int64)
    st = C.SignalStore(oracle_api, K, cap.FAMILY_CATEGORICAL, cap.F64, 0)
    api, h = oracle_api, st.h
    st.check(api.graph_build(h, n_ids, is_factor.ctypes.data_as(cap.u8p), ftype.ctypes.data_as(cap.i32p), len(ev), ev.ctypes.data_as(cap.i64p),
                             ef.ctypes.data_as(cap.i64p)))
    for t in range(n_tables):
        tb = np.ascontiguousarray(tables[t].astype(np.float64).ravel())
        st.check(api.register_rule(h, t, cap.RULE_CAT_TABLE, tb.ctypes.data_as(cap.f64p), tb.size))
    st.check(api.resolve_dependencies(h, cap.RESOLVER_DEFAULT_BP))
    pair_conn = nv + np.arange(2 * d)
    pair_m2f = np.ascontiguousarray(nv + 2 * pair_conn + 1, dtype=np.int64)
    lv = np.ascontiguousarray(ev[nv:], dtype=np.int64)
    st.check(api.link_signals(h, 2 * d, lv.ctypes.data_as(cap.i64p), pair_m2f.ctypes.data_as(cap.i64p)))
    init = np.full((2 * d, K), 1.0 / K)
    st.check(api.set_values(h, 2 * d, pair_m2f.ctypes.data_as(cap.i64p), init.ctypes.data_as(cap.f64p), K))
    usig = np.ascontiguousarray(nv + 2 * np.arange(nv), dtype=np.int64)
    u = np.ascontiguousarray(unary[ids].astype(np.float64))
    st.check(api.set_values(h, nv, usig.ctypes.data_as(cap.i64p), u.ctypes.data_as(cap.f64p), K))
    xs = np.arange(nv, dtype=np.int64)
    stats = cap.UpdateStats()
    st.check(api.set_schedule(h, cap.SCHEDULE_SEQUENTIAL))
    st.check(api.update_marginals(h, nv, xs.ctypes.data_as(cap.i64p), ctypes.byref(stats)))
    side = np.array([0 if v < int(uu) else 1 for uu in nbr_u])
    conn_v = nv + 2 * np.arange(d) + side  # connection index of (v, f_j)
    sig = np.ascontiguousarray(np.concatenate([[new[v]], nv + 2 * conn_v, nv + 2 * conn_v + 1]), dtype=np.int64)
    out = np.zeros((len(sig), K))
    st.check(api.get_values(h, len(sig), sig.ctypes.data_as(cap.i64p), out.ctypes.data_as(cap.f64p), K))
    return out[0], out[1:1 + d], out[1 + d:]


def test_powerlaw_graph_of_config_5_at_full_size(oracle_api):
    n, K, n_tables = 10_000_000, 8, 16
    m = 2 * n
    rng = np.random.Generator(np.random.PCG64(1235))
    edges = np.asarray(models.chung_lu_edges_fast(n, m), dtype=np.int64)
    ttype = rng.integers(0, n_tables, size=m).astype(np.int32)
    tables = np.exp(rng.standard_normal((n_tables, K, K))).astype(np.float32).astype(np.float64)
    e = rng.standard_exponential((n, K), dtype=np.float32)
    unary = e / e.sum(axis=-1, keepdims=True)
    pw = C.PairwiseGraph(n, edges[:, 0], edges[:, 1], ttype, tables, dtype=cap.F32)
    pw.set_unary(unary)
    pw.reset_messages()
    pw.sweep()
    marg = pw.get_marginals()
    m2v, m2f = pw.get_messages(0), pw.get_messages(1)
    sums = marg.sum(axis=-1, dtype=np.float64)
    assert np.all(np.abs(sums - 1.0) < 1e-5) and np.all(marg >= 0)
    deg = np.bincount(edges.ravel(), minlength=n)
    order = np.argsort(edges.ravel(), kind="stable")  # incidences grouped by variable, ascending factor inside a variable
    start = np.concatenate([[0], np.cumsum(deg)])
    hubs = list(np.argsort(-deg)[:3])
    mid = [int(v) for v in np.flatnonzero((deg >= 6) & (deg <= 40))[[5, 5000, 50000]]]
    leaves = [int(v) for v in np.flatnonzero((deg >= 1) & (deg <= 4))[[7, 70000, 700000]]]
    for v in [int(h) for h in hubs] + mid + leaves:
        inc = order[start[v]:start[v + 1]]  # indices into edges.ravel(): factor = inc // 2, side = inc % 2
        f = inc // 2
        nbr_u = edges[f, 1 - (inc % 2)]
        want_marg, want_m2v, want_m2f = _oracle_star(oracle_api, v, nbr_u, ttype[f], K, tables, unary)
        side = inc % 2  # get_messages: [factor][endpoint side][K]
        models.assert_values_close(marg[v], want_marg, cap.F32, kind="prob", err_msg=f"marginal of variable {v} (degree {deg[v]})")
        models.assert_values_close(m2v[f, side], want_m2v, cap.F32, kind="prob", err_msg=f"m2v of variable {v} (degree {deg[v]})")
        models.assert_values_close(m2f[f, side], want_m2f, cap.F32, kind="prob", err_msg=f"m2f of variable {v} (degree {deg[v]})")


def test_hmm_k64_of_config_3_at_t_1e5():
    B, T, K, M = 1024, 100000, 64, 32
    rng = np.random.Generator(np.random.PCG64(1234))
    A = rng.dirichlet(np.ones(K), size=K).astype(np.float32).astype(np.float64)
    E = (rng.dirichlet(np.ones(K), size=M).T * K).astype(np.float32).astype(np.float64)
    obs = rng.integers(0, M, size=(T, B), dtype=np.uint8)
    hm = C.HmmBatch(B, T, K, M, dtype=cap.F32)
    hm.set_tables(A, E)
    hm.set_observations(obs)
    assert hm.update_marginals() == B * (6 * T - 4)
    windows = [(0, 400), (T // 2 - 200, T // 2 + 200), (T - 400, T)]
    got = [hm.get_marginals(t0, t1) for (t0, t1) in windows]
    for b in (0, 517, B - 1):
        _, want = _hmm_numpy(A, E, obs[:, b])  # dense fp64 scaled forward-backward (tests/test_device_parity.py)
        for (t0, t1), g in zip(windows, got):
            models.assert_values_close(g[:, b, :], want[t0:t1], cap.F32, kind="prob", err_msg=f"chain {b}, steps {t0}..{t1}")


@pytest.mark.parametrize("K,T,B", [(64, 2000, 3), (512, 24, 130)])
def test_hmm_kernels_against_the_oracle_explicit_graph(oracle_api, K, T, B):
    """The structured HMM kernels (K = 64: FFMA; K = 512: tcgen05 + TMEM, 3-piece bf16 operands) against the ORACLE running
    update_marginals! on the explicit graph (the SURVEY 8 parity chain), not only against numpy."""
    M = 16
    rng = np.random.Generator(np.random.PCG64(77))
    A = rng.dirichlet(np.ones(K) * 0.5, size=K).astype(np.float32).astype(np.float64)
    E = (rng.dirichlet(np.ones(K), size=M).T * K).astype(np.float32).astype(np.float64)
    obs = rng.integers(0, M, size=(T, B), dtype=np.uint8)
    hm = C.HmmBatch(B, T, K, M, dtype=cap.F32)
    hm.set_tables(A, E)
    hm.set_observations(obs)
    hm.update_marginals()
    got = hm.get_marginals()
    for b in (0, B - 1):
        e, z, y, prior, em, tr = models.make_hmm_model(T, K, M, A, E, oracle_api)
        models.hmm_set_data(e, z, y, prior, em, obs[:, b], K)
        C.update_marginals(e, z, schedule="seq")
        want = C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in z])
        models.assert_values_close(got[:, b, :], want, cap.F32, kind="prob", err_msg=f"K={K}, chain {b}")
