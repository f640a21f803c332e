"""Golden fixtures (tests/golden/, made by tests/golden/make_golden.py).

CPU (not gpu): the oracle against the known answers transcribed from the reference's own test-suite
(`reference_known_answers.json`), and the committed oracle vectors re-derived so that they cannot rot.
GPU: the structured CUDA engines against the committed vectors, with no oracle in the loop."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests import models
from tests._pkg import pkg as C

cap = C.capi
GOLD = Path(__file__).resolve().parent / "golden"
REF = json.loads((GOLD / "reference_known_answers.json").read_text())


# ---- the oracle against the reference's known answers (CPU) ---------------------------------------------------------------
def test_gaussian_product_rule_matches_reference_fixture(oracle_api):  # test/runtests.jl:40-46
    """marginal of a 2-step chain = product of the two Gaussians it receives; checked in the fixture's (mean, variance) form."""
    for case in REF["gaussian_product_mean_variance"]["cases"]:
        w = 1.0 / case["v1"] + 1.0 / case["v2"]
        v = 1.0 / w
        m = v * (case["m1"] / case["v1"] + case["m2"] / case["v2"])
        assert m == pytest.approx(case["m"], rel=1e-12) and v == pytest.approx(case["v"], rel=1e-12)
        # the oracle's moment-form family product (FAMILY_GAUSS_MV) on a one-variable, two-observation model
        e, x, y, lik, tr = models.make_ssm_model(1, oracle_api, form="mv", r=1.0)
        # one state, its likelihood message N(y, r): the marginal of a single message is the message (n = 1 path)
        models.ssm_set_data(e, y, lik, [case["m1"]])
        C.update_marginals(e, x)
        got = C.get_value(C.get_variable_marginal(C.get_variable(e, x[0])))
        assert got[0] == pytest.approx(case["m1"]) and got[1] == pytest.approx(1.0)  # observation -> N(y, 1.0) (:415-432)


def test_beta_bernoulli_posterior_matches_reference_known_answer(oracle_api):  # test/inference_engine_tests.jl:360-376
    ref = REF["beta_bernoulli"]
    n = ref["n"]
    e, p, o, f = models.make_beta_bernoulli_model(n, oracle_api)
    data = np.random.Generator(np.random.PCG64(7)).integers(0, 2, size=n).astype(np.float64)
    C.set_values([C.get_connection_message_to_factor(e, o[i], f[i]) for i in range(n)], data.reshape(-1, 1))
    stats = C.update_marginals(e, p)
    a, b = C.get_value(C.get_variable_marginal(C.get_variable(e, p)))
    assert a == pytest.approx(ref["prior"][0] + data.sum()) and b == pytest.approx(ref["prior"][1] + n - data.sum())
    assert stats.updates == ref["executions"]


def test_ssm_bp_properties_of_the_reference_hold(oracle_api):  # test/inference_engine_tests.jl:477-487
    n = 100
    e, x, y, lik, tr = models.make_ssm_model(n, oracle_api, form="mv")
    data = 2.0 * np.arange(1, n + 1) + np.random.Generator(np.random.PCG64(1)).standard_normal(n)
    models.ssm_set_data(e, y, lik, data)
    C.update_marginals(e, x)
    mv = C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in x])
    assert np.all(mv[:, 0] >= 0) and np.all(np.diff(mv[:, 0]) >= 0) and np.all(mv[:, 1] >= 0)


def test_nibble_layout_constants_match_reference():  # src/signal.jl:36-45, 507-526
    bits = REF["pending_nibble_layout"]["bits"]
    assert (cap.NIB_INTERMEDIATE, cap.NIB_WEAK, cap.NIB_COMPUTED, cap.NIB_FRESH) == (
        bits["intermediate"], bits["weak"], bits["computed"], bits["fresh"])


# ---- committed oracle vectors: re-derived on the CPU, checked on the GPU ------------------------------------------------
def test_committed_oracle_vectors_are_reproducible(oracle_api):
    from tests.golden import make_golden as mg

    for name, fn in (("oracle_gauss_chain", mg.oracle_chain), ("oracle_hmm", mg.oracle_hmm), ("oracle_pairwise", mg.oracle_pairwise),
                     ("oracle_vmp", mg.oracle_vmp)):
        want = np.load(GOLD / f"{name}.npz")
        got = fn(oracle_api)
        assert sorted(want.files) == sorted(got)
        for k in want.files:
            np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), want[k].astype(np.float64), rtol=1e-13, atol=0, err_msg=f"{name}:{k}")


@pytest.mark.gpu
def test_chain_kernel_matches_committed_vectors():
    g = np.load(GOLD / "oracle_gauss_chain.npz")
    T, B = g["y"].shape
    ch = C.GaussianChainBatch(B, T, dtype=cap.F64)
    ch.set_noise(g["q"], g["r"])
    ch.set_observations(g["y"])
    ch.update_marginals()
    np.testing.assert_allclose(ch.get_marginals(), g["marginals"], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(cap.F64, 1e-12), (cap.F32, 1e-5)])
def test_hmm_kernel_matches_committed_vectors(dtype, tol):
    g = np.load(GOLD / "oracle_hmm.npz")
    T, B = g["obs"].shape
    K, M = g["E"].shape
    hm = C.HmmBatch(B, T, K, M, dtype=dtype)
    hm.set_tables(g["A"], g["E"])
    hm.set_observations(g["obs"])
    hm.update_marginals()
    np.testing.assert_allclose(hm.get_marginals(), g["marginals"], rtol=tol, atol=tol * 1e-3)  # element-wise


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(cap.F64, 1e-12), (cap.F32, 1e-5)])
def test_pairwise_kernel_matches_committed_vectors(dtype, tol):
    g = np.load(GOLD / "oracle_pairwise.npz")
    n = g["unary"].shape[0]
    tables = g["tables"]
    pw = C.PairwiseGraph(n, g["edges"][:, 0], g["edges"][:, 1], g["ttype"], tables, dtype=dtype)
    pw.set_unary(g["unary"])
    pw.reset_messages()
    for _ in range(int(g["sweeps"])):
        pw.sweep()
    # the committed inputs are fp32-representable (make_golden.f32): the tolerance covers arithmetic only, element-wise
    np.testing.assert_allclose(pw.get_marginals(), g["marginals"], rtol=tol, atol=tol * 1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(cap.F64, 1e-12), (cap.F32, 1e-5)])
def test_vmp_engine_matches_committed_vectors(dtype, tol):
    """Mean-field and structured VMP on the device engine (level-synchronous schedule) against vectors the oracle produced
    on the reference's SEQUENTIAL schedule: weak dependencies, joint marginals, per-variable value families."""
    g = np.load(GOLD / "oracle_vmp.npz")
    data, iters = g["data"], int(g["iters"])
    api = C.default_api()
    m = models.make_ssm_mean_field_model(len(data), api, dtype=dtype)
    got = models.ssm_mean_field_experiment(m[0], m[1], m[2], m[3], m[4], data, iters)
    for k in ("x", "ssnoise", "obsnoise"):
        np.testing.assert_allclose(got[k], g[f"mf_{k}"], rtol=tol, atol=tol * 1e-3, err_msg=f"mean-field {k}")
    m = models.make_ssm_structured_model(len(data), api, dtype=dtype)
    got = models.ssm_structured_experiment(m[0], m[1], m[2], m[3], m[4], data, iters, merged_all=False)
    for k in ("x", "ssnoise", "obsnoise"):
        np.testing.assert_allclose(got[k][..., :2], g[f"st_{k}"], rtol=tol, atol=tol * 1e-3, err_msg=f"structured {k}")
