"""Device schedules against the oracle on random HAND-WIRED signal DAGs (random weak / intermediate / non-listening
dependencies, random set_value! / update_marginals! scripts: tests/fuzz_schedules.py builds them).

  * AUTO (the default): the device answers every request exactly as the reference's sequential loop does (the oracle's
    `seq` schedule): same values, computed / pending flags and nibbles after every operation, nothing refused;
  * LEVEL: the device's level-synchronous schedule refuses exactly the requests the oracle's level schedule refuses
    (strict rules A / B / D / E and rule F included) and leaves the same state on the accepted ones; a refusal leaves the
    engine untouched;
  * the reproduced final-phase case of ADVICE.md (a linked signal that depends on a later-requested marginal).
"""
import numpy as np
import pytest

from tests import fuzz_schedules as fz
from tests import models
from tests._pkg import pkg as C

cap = C.capi
MIXES = {"strong": (0.0, 1.0), "nonlisten": (0.0, 0.9), "weak": (0.35, 1.0), "weak+nonlisten": (0.35, 0.9)}


def _pair(oracle_api, device_api, seed, mix, big=False):
    p_weak, p_listen = MIXES[mix]
    rng = np.random.Generator(np.random.PCG64(9000 + seed))
    hi = 8 if big else 6
    n_var, n_fac = int(rng.integers(2, hi)), int(rng.integers(1, hi))
    dep_p = float(rng.uniform(0.3, 0.9))
    build_seed = int(rng.integers(1 << 30))
    eo, vso, inputs = fz._build(oracle_api, np.random.Generator(np.random.PCG64(build_seed)), n_var, n_fac, dep_p, p_weak=p_weak, p_listen=p_listen)
    ed, vsd, _ = fz._build(device_api, np.random.Generator(np.random.PCG64(build_seed)), n_var, n_fac, dep_p, p_weak=p_weak, p_listen=p_listen)
    return rng, n_var, inputs, (eo, vso), (ed, vsd)


@pytest.mark.gpu
@pytest.mark.parametrize("mix", list(MIXES))
@pytest.mark.parametrize("seed", range(30))
def test_device_default_schedule_equals_the_sequential_reference(oracle_api, device_api, seed, mix):
    rng, n_var, inputs, (eo, vso), (ed, vsd) = _pair(oracle_api, device_api, seed, mix, big=True)
    n_req = 0
    for op in fz._script(rng, n_var, inputs, 16):
        r_o = fz._run(eo, vso, op, "seq")
        r_d = fz._run(ed, vsd, op, "auto")
        assert r_o == r_d and r_d != "refused", (seed, mix, op, r_o, r_d)
        if r_o != "ok":
            break  # a free signal without a rule: the reference throws midway as well
        assert fz._state(eo) == fz._state(ed), (seed, mix, op)
        n_req += op[0] == "update"
        if op[0] == "update":
            assert C.last_schedule(ed) == cap.SCHEDULE_SEQUENTIAL  # hand-wired graph: the literal loop on the device


@pytest.mark.gpu
@pytest.mark.parametrize("mix", list(MIXES))
@pytest.mark.parametrize("seed", range(30))
def test_device_level_schedule_equals_oracle_level_schedule(oracle_api, device_api, seed, mix):
    rng, n_var, inputs, (eo, vso), (ed, vsd) = _pair(oracle_api, device_api, seed, mix)
    for op in fz._script(rng, n_var, inputs, 14):
        before = fz._state(ed)
        r_o = fz._run(eo, vso, op, "lvl")
        r_d = fz._run(ed, vsd, op, "lvl")
        assert r_o == r_d, (seed, mix, op, r_o, r_d)  # the same requests are refused
        if r_d == "refused":
            assert fz._state(ed) == before, (seed, mix, op)  # a refusal is side-effect free on the device
        if r_o != "ok":
            break
        assert fz._state(eo) == fz._state(ed), (seed, mix, op)


def _linked_depends_on_later_marginal(api):
    """ADVICE.md (round 1, high): variables a < b; a signal L linked to a depends on marginal(b); request [a, b]."""
    g = C.BipartiteFactorGraph()
    a = g.add_variable(C.Variable(name="a"))
    b = g.add_variable(C.Variable(name="b"))
    fa = g.add_factor(C.Factor(functional_form="f"))
    fb = g.add_factor(C.Factor(functional_form="f"))
    g.add_edge(a, fa, C.Connection(label="e"))
    g.add_edge(b, fb, C.Connection(label="e"))
    proc = C.RuleProcessor({"f": (cap.RULE_SCALE2, [])}, family=cap.FAMILY_SUM, value_dim=1)
    e = C.InferenceEngine(model_engine=g, inference_request_processor=proc, api=api)
    L = C.create_inference_signal(e)
    C.set_variant(L, C.MessageToFactor(a, fa))  # any variant with a family-reduce rule
    C.add_dependency(L, C.get_variable_marginal(C.get_variable(e, b)))
    C.link_signal_to_variable(C.get_variable(e, a), L)
    return e, a, b, fa, fb, L


def _run_linked_case(api, schedule):
    e, a, b, fa, fb, L = _linked_depends_on_later_marginal(api)
    out = []
    for k in range(2):
        C.set_value(C.get_connection_message_to_variable(e, a, fa), 1.0 + k)
        C.set_value(C.get_connection_message_to_variable(e, b, fb), 5.0 * (k + 1))
        C.update_marginals(e, [a, b], schedule=schedule)
        out.append((C.is_computed(L), C.get_value(L) if C.is_computed(L) else None))
    return e, out


def test_linked_signal_that_depends_on_a_later_marginal_oracle(oracle_api):
    _, seq = _run_linked_case(oracle_api, "seq")
    assert seq == [(False, None), (True, 5.0)]  # the reference reaches L before marginal(b) is computed
    with pytest.raises(C.OutOfContractError, match="final-phase"):
        _run_linked_case(oracle_api, "lvl")


@pytest.mark.gpu
def test_linked_signal_that_depends_on_a_later_marginal_device(oracle_api, device_api):
    eo, seq = _run_linked_case(oracle_api, "seq")
    ed, got = _run_linked_case(device_api, "auto")
    assert got == seq
    assert models.engine_state(eo)[0] == models.engine_state(ed)[0]
    with pytest.raises(C.OutOfContractError, match="final-phase"):
        _run_linked_case(device_api, "lvl")
    # wired by the resolver only (no hand-made dependency), so that AUTO goes through the level schedule, is refused by rule F,
    # rolled back and answered by the sequential executor: the structured VMP model below covers that path


@pytest.mark.gpu
def test_reference_vmp_call_sequence_runs_on_the_device(oracle_api, device_api):
    """ADVICE.md (round 1, medium): the reference's structured-VMP experiment ends every iteration with
    update_marginals!(engine, [ssnoise, obsnoise, x...]) (test/inference_engine_tests.jl:1089-1113). The level schedule
    refuses that request; the default schedule runs it as the reference does."""
    n, iters = 12, 4
    data = models.ssm_mean_field_dataset(n, seed=5)
    mo = models.make_ssm_structured_model(n, oracle_api)
    md = models.make_ssm_structured_model(n, device_api)
    want = models.ssm_structured_experiment(mo[0], mo[1], mo[2], mo[3], mo[4], data, iters, schedule="seq", merged_all=True)
    got = models.ssm_structured_experiment(md[0], md[1], md[2], md[3], md[4], data, iters, schedule="auto", merged_all=True)
    for k in ("x", "ssnoise", "obsnoise"):
        np.testing.assert_allclose(got[k][..., :2], want[k][..., :2], rtol=1e-12, atol=0, err_msg=k)
    assert models.engine_state(md[0])[0] == models.engine_state(mo[0])[0]


@pytest.mark.gpu
def test_incremental_chain_scripts_default_schedule_equals_reference(oracle_api, device_api):
    """Resolver-built graph (AUTO = level schedule, rollback + sequential executor when refused): random scripts of
    incremental evidence on the chain; the device must leave the reference's state after every request."""
    fell_back = 0
    for seed in range(12):
        rng = np.random.Generator(np.random.PCG64(777 + seed))
        T = int(rng.integers(3, 10))
        eo = models.make_ssm_model(T, oracle_api, form="canon")
        ed = models.make_ssm_model(T, device_api, form="canon")
        for _ in range(12):
            ids = [int(i) for i in rng.choice(T, size=int(rng.integers(1, T + 1)), replace=False)]
            if rng.random() < 0.5:
                vals = np.stack([rng.standard_normal(len(ids)), np.zeros(len(ids))], axis=1)
                for (e, x, y, lik, tr) in (eo, ed):
                    C.set_values([C.get_connection_message_to_factor(e, y[i], lik[i]) for i in ids], vals)
            else:
                C.update_marginals(eo[0], [eo[1][i] for i in ids], schedule="seq")
                C.update_marginals(ed[0], [ed[1][i] for i in ids], schedule="auto")
                fell_back += C.last_schedule(ed[0]) == cap.SCHEDULE_SEQUENTIAL
                so, sd = models.engine_state(eo[0]), models.engine_state(ed[0])
                assert so[0] == sd[0], (seed, ids)
                np.testing.assert_allclose(sd[1], so[1], rtol=1e-12, atol=0, equal_nan=True)
    assert fell_back > 0  # some requests were refused by the level schedule and answered by the sequential executor


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(40)) + [112, 768, 949, 482])
def test_default_schedule_on_resolver_built_graphs_equals_the_reference(oracle_api, device_api, seed):
    """Graphs wired only by the default resolver (AUTO = level schedule, certified against the sequential executor when first
    recorded): random bipartite graphs with loops, leaves and hubs, messages overwritten by the user, random links, random
    request subsets and orders (tests/fuzz_bp_graphs.py). Seeds 112, 768, 949: scripts on which the oracle's level schedule
    ACCEPTS a request and answers differently from the reference; 482: the reference's own loop does not terminate."""
    from tests import fuzz_bp_graphs as fb

    assert fb.one_seed(device_api, oracle_api, seed, "auto", "seq") in ("equal", "diverges")
