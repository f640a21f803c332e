"""User-defined rules without rebuilding the library (CXB_RULE_PROGRAM): a small stack program per factor type, evaluated by
the device rule kernels (every schedule: level kernels, resident loop, sequential executor, memoised replay) and by the oracle's
restatement of the same machine. VERDICT r1, missing #7: "no user-supplied rule can reach the device"."""
import numpy as np
import pytest

from tests import models
from tests._pkg import pkg as C

cap = C.capi


def _random_walk_program(q):
    p = C.RuleProgram(default_param=q)
    p.const(1.0).param().dep(0, 0).mul().add().tset(0)  # den = 1 + q * L
    p.dep(0, 0).tget(0).div().store(0).dep(0, 1).tget(0).div().store(1)
    return p.rule()


def _observation_program(r):
    p = C.RuleProgram(default_param=r)
    p.const(1.0).param().div().store(0).dep(0, 0).param().div().store(1)  # (1 / r, y / r)
    return p.rule()


def _chain(api, T, rules, dtype=cap.F64):
    proc = C.RuleProcessor(rules, family=cap.FAMILY_GAUSS_CANON, value_dim=2)
    return models.make_ssm_model(T, api, processor=proc, dtype=dtype)


def test_programs_equal_the_built_in_rules_on_the_oracle(oracle_api):
    T = 30
    data = np.cumsum(np.random.Generator(np.random.PCG64(1)).standard_normal(T))
    a = _chain(oracle_api, T, {"likelihood": (cap.RULE_GAUSS_OBS, [1.3]), "transition": (cap.RULE_GAUSS_RW, [0.7])})
    b = _chain(oracle_api, T, {"likelihood": _observation_program(1.3), "transition": _random_walk_program(0.7)})
    for (e, x, y, lik, tr) in (a, b):
        models.ssm_set_data(e, y, lik, data)
        C.update_marginals(e, x, schedule="seq")
    sa, va = models.engine_state(a[0])
    sb, vb = models.engine_state(b[0])
    assert sa == sb
    np.testing.assert_array_equal(va, vb)  # the same operations in the same order


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [cap.F64, cap.F32])
@pytest.mark.parametrize("schedule,resident", [("lvl", "1"), ("lvl", "0"), ("seq", "1"), ("auto", "1")])
def test_programs_run_on_the_device(oracle_api, device_api, monkeypatch, dtype, schedule, resident):
    monkeypatch.setenv("CXB_ENGINE_RESIDENT", resident)
    T = 25
    rng = np.random.Generator(np.random.PCG64(2))
    rules = {"likelihood": _observation_program(1.3), "transition": _random_walk_program(0.7)}
    eo = _chain(oracle_api, T, rules)
    ed = _chain(device_api, T, rules, dtype=dtype)
    # per-factor parameters override the program's default on both sides
    for (e, x, y, lik, tr) in (eo, ed):
        fids = np.ascontiguousarray(tr[:5], dtype=np.int64)
        vals = np.ascontiguousarray([0.2, 0.4, 0.6, 0.8, 1.0])
        e.store.check(e.api.set_factor_params(e.store.h, 5, fids.ctypes.data_as(cap.i64p), vals.ctypes.data_as(cap.f64p)))
    for rep in range(4):  # with "auto" the later requests are memoised replays
        data = rng.standard_normal(T).astype(np.float32).astype(np.float64)
        for (e, x, y, lik, tr) in (eo, ed):
            models.ssm_set_data(e, y, lik, data)
        C.update_marginals(eo[0], eo[1], schedule="seq")
        C.update_marginals(ed[0], ed[1], schedule=schedule)
    so, vo = models.engine_state(eo[0])
    sd, vd = models.engine_state(ed[0])
    assert so == sd
    models.assert_values_close(vd, vo, dtype, kind="canon")


def test_a_new_factor_type_and_program_errors(backend):
    """out = a * x + b on scalars: a rule the library does not ship; and a malformed program is a NoRuleError."""
    g = C.BipartiteFactorGraph()
    v = g.add_variable(C.Variable(name="v"))
    w = g.add_variable(C.Variable(name="w"))
    f = g.add_factor(C.Factor(functional_form="affine"))
    g.add_edge(v, f, C.Connection(label="out"))
    g.add_edge(w, f, C.Connection(label="in"))
    prog = C.RuleProgram(default_param=3.0)
    prog.param().dep(0, 0).mul().const(0.5).add().store(0)
    e = C.InferenceEngine(model_engine=g, inference_request_processor=C.RuleProcessor({"affine": prog.rule()}, family=cap.FAMILY_SUM, value_dim=1),
                          api=backend)
    C.set_value(C.get_connection_message_to_factor(e, w, f), 4.0)
    C.update_marginals(e, v)
    assert C.get_value(C.get_variable_marginal(C.get_variable(e, v))) == 3.0 * 4.0 + 0.5
    bad = C.RuleProgram()
    bad.add().store(0)  # pops from an empty stack
    e2 = C.InferenceEngine(model_engine=_copy_graph(), inference_request_processor=C.RuleProcessor(
        {"affine": bad.rule()}, family=cap.FAMILY_SUM, value_dim=1), api=backend)
    C.set_value(C.get_connection_message_to_factor(e2, 1, 2), 4.0)
    with pytest.raises(C.NoRuleError):
        C.update_marginals(e2, 0)


def _copy_graph():
    g = C.BipartiteFactorGraph()
    v = g.add_variable(C.Variable(name="v"))
    w = g.add_variable(C.Variable(name="w"))
    f = g.add_factor(C.Factor(functional_form="affine"))
    g.add_edge(v, f, C.Connection(label="out"))
    g.add_edge(w, f, C.Connection(label="in"))
    return g
