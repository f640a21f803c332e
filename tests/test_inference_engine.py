"""Port of the reference's test/inference_engine_tests.jl and test/dependencies_tests.jl.
Every test runs against the oracle and (gpu-marked) the device engine through the same frontend.
Citations are line ranges of the reference test files."""
import numpy as np
import pytest

from tests._pkg import pkg
from tests.oracle_frontend import CallbackProcessor
from tests import models

C = pkg
cap = pkg.capi
G = C.BipartiteFactorGraph


def m2v(e, v, f):
    return C.get_connection_message_to_variable(e, v, f)


def m2f(e, v, f):
    return C.get_connection_message_to_factor(e, v, f)


def marginal(e, v):
    return C.get_variable_marginal(C.get_variable(e, v))


def test_warning_for_variable_without_factors(backend):  # inference_engine_tests.jl:33-46
    g = G()
    v = g.add_variable(C.Variable(name="v"))
    e = C.InferenceEngine(model_engine=g, api=backend)
    w = C.get_warnings(e)
    assert len(w) == 1 and w[0].description == "Variable has no connected factors" and w[0].context == v


def test_signals_metadata_prepared_by_default(backend):  # :48-91
    g = G()
    v1, v2, v3 = (g.add_variable(C.Variable(name=n)) for n in ("v1", "v2", "v3"))
    f1, f2 = (g.add_factor(C.Factor(functional_form=n)) for n in ("f1", "f2"))
    g.add_edge(v1, f1, C.Connection(label="out"))
    g.add_edge(v2, f2, C.Connection(label="out"))
    g.add_edge(v3, f1, C.Connection(label="in"))
    g.add_edge(v3, f2, C.Connection(label="in"))
    e = C.InferenceEngine(model_engine=g, api=backend)
    for v, f in ((v1, f1), (v2, f2), (v3, f1), (v3, f2)):
        assert C.get_variant(m2v(e, v, f)) == C.MessageToVariable(v, f)
        assert C.get_variant(m2f(e, v, f)) == C.MessageToFactor(v, f)
    for v in (v1, v2, v3):
        assert C.get_variant(marginal(e, v)) == C.IndividualMarginal(v)


def test_engine_rejects_unsupported_backend_and_bad_arguments(backend):  # model_engine.jl:252-266, inference_engine.jl:68-70
    with pytest.raises(C.UnsupportedModelEngineError):
        C.InferenceEngine(model_engine=object(), api=backend)
    with pytest.raises(TypeError):
        C.InferenceEngine(model_engine=G(), dependency_resolver=lambda *a: None, api=backend)
    with pytest.raises(TypeError):
        C.InferenceEngine(model_engine=G(), inference_request_processor=lambda *a: None, api=backend)


def test_empty_scan_for_model_without_pending_messages(backend):  # :93-114
    g = G()
    f1 = g.add_factor(C.Factor(functional_form="left"))
    f2 = g.add_factor(C.Factor(functional_form="right"))
    vc = g.add_variable(C.Variable(name="center"))
    g.add_edge(vc, f1, C.Connection(label="param"))
    g.add_edge(vc, f2, C.Connection(label="param"))
    e = C.InferenceEngine(model_engine=g, api=backend)
    assert C.scan_inference_request(C.request_inference_for(e, vc)) == []


def _small_node_variable_node(backend):  # :120-145
    g = G()
    f1 = g.add_factor(C.Factor(functional_form="left"))
    f2 = g.add_factor(C.Factor(functional_form="right"))
    vc = g.add_variable(C.Variable(name="center"))
    g.add_edge(vc, f1, C.Connection(label="param"))
    g.add_edge(vc, f2, C.Connection(label="param"))
    e = C.InferenceEngine(model_engine=g, resolve_dependencies=False, api=backend)
    vm = marginal(e, vc)
    left, right = e.store.Signal(), e.store.Signal()
    C.add_dependency(m2v(e, vc, f1), left)
    C.add_dependency(m2v(e, vc, f2), right)
    C.add_dependency(vm, m2v(e, vc, f1))
    C.add_dependency(vm, m2v(e, vc, f2))
    return e, f1, f2, vc, left, right


def test_non_empty_scan_for_pending_messages(backend):  # :116-181
    e, f1, f2, vc, left, right = _small_node_variable_node(backend)
    C.set_value(left, 1.0)
    assert C.scan_inference_request(C.request_inference_for(e, vc)) == [m2v(e, vc, f1)]
    e, f1, f2, vc, left, right = _small_node_variable_node(backend)
    C.set_value(right, 1.0)
    assert C.scan_inference_request(C.request_inference_for(e, vc)) == [m2v(e, vc, f2)]
    e, f1, f2, vc, left, right = _small_node_variable_node(backend)
    C.set_value(left, 1.0)
    C.set_value(right, 1.0)
    assert C.scan_inference_request(C.request_inference_for(e, vc)) == [m2v(e, vc, f1), m2v(e, vc, f2)]


def test_scan_resolves_dependencies_of_required_messages(backend):  # :183-239
    g = G()
    v1, v2, v3 = (g.add_variable(C.Variable(name=n)) for n in ("v1", "v2", "v3"))
    f1, f2 = (g.add_factor(C.Factor(functional_form=n)) for n in ("f1", "f2"))
    g.add_edge(v1, f1, C.Connection(label="out"))
    g.add_edge(v2, f1, C.Connection(label="in"))
    g.add_edge(v2, f2, C.Connection(label="out"))
    g.add_edge(v3, f2, C.Connection(label="in"))
    e = C.InferenceEngine(model_engine=g, resolve_dependencies=False, api=backend)
    C.add_dependency(m2v(e, v2, f1), m2f(e, v1, f1))
    C.add_dependency(m2v(e, v2, f2), m2f(e, v3, f2))
    C.add_dependency(marginal(e, v2), m2v(e, v2, f1))
    C.add_dependency(marginal(e, v2), m2v(e, v2, f2))
    C.set_value(m2f(e, v1, f1), 1.0)
    C.set_value(m2f(e, v3, f2), 1.0)
    steps = C.scan_inference_request(C.request_inference_for(e, v2))
    assert steps == [m2v(e, v2, f1), m2v(e, v2, f2)]


def test_default_resolution_1(backend):  # dependencies_tests.jl:39-99
    g = G()
    v1, v2, v3 = (g.add_variable(C.Variable(name=n)) for n in ("v1", "v2", "v3"))
    f1, f2 = (g.add_factor(C.Factor(functional_form=n)) for n in ("f1", "f2"))
    g.add_edge(v1, f1, C.Connection(label="out"))
    g.add_edge(v2, f1, C.Connection(label="out"))
    g.add_edge(v2, f2, C.Connection(label="out"))
    g.add_edge(v3, f2, C.Connection(label="out"))
    e = C.InferenceEngine(model_engine=g, dependency_resolver=C.DefaultDependencyResolver(), api=backend)
    assert C.get_dependencies(marginal(e, v1)) == [m2v(e, v1, f1)]
    assert C.get_dependencies(marginal(e, v2)) == [m2v(e, v2, f1), m2v(e, v2, f2)]
    assert C.get_dependencies(marginal(e, v3)) == [m2v(e, v3, f2)]
    assert C.get_dependencies(m2v(e, v2, f1)) == [m2f(e, v1, f1)]
    assert C.get_dependencies(m2v(e, v2, f2)) == [m2f(e, v3, f2)]
    assert C.get_dependencies(m2f(e, v2, f1)) == [m2v(e, v2, f2)]
    assert C.get_dependencies(m2f(e, v2, f2)) == [m2v(e, v2, f1)]
    # marginal dependencies are intermediate, factor-side ones are not (src/dependencies.jl:25-28 vs :52,67,82)
    assert C.get_dependency_props(marginal(e, v2)) == [cap.NIB_INTERMEDIATE] * 2
    assert C.get_dependency_props(m2v(e, v2, f1)) == [0]


def test_segment_tree_wiring_degree_100(backend):  # src/dependencies.jl:90-173, SURVEY A.3
    e, p, o, f = models.make_beta_bernoulli_model(100, backend)
    n_before = 1 + 100 + 4 * 100  # marginals + 2 signals per connection (2 connections per factor)
    assert e.store.n_signals() == n_before + 98  # n-2 ProductOfMessages nodes
    md = C.get_dependencies(marginal(e, p))
    assert [C.get_variant(s).range for s in md] == [(0, 49), (50, 99)]
    deps = C.get_dependencies(m2f(e, p, f[0]))
    assert [C.get_variant(s).range for s in deps] == [(1, 2), (3, 5), (6, 11), (12, 24), (25, 49), (50, 99)]
    assert all(isinstance(C.get_variant(s), C.ProductOfMessages) for s in deps)


def test_beta_bernoulli_known_answer(backend):  # :241-377
    n = 100
    e, p, o, f = models.make_beta_bernoulli_model(n, backend)
    rng = np.random.Generator(np.random.PCG64(1234))
    data = rng.integers(0, 2, size=n).astype(np.float64)
    C.set_values([m2f(e, o[i], f[i]) for i in range(n)], data.reshape(-1, 1))
    stats = C.update_marginals(e, p)
    a, b = C.get_value(marginal(e, p))
    assert a == pytest.approx(1.0 + data.sum()) and b == pytest.approx(1.0 + n - data.sum())
    assert stats.updates == 199  # 100 m2v + 98 products + 1 marginal
    for again in ((p,), [p]):  # repeated calls are no-ops and must not fail (:353-355)
        assert C.update_marginals(e, again).updates == 0


@pytest.mark.parametrize("form", ["mv", "canon"])
def test_ssm_belief_propagation(backend, form):  # :379-488
    n = 100
    e, x, y, lik, tr = models.make_ssm_model(n, backend, form=form)
    rng = np.random.Generator(np.random.PCG64(1234))
    data = 2.0 * np.arange(1, n + 1) + rng.standard_normal(n)
    models.ssm_set_data(e, y, lik, data)
    stats = C.update_marginals(e, x)
    assert stats.updates == 6 * n - 4
    vals = C.get_values([marginal(e, v) for v in x])
    mv = vals if form == "mv" else models.canon_to_mv(vals)
    assert np.all(mv[:, 0] >= 0) and np.all(np.diff(mv[:, 0]) >= 0) and np.all(mv[:, 1] >= 0)
    ms, Ps = models.rts_smoother(data, 1.0, 1.0)  # independent check
    np.testing.assert_allclose(mv[:, 0], ms, rtol=1e-10)
    np.testing.assert_allclose(mv[:, 1], Ps, rtol=1e-10)


def test_tracing_iid_model(backend):  # :1149-1280
    g = G()
    p = g.add_variable(C.Variable(name="p"))
    o1 = g.add_variable(C.Variable(name="y1"))
    o2 = g.add_variable(C.Variable(name="y2"))
    fp = g.add_factor(C.Factor(functional_form="prior"))
    f1 = g.add_factor(C.Factor(functional_form="likelihood1"))
    f2 = g.add_factor(C.Factor(functional_form="likelihood2"))
    g.add_edge(p, fp, C.Connection(label="out"))
    g.add_edge(p, f1, C.Connection(label="in"))
    g.add_edge(p, f2, C.Connection(label="in"))
    g.add_edge(o1, f1, C.Connection(label="out"))
    g.add_edge(o2, f2, C.Connection(label="out"))
    proc = C.RuleProcessor({"likelihood1": (cap.RULE_SCALE2, []), "likelihood2": (cap.RULE_SCALE2, [])},
                           family=cap.FAMILY_SUM, value_dim=1)
    e = C.InferenceEngine(model_engine=g, dependency_resolver=C.DefaultDependencyResolver(),
                          inference_request_processor=proc, trace=True, api=backend)
    C.set_value(m2f(e, o1, f1), 1)
    C.set_value(m2f(e, o2, f2), 2)
    C.set_value(m2v(e, p, fp), 3)
    C.update_marginals(e, p)
    assert C.get_value(marginal(e, p)) == 9
    trace = C.get_trace(e)
    assert len(trace.inference_requests) == 1
    req = trace.inference_requests[0]
    assert req.request.variable_ids == (p,) and req.total_time_in_ns > 0
    assert len(req.rounds) == 2
    r1, r2 = req.rounds
    assert [C.get_variant(x.signal) for x in r1.executions] == [C.MessageToVariable(p, f1), C.MessageToVariable(p, f2)]
    assert [x.variable_id for x in r1.executions] == [p, p]
    assert [x.value_after_execution for x in r1.executions] == [2, 4]
    assert all(x.value_before_execution == C.UndefValue() for x in r1.executions + r2.executions)  # :1246, 1253, 1261
    assert [C.get_variant(x.signal) for x in r2.executions] == [C.IndividualMarginal(p)]
    assert r2.executions[0].value_after_execution == 9
    # times are measured (per execution on the oracle; per level on the device, CUDA events), not the request time split evenly
    for r in (r1, r2):
        assert all(x.total_time_in_ns > 0 for x in r.executions)
        assert r.total_time_in_ns == sum(x.total_time_in_ns for x in r.executions) <= req.total_time_in_ns
    # a second traced request after ALL inputs were set again (a signal is pending only when every strong dependency is
    # fresh, src/signal.jl:668-730): the old values are reported, no longer UndefValue()
    C.set_value(m2f(e, o1, f1), 5)
    C.set_value(m2f(e, o2, f2), 2)
    C.set_value(m2v(e, p, fp), 3)
    C.update_marginals(e, p)
    req2 = C.get_trace(e).inference_requests[1]
    ex = [x for r in req2.rounds for x in r.executions]
    assert [(x.value_before_execution, x.value_after_execution) for x in ex] == [(2, 10), (4, 4), (9, 17)]


def test_missing_rule_raises(backend):  # src/inference_engine.jl:358-360
    e, x, y, lik, tr = models.make_ssm_model(3, backend, processor=C.RuleProcessor({}, cap.FAMILY_GAUSS_CANON, 2))
    models.ssm_set_data(e, y, lik, [1.0, 2.0, 3.0])
    with pytest.raises(C.NoRuleError):
        C.update_marginals(e, x)


# ---- oracle-only: user-defined (Python) rules, the reference's own extension mechanism ---------------
class _SSMCallback(CallbackProcessor):  # :383-432, literally
    def __init__(self):
        super().__init__(value_dim=2)

    @staticmethod
    def _product(a, b):  # test/runtests.jl:40-46
        xi = a[0] / a[1] + b[0] / b[1]
        w = 1 / a[1] + 1 / b[1]
        return np.array([(1 / w) * xi, 1 / w])

    def _reduce(self, deps):
        acc = C.get_value(deps[0])
        for d in deps[1:]:
            acc = self._product(acc, C.get_value(d))
        return acc

    def compute_individual_marginal(self, engine, variant, signal, dependencies):
        return self._reduce(dependencies)

    def compute_message_to_factor(self, engine, variant, signal, dependencies):
        return self._reduce(dependencies)

    def compute_message_to_variable(self, engine, variant, signal, dependencies):
        assert len(dependencies) == 1
        form = C.get_factor_functional_form(C.get_factor(engine, variant.factor_id))
        v = C.get_value(dependencies[0])
        return np.array([v[0], 1.0]) if form == "likelihood" else np.array([v[0], v[1] + 1.0])


def test_user_defined_python_rules_match_builtin(oracle_api):
    n = 30
    data = 2.0 * np.arange(1, n + 1) + np.random.Generator(np.random.PCG64(7)).standard_normal(n)
    e1, x1, y1, l1, _ = models.make_ssm_model(n, oracle_api, processor=_SSMCallback())
    models.ssm_set_data(e1, y1, l1, data)
    C.update_marginals(e1, x1, schedule="seq")
    e2, x2, y2, l2, _ = models.make_ssm_model(n, oracle_api, form="mv")
    models.ssm_set_data(e2, y2, l2, data)
    C.update_marginals(e2, x2, schedule="seq")
    a = C.get_values([marginal(e1, v) for v in x1])
    b = C.get_values([marginal(e2, v) for v in x2])
    assert np.array_equal(a, b)


def test_unimplemented_python_rule_raises(oracle_api):
    e, x, y, lik, tr = models.make_ssm_model(3, oracle_api, processor=CallbackProcessor(value_dim=2))
    models.ssm_set_data(e, y, lik, [1.0, 2.0, 3.0])
    with pytest.raises(C.NoRuleError):
        C.update_marginals(e, x, schedule="seq")


def test_signal_to_dot_shows_state_and_dependency_flags(backend):  # ext/GraphVizExt/GraphVizExt.jl:292-339
    e, x, y, lik, tr = models.make_ssm_model(3, backend)
    models.ssm_set_data(e, y, lik, [1.0, 2.0, 3.0])
    mg = marginal(e, x[1])
    dot_before = C.signal_to_dot(mg, max_depth=2)
    assert dot_before.startswith("digraph G {") and dot_before.rstrip().endswith("}")
    assert "MainSignal" in dot_before and "IndividualMarginal" in dot_before and "UndefValue()" in dot_before
    assert 'color="gray"' in dot_before  # the marginal's dependencies are intermediate
    C.update_marginals(e, x)
    dot_after = C.signal_to_dot(mg, max_depth=1, show_listeners=False)
    assert 'fillcolor="palegreen"' in dot_after and "UndefValue()" not in dot_after.split("main [")[1].split("];")[0]
