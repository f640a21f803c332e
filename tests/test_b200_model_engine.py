"""`B200ModelEngine` (cortex.jl_b200/model_engine.py; Julia twin: cortex.jl_b200/julia/CortexB200.jl): the device-backed
model-engine backend of SURVEY 8b / 8f.3. The seven generics of src/model_engine.jl:329-391 answer lazily from flat arrays,
and the reference-style read-out works with no signal id in sight."""
import numpy as np
import pytest

from tests import models
from tests._pkg import pkg as C

cap = C.capi


def _chain_arrays(T):
    """x_0..x_{T-1}, y_0..y_{T-1}, likelihood_t, transition_t (test/inference_engine_tests.jl:436-462) as flat arrays."""
    n_ids = 4 * T - 1
    is_factor = np.zeros(n_ids, dtype=np.uint8)
    is_factor[2 * T:] = 1
    forms = {f: ("likelihood" if f < 3 * T else "transition") for f in range(2 * T, n_ids)}
    ev, ef = [], []
    for i in range(T):
        ev += [T + i, i]
        ef += [2 * T + i, 2 * T + i]
    for i in range(T - 1):
        ev += [i, i + 1]
        ef += [3 * T + i, 3 * T + i]
    return n_ids, is_factor, forms, ev, ef


def _engine(api, T, dtype=cap.F64):
    me = C.B200ModelEngine(*_chain_arrays(T))
    proc = C.RuleProcessor({"likelihood": (cap.RULE_GAUSS_OBS, [1.3]), "transition": (cap.RULE_GAUSS_RW, [0.7])},
                           family=cap.FAMILY_GAUSS_CANON, value_dim=2)
    return C.InferenceEngine(model_engine=me, inference_request_processor=proc, dtype=dtype, api=api), me


def test_generics_answer_lazily(backend):
    T = 6
    e, me = _engine(backend, T)
    assert isinstance(C.is_engine_supported(me), C.SupportedModelEngine)
    assert C.get_variable_ids(e) == list(range(2 * T)) and C.get_factor_ids(e) == list(range(2 * T, 4 * T - 1))
    assert C.get_connected_factor_ids(e, 2) == [2 * T + 2, 3 * T + 1, 3 * T + 2]  # ascending ids
    assert C.get_connected_variable_ids(e, 3 * T) == [0, 1]
    assert C.get_factor(e, 2 * T).functional_form == "likelihood"
    v = C.get_variable(e, 3)
    assert C.get_variant(C.get_variable_marginal(v)) == C.IndividualMarginal(3)
    c = C.get_connection(e, 3, 2 * T + 3)
    assert C.get_variant(c.message_to_variable) == C.MessageToVariable(3, 2 * T + 3)
    assert C.get_variant(C.get_connection_message_to_factor(e, 3, 2 * T + 3)) == C.MessageToFactor(3, 2 * T + 3)
    with pytest.raises(KeyError):
        C.get_variable(e, 2 * T)  # a factor id
    with pytest.raises(KeyError):
        C.get_connection(e, 0, 3 * T + 3)  # not connected
    assert not C.is_computed(C.get_variable_marginal(v))


def test_reference_style_readout_equals_the_object_graph_engine(backend):
    """The same model through the lazy backend and through the BipartiteFactorGraph stand-in: identical wiring, and
    get_value(get_variable_marginal(get_variable(engine, v))) gives the same marginals."""
    T = 40
    data = np.cumsum(np.random.Generator(np.random.PCG64(2)).standard_normal(T))
    e, me = _engine(backend, T)
    eg, x, y, lik, tr = models.make_ssm_model(T, backend, form="canon", q=0.7, r=1.3)
    assert e.store.n_signals() == eg.store.n_signals()
    for s in range(e.store.n_signals()):
        a, b = C.Signal(e.store, s), C.Signal(eg.store, s)
        assert [d.sid for d in C.get_dependencies(a)] == [d.sid for d in C.get_dependencies(b)]
        assert C.get_dependency_props(a) == C.get_dependency_props(b)
    for t in range(T):  # set_value!(get_connection_message_to_factor(engine, y_t, likelihood_t), obs)
        C.set_value(C.get_connection_message_to_factor(e, T + t, 2 * T + t), [data[t], 0.0])
    models.ssm_set_data(eg, y, lik, data)
    C.update_marginals(e, list(range(T)))
    C.update_marginals(eg, x)
    got = np.array([C.get_value(C.get_variable_marginal(C.get_variable(e, v))) for v in range(T)])
    want = np.array([C.get_value(C.get_variable_marginal(C.get_variable(eg, v))) for v in x])
    np.testing.assert_array_equal(got, want)
    mv = models.canon_to_mv(got)
    assert np.all(mv[:, 1] > 0)


def test_from_engine_walks_a_source_engine_once(backend):
    g = C.BipartiteFactorGraph()
    p = g.add_variable(C.Variable(name="p"))
    obs = [g.add_variable(C.Variable(name="o", index=(i,))) for i in range(8)]
    fs = [g.add_factor(C.Factor(functional_form="bernoulli")) for _ in range(8)]
    for o, f in zip(obs, fs):
        g.add_edge(p, f, C.Connection(label="out"))
        g.add_edge(o, f, C.Connection(label="in"))
    me = C.B200ModelEngine.from_engine(g)
    proc = C.RuleProcessor({"bernoulli": (cap.RULE_BETA_BERNOULLI, [])}, family=cap.FAMILY_BETA, value_dim=2)
    e = C.InferenceEngine(model_engine=me, inference_request_processor=proc, api=backend)
    data = [1, 0, 1, 1, 0, 1, 1, 1]
    for o, f, d in zip(obs, fs, data):
        C.set_value(C.get_connection_message_to_factor(e, o, f), [float(d), 0.0])
    C.update_marginals(e, p)
    a, b = C.get_value(C.get_variable_marginal(C.get_variable(e, p)))
    assert (a, b) == (1 + sum(data), 1 + len(data) - sum(data))  # Beta(1 + sum, 1 + n - sum), test/inference_engine_tests.jl:360-376
    assert C.get_variable(e, obs[3]).name == "o" and C.get_variable(e, obs[3]).index == (3,)
    assert C.get_connection(e, p, fs[0]).label == "out"
    # 8 factors > 5: the segment tree of src/dependencies.jl:90-173; its ProductOfMessages variants name the neighbour list
    prods = [C.get_variant(C.Signal(e.store, s)) for s in range(e.store.n_signals())]
    prods = [v for v in prods if isinstance(v, C.ProductOfMessages)]
    assert len(prods) == 6 and all(v.factors_connected_to_variable == tuple(fs) for v in prods)
