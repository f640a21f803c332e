"""The model-engine interface of the drop-in boundary: ports of test/model_engine_tests.jl,
test/ext/bipartite_factor_graphs_ext_tests.jl, test/dependencies_tests.jl:1-37 and test/inference_engine_tests.jl:1-31
against the host-side mirror (cortex.jl_b200/model_engine.py, inference_engine.py)."""
import pytest

from tests._pkg import pkg

C = pkg
cap = pkg.capi


def test_create_variable():  # model_engine_tests.jl:1-21
    for name in ("v", "v1", "v2", "v3"):
        v = C.Variable(name=name)
        assert C.get_variable_name(v) == name and C.get_variable_index(v) is None
        assert isinstance(C.get_variable_linked_signals(v), list)
    for index in (1, 2, 3):
        v = C.Variable(name="v", index=index)
        assert C.get_variable_name(v) == "v" and C.get_variable_index(v) == index


def test_linked_signals_empty_by_default_and_linkable(oracle_api):  # :23-52
    assert C.get_variable_linked_signals(C.Variable(name="v")) == []
    g = C.BipartiteFactorGraph()
    v1 = g.add_variable(C.Variable(name="v1"))
    e = C.InferenceEngine(model_engine=g, api=oracle_api)
    some_other_signal = C.create_inference_signal(e)
    var = C.get_variable(e, v1)
    C.link_signal_to_variable(var, some_other_signal)
    assert C.get_variable_linked_signals(var) and some_other_signal in C.get_variable_linked_signals(var)


def test_variable_marginal_is_a_signal_once_bound(oracle_api):  # :31-39 (signals live in the engine's store here)
    g = C.BipartiteFactorGraph()
    v = g.add_variable(C.Variable(name="v"))
    e = C.InferenceEngine(model_engine=g, api=oracle_api)
    mg = C.get_variable_marginal(C.get_variable(e, v))
    assert isinstance(mg, C.Signal) and isinstance(C.get_variant(mg), C.IndividualMarginal)


def test_create_factor_and_local_marginals(oracle_api):  # :54-84
    for form in ("f", "g", "h"):
        f = C.Factor(functional_form=form)
        assert C.get_factor_functional_form(f) == form and C.get_factor_local_marginals(f) == []
    g = C.BipartiteFactorGraph()
    fid = g.add_factor(C.Factor(functional_form="f"))
    e = C.InferenceEngine(model_engine=g, api=oracle_api)
    local_marginal = C.create_inference_signal(e)
    C.add_local_marginal_to_factor(C.get_factor(e, fid), local_marginal)
    assert local_marginal in C.get_factor_local_marginals(C.get_factor(e, fid))


def test_create_connection():  # :86-112
    for label in ("c", "d", "e"):
        for index in (1, 2, 3):
            c = C.Connection(label=label, index=index)
            assert C.get_connection_label(c) == label and C.get_connection_index(c) == index
    assert C.get_connection_index(C.Connection(label="c")) == 0


def test_unsupported_model_engine_error_message():  # :114-126
    assert str(C.UnsupportedModelEngineError(1, None)) == "The model engine of type `int` is not supported."
    assert str(C.UnsupportedModelEngineError(1, "get_variable")) == \
        "The model engine of type `int` does not implement the function `get_variable`."
    assert str(C.UnsupportedModelEngineError(1, "get_factor")) == \
        "The model engine of type `int` does not implement the function `get_factor`."


def test_engine_for_unsupported_backend_throws(oracle_api):  # :128-142
    class MyDummyUnsupportedEngine:
        pass

    for bad in (1, "string", MyDummyUnsupportedEngine()):
        with pytest.raises(C.UnsupportedModelEngineError, match="is not supported"):
            C.InferenceEngine(model_engine=bad, api=oracle_api)


def test_backend_without_interface_methods_throws():  # :144-175
    class NoMethods:
        pass

    dummy = NoMethods()
    calls = [(C.backend_get_variable, (1,), "get_variable"), (C.backend_get_factor, (1,), "get_factor"),
             (C.backend_get_variable_ids, (), "get_variable_ids"), (C.backend_get_factor_ids, (), "get_factor_ids"),
             (C.backend_get_connection, (1, 1), "get_connection"),
             (C.backend_get_connected_variable_ids, (1,), "get_connected_variable_ids"),
             (C.backend_get_connected_factor_ids, (1,), "get_connected_factor_ids")]
    for fn, args, name in calls:
        with pytest.raises(C.UnsupportedModelEngineError) as ei:
            fn(dummy, *args)
        assert ei.value.missing_function == name and ei.value.model_engine is dummy


def test_bipartite_backend_supported_with_every_constructor_switch(backend):  # ext tests :1-16
    g = C.BipartiteFactorGraph()
    assert isinstance(C.InferenceEngine(model_engine=g, api=backend), C.InferenceEngine)
    assert isinstance(C.InferenceEngine(model_engine=C.BipartiteFactorGraph(), resolve_dependencies=False, api=backend), C.InferenceEngine)
    assert isinstance(C.InferenceEngine(model_engine=C.BipartiteFactorGraph(), prepare_signals_metadata=False, api=backend), C.InferenceEngine)


def test_backend_generics_through_the_engine(backend):  # ext tests :18-101
    g = C.BipartiteFactorGraph()
    v1 = g.add_variable(C.Variable(name="a"))
    v2 = g.add_variable(C.Variable(name="b", index=(1,)))
    v3 = g.add_variable(C.Variable(name="c", index=(2, 3)))
    f1 = g.add_factor(C.Factor(functional_form="f1"))
    f2 = g.add_factor(C.Factor(functional_form="f2"))
    g.add_edge(v1, f1, C.Connection(label="out"))
    g.add_edge(v2, f2, C.Connection(label="theta"))
    e = C.InferenceEngine(model_engine=g, api=backend)
    assert [C.get_variable_name(C.get_variable(e, v)) for v in (v1, v2, v3)] == ["a", "b", "c"]
    assert [C.get_variable_index(C.get_variable(e, v)) for v in (v1, v2, v3)] == [None, (1,), (2, 3)]
    for v in (v1, v2, v3):
        assert isinstance(C.get_variable_marginal(C.get_variable(e, v)), C.Signal)
    assert C.get_factor_functional_form(C.get_factor(e, f1)) == "f1"
    assert C.get_factor_functional_form(C.get_factor(e, f2)) == "f2"
    for (v, f, label) in ((v1, f1, "out"), (v2, f2, "theta")):
        c = C.get_connection(e, v, f)
        assert isinstance(c, C.Connection) and C.get_connection_label(c) == label
        assert isinstance(C.get_connection_message_to_variable(c), C.Signal)
        assert isinstance(C.get_connection_message_to_factor(c), C.Signal)
        assert C.get_connection_message_to_variable(e, v, f) == C.get_connection_message_to_variable(c)
        assert C.get_connection_message_to_factor(e, v, f) == C.get_connection_message_to_factor(c)
        assert C.get_variant(C.get_connection_message_to_variable(c)) == C.MessageToVariable(v, f)
        assert C.get_variant(C.get_connection_message_to_factor(c)) == C.MessageToFactor(v, f)
    with pytest.raises(Exception):
        C.get_connection(e, v1, f2)
    with pytest.raises(Exception):
        C.get_connection(e, v2, f1)
    assert set(C.get_variable_ids(e)) == {v1, v2, v3} and set(C.get_factor_ids(e)) == {f1, f2}
    assert set(C.get_connected_variable_ids(e, f1)) == {v1} and set(C.get_connected_variable_ids(e, f2)) == {v2}
    assert set(C.get_connected_factor_ids(e, v1)) == {f1} and set(C.get_connected_factor_ids(e, v2)) == {f2}
    assert set(C.get_connected_factor_ids(e, v3)) == set()


def test_resolve_dependencies_visits_every_factor_then_every_variable(backend):  # dependencies_tests.jl:1-37
    class CustomDependencyResolver(C.AbstractDependencyResolver):
        def __init__(self):
            self.order = []

        def resolve_variable_dependencies(self, engine, variable_id):
            self.order.append(("variable", variable_id))

        def resolve_factor_dependencies(self, engine, factor_id):
            self.order.append(("factor", factor_id))

    g = C.BipartiteFactorGraph()
    x, y, z = (g.add_variable(C.Variable(name=n)) for n in "xyz")
    f1, f2 = (g.add_factor(C.Factor(functional_form=n)) for n in ("f1", "f2"))
    e = C.InferenceEngine(model_engine=g, api=backend)
    resolver = CustomDependencyResolver()
    resolver.resolve_dependencies(e)
    assert {i for k, i in resolver.order if k == "variable"} == {x, y, z}
    assert {i for k, i in resolver.order if k == "factor"} == {f1, f2}
    kinds = [k for k, _ in resolver.order]
    assert kinds == ["factor"] * 2 + ["variable"] * 3  # src/dependencies.jl:5-15: factors first


def test_isa_variant_and_variant_reprs(backend):  # inference_engine_tests.jl:1-31
    g = C.BipartiteFactorGraph()
    v = g.add_variable(C.Variable(name="v"))
    f = g.add_factor(C.Factor(functional_form="f"))
    g.add_edge(v, f, C.Connection(label="out"))
    e = C.InferenceEngine(model_engine=g, api=backend)
    s = C.create_inference_signal(e)
    assert C.isa_variant(s, C.Unspecified)
    for variant, others in ((C.MessageToFactor(v, f), (C.MessageToVariable, C.IndividualMarginal, C.JointMarginal)),
                            (C.MessageToVariable(v, f), (C.MessageToFactor, C.IndividualMarginal, C.JointMarginal)),
                            (C.IndividualMarginal(v), (C.MessageToFactor, C.MessageToVariable, C.JointMarginal)),
                            (C.JointMarginal(f, (v,)), (C.MessageToFactor, C.MessageToVariable, C.IndividualMarginal))):
        C.set_variant(s, variant)
        assert C.isa_variant(s, type(variant)) and C.get_variant(s) == variant
        assert not any(C.isa_variant(s, T) for T in others)
        assert type(variant).__name__ in repr(C.get_variant(s))


def test_format_time_ns():  # test/util_tests.jl
    cases = {100: "100 ns", 999: "999 ns", 1_000: "1.0 μs", 1_234: "1.23 μs", 999_999: "1000.0 μs", 999_000: "999.0 μs",
             500_500: "500.5 μs", 1_000_000: "1.0 ms", 1_234_567: "1.23 ms", 999_000_000: "999.0 ms", 1_000_000_000: "1.0 s",
             1_234_000_000: "1.23 s", 59_000_000_000: "59.0 s", 60_000_000_000: "1.0 min", 90_000_000_000: "1.5 min",
             3_540_000_000_000: "59.0 min", 3_600_000_000_000: "1.0 hr", 5_400_000_000_000: "1.5 hr",
             999_999_999: "1000.0 ms", 59_999_999_999: "60.0 s", 3_599_999_999_999: "60.0 min", 1_235: "1.24 μs",
             1_230: "1.23 μs", 1_234_000: "1.23 ms", 1_235_000: "1.24 ms", 0: "0 ns"}
    for ns, want in cases.items():
        assert C.format_time_ns(ns) == want, ns


def test_add_warning_and_resolve_dependencies_entry_points(oracle_api):  # src/inference_engine.jl:127-129, src/dependencies.jl:5-15
    g = C.BipartiteFactorGraph()
    v = g.add_variable(C.Variable(name="v"))
    f = g.add_factor(C.Factor(functional_form="f"))
    g.add_edge(v, f, C.Connection(label="out"))
    e = C.InferenceEngine(model_engine=g, resolve_dependencies=False, api=oracle_api)
    assert C.get_dependencies(C.get_variable_marginal(C.get_variable(e, v))) == []
    C.resolve_dependencies(C.DefaultDependencyResolver(), e)
    assert len(C.get_dependencies(C.get_variable_marginal(C.get_variable(e, v)))) == 1
    C.add_warning(e, "something", 3)
    assert [(w.description, w.context) for w in C.get_warnings(e)] == [("something", 3)]
