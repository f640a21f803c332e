"""The C++ oracle against a second, structurally independent restatement of the reference (oracle/pyref.py: one Python
object per signal, object references, the reference's 1-based chunk arithmetic). Same random scripts on both - random
signal DAGs with weak / intermediate / non-listening dependencies, `set_value!` on inputs, `update_marginals!` on random
variable subsets - and identical observable state (computed, pending, nibbles, values) plus identical execution order
after every operation. Pins the oracle's SEQUENTIAL schedule (what the reference does) beyond the ported known-answer
tests; the level schedule is then pinned to the sequential one by tests/test_schedules.py."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

from tests import fuzz_schedules as fz
from tests._pkg import pkg
from tests.test_schedules import _trace

C = pkg

_spec = importlib.util.spec_from_file_location("cortex_pyref", Path(__file__).resolve().parent.parent / "oracle" / "pyref.py")
R = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(R)


def _strategy(signal, dependencies):
    """The rules of the fuzzer's processor (tests/fuzz_schedules.py::_build): m2v = 2 * its dependency (RULE_SCALE2,
    test/inference_engine_tests.jl:1163-1166), everything else the left-to-right sum (FAMILY_SUM, :1179)."""
    if signal.variant == "MessageToVariable":
        return 2.0 * dependencies[0].value
    acc = dependencies[0].value
    for d in dependencies[1:]:
        acc = acc + d.value
    return acc


def _mirror(engine):
    n = engine.store.n_signals()
    sigs = [R.Signal(variant=type(C.get_variant(C.Signal(engine.store, s))).__name__) for s in range(n)]
    for s, d, weak, inter, listen in engine.fuzz_plan:
        R.add_dependency(sigs[s], sigs[d], weak=weak, intermediate=inter, listen=listen)
    return sigs


def _py_state(sigs):
    st, vals = [], []
    for s in sigs:
        pend = R.is_pending(s)  # like models.engine_state: evaluates the lazy flag of every signal, in id order
        st.append((R.is_computed(s), pend, tuple(s.dependencies_props.nibble(i) for i in range(1, s.dependencies_props.length + 1))))
        vals.append(float(s.value) if R.is_computed(s) else None)
    return st, vals


@pytest.mark.parametrize("flags", [(0.0, 1.0), (0.35, 1.0), (0.0, 0.9), (0.35, 0.8)], ids=["strong", "weak", "nonlistening", "mixed"])
@pytest.mark.parametrize("seed", range(25))
def test_oracle_sequential_schedule_equals_the_object_graph_witness(oracle_api, seed, flags):
    p_weak, p_listen = flags
    rng = np.random.Generator(np.random.PCG64(4242 + seed))
    n_var, n_fac = int(rng.integers(2, 8)), int(rng.integers(1, 8))
    dep_p = float(rng.uniform(0.3, 0.9))
    e, vs, inputs = fz._build(oracle_api, np.random.Generator(np.random.PCG64(int(rng.integers(1 << 30)))), n_var, n_fac, dep_p,
                              p_weak=p_weak, p_listen=p_listen)
    sigs = _mirror(e)
    index_of = {id(s): i for i, s in enumerate(sigs)}
    marg = [C.get_variable_marginal(C.get_variable(e, v)).sid for v in vs]
    # linked signals (src/model_engine.jl:80-83; the final phase of update_marginals! computes the pending ones, :620-627):
    # a few random non-input signals linked to random variables, in the same order on both
    linked = {i: [] for i in range(n_var)}
    candidates = [s for s in range(len(sigs)) if sigs[s].dependencies]
    for s in (rng.choice(candidates, size=min(3, len(candidates)), replace=False) if candidates else []):
        i = int(rng.integers(n_var))
        C.link_signal_to_variable(C.get_variable(e, vs[i]), C.Signal(e.store, int(s)))
        linked[i].append(sigs[int(s)])
    assert fz._state(e) == _py_state(sigs)
    n_updates = n_scanned = 0
    for op in fz._script(rng, n_var, inputs, 20):
        kind, ids, vals = op
        if kind == "set":
            C.set_values([C.Signal(e.store, s) for s in ids], vals.reshape(-1, 1))
            for s, v in zip(ids, vals):
                R.set_value(sigs[s], float(v))
        else:
            # the scanner first (src/inference_engine.jl:540-546): same pending signals in the same DFS order, duplicates included
            request = C.request_inference_for(e, [vs[i] for i in ids])
            scanned = [s.sid for s in C.scan_inference_request(request, order="dfs")]
            R.request_inference_for([sigs[marg[i]] for i in ids], [linked[i] for i in ids])
            assert [index_of[id(s)] for s in R.scan_inference_request([sigs[marg[i]] for i in ids])] == scanned, (seed, op)
            n_scanned += len(scanned)
            C.update_marginals(e, [vs[i] for i in ids], schedule="seq")
            executed = R.update_marginals([sigs[marg[i]] for i in ids], [linked[i] for i in ids], _strategy)
            assert [index_of[id(s)] for s in executed] == _trace(e)[1].tolist(), (seed, op)  # same executions, same order
            n_updates += len(executed)
        assert fz._state(e) == _py_state(sigs), (seed, op)
    assert n_updates >= 0 and n_scanned >= 0
    _SCANNED.append(n_scanned)


_SCANNED = []


def test_scanner_comparisons_were_not_vacuous():
    assert sum(_SCANNED) > 200, sum(_SCANNED)


def test_witness_pending_criteria_across_chunk_boundaries():
    """17 and 33 dependencies: the padded last chunk of is_meeting_pending_criteria (src/signal.jl:708-717) on both."""
    for n in (1, 15, 16, 17, 32, 33):
        s = R.Signal()
        deps = [R.Signal() for _ in range(n)]
        for k, d in enumerate(deps):
            R.add_dependency(s, d, weak=(k % 3 == 0))
        assert not R.is_pending(s)
        for d in deps[:-1]:
            R.set_value(d, 1.0)
        assert not R.is_pending(s)
        R.set_value(deps[-1], 1.0)
        assert R.is_pending(s)
        R.compute(lambda sig, dd: sum(x.value for x in dd), s)
        assert s.value == float(n) and not R.is_pending(s)
        R.set_value(deps[0], 2.0)  # weak: (n == 1 -> the only dependency) fresh again, the others computed but not fresh
        assert R.is_pending(s) == (n == 1)


# ---- the default BP wiring (src/dependencies.jl) on both restatements ---------------------------------------------------
_EXECUTED_ON_WIRED_GRAPHS = []

def _key_of_oracle_signal(sig):
    v = C.get_variant(sig)
    name = type(v).__name__
    if name == "IndividualMarginal":
        return ("marginal", v.variable_id)
    if name == "MessageToVariable":
        return ("m2v", v.variable_id, v.factor_id)
    if name == "MessageToFactor":
        return ("m2f", v.variable_id, v.factor_id)
    if name == "ProductOfMessages":
        return ("product", v.variable_id, v.range[0] + 1, v.range[1] + 1)  # the mirror stores 0-based inclusive positions
    raise AssertionError(name)


@pytest.mark.parametrize("shape", ["loopy", "tree"])
@pytest.mark.parametrize("seed", range(12))
def test_default_wiring_equals_the_object_graph_witness(oracle_api, seed, shape):
    """Random bipartite graphs with leaves, small-degree variables and hubs (segment trees, degree up to 40): every
    signal's dependency list with its nibbles AND its listener list with its listen mask - i.e. the global order of all
    add_dependency! calls - are the same in the C++ oracle and in the literal object-graph transcription."""
    rng = np.random.Generator(np.random.PCG64(31337 + seed))
    n_var, n_fac = int(rng.integers(3, 30)), int(rng.integers(20, 60))
    g = C.BipartiteFactorGraph()
    vs = [g.add_variable(C.Variable(name="v", index=(i,))) for i in range(n_var)]
    edges = []
    hubs = rng.choice(n_var, size=min(2, n_var), replace=False)
    fs = []
    if shape == "tree":  # a random tree with hubs (variable i hangs under an earlier one, often under variable 0) and
        n_var = max(n_var, 16)  # unary factors whose messages are the inputs: BP computes every message in one request
        vs = vs + [g.add_variable(C.Variable(name="v", index=(i,))) for i in range(len(vs), n_var)]
        groups = [(int(rng.integers(0, i)) if i > 8 and rng.random() < 0.5 else 0, i) for i in range(1, n_var)]
        groups += [(i,) for i in range(n_var) if rng.random() < 0.7 or i == 0]
    else:
        groups = []
        for _ in range(n_fac):
            k = int(rng.integers(1, min(4, n_var) + 1))
            members = set(int(x) for x in rng.choice(n_var, size=k, replace=False))
            if rng.random() < 0.6:
                members.add(int(hubs[int(rng.integers(len(hubs)))]))  # hubs collect many factors
            groups.append(tuple(members))
    for members in groups:
        f = g.add_factor(C.Factor(functional_form="f"))
        fs.append(f)
        for v in sorted(members):  # ascending ids: insertion order == sorted order
            g.add_edge(vs[v], f, C.Connection(label="e"))
            edges.append((vs[v], f))
    proc = C.RuleProcessor({"f": (pkg.capi.RULE_SCALE2, [])}, family=pkg.capi.FAMILY_SUM, value_dim=1)
    e = C.InferenceEngine(model_engine=g, dependency_resolver=C.DefaultDependencyResolver(), inference_request_processor=proc,
                          dtype=pkg.capi.F64, api=oracle_api)
    m = R.Model(vs, fs, edges)
    R.resolve_dependencies(m)
    witness = {}
    for s in m.signals():
        witness[s.variant] = ([(d.variant, s.dependencies_props.nibble(i + 1)) for i, d in enumerate(s.dependencies)],
                              [(l.variant, bool(b)) for l, b in zip(s.listeners, s.listenmask)])
    got = {}
    for sid in range(e.store.n_signals()):
        sig = C.Signal(e.store, sid)
        deps = list(zip([_key_of_oracle_signal(d) for d in C.get_dependencies(sig)], C.get_dependency_props(sig)))
        lis = list(zip([_key_of_oracle_signal(l) for l in C.get_listeners(sig)], [bool(b) for b in C.get_listenmask(sig)]))
        got[_key_of_oracle_signal(sig)] = (deps, lis)
    assert max(len(m.factors_of[v]) for v in vs) > 5  # at least one segment tree
    assert got.keys() == witness.keys()
    for k in witness:
        assert got[k] == witness[k], k
    no_factor = [v for v in vs if not m.factors_of[v]]
    assert sorted(w.context for w in C.get_warnings(e)) == sorted(no_factor)

    # ... and the same requests on both: values into every dependency-free signal (messages of single-variable factors),
    # then update_marginals! over random subsets of the variables - same executions in the same order, same state
    by_key = {_key_of_oracle_signal(C.Signal(e.store, sid)): sid for sid in range(e.store.n_signals())}
    py_of = {s.variant: s for s in m.signals()}

    def strategy(signal, dependencies):
        if signal.variant[0] == "m2v":
            return 2.0 * dependencies[0].value
        acc = dependencies[0].value
        for d in dependencies[1:]:
            acc = acc + d.value
        return acc

    def states():
        st_o, val_o = fz._state(e)
        order = [_key_of_oracle_signal(C.Signal(e.store, sid)) for sid in range(e.store.n_signals())]
        st_w, val_w = _py_state([py_of[k] for k in order])
        return (st_o, val_o), (st_w, val_w)

    if shape != "tree":
        return  # loopy graphs: the wiring comparison above is the point (BP does not start without initial messages)
    free = [k for k, s in py_of.items() if not s.dependencies]
    n_executed = 0
    for it in range(4):  # first all inputs (a full BP pass), then new values into half of them (incremental requests)
        pick = list(free) if it == 0 else [free[i] for i in rng.choice(len(free), size=max(1, len(free) // 2), replace=False)]
        vals = rng.integers(1, 5, size=len(pick)).astype(np.float64)
        if pick:
            C.set_values([C.Signal(e.store, by_key[k]) for k in pick], vals.reshape(-1, 1))
            for k, x in zip(pick, vals):
                R.set_value(py_of[k], float(x))
        k_req = n_var if it == 0 else int(rng.integers(1, n_var + 1))
        ids = [int(i) for i in rng.choice(n_var, size=k_req, replace=False)]
        C.update_marginals(e, [vs[i] for i in ids], schedule="seq")
        executed = R.update_marginals([m.marginal[vs[i]] for i in ids], [[] for _ in ids], strategy)
        assert [by_key[s.variant] for s in executed] == _trace(e)[1].tolist()
        a, b = states()
        assert a == b
        n_executed += len(executed)
    _EXECUTED_ON_WIRED_GRAPHS.append(n_executed)


def test_wired_graph_requests_executed_something():
    """Guards the test above against vacuity (runs after it): the requests on the wired graphs did compute signals."""
    assert len(_EXECUTED_ON_WIRED_GRAPHS) == 12 and min(_EXECUTED_ON_WIRED_GRAPHS) > 20, _EXECUTED_ON_WIRED_GRAPHS


@pytest.mark.parametrize("seed", range(20))
def test_duplicate_circular_and_self_dependencies_on_both(oracle_api, seed):
    """The edge cases of test/signal_tests.jl:442-521 at random: duplicate dependencies (only the FIRST matching slot is ever
    notified), cycles and self dependencies (ignored), any mix of flags, dependencies added between computed and uncomputed
    signals (check_computed on and off) - `set_value!` / `compute!(force)` on random signals, identical state on both."""
    rng = np.random.Generator(np.random.PCG64(777 + seed))
    store = pkg.SignalStore(oracle_api, value_dim=1, family=pkg.capi.FAMILY_SUM, dtype=pkg.capi.F64)
    n = int(rng.integers(2, 9))
    co = [C.Signal(store, store.api.create_signal(store.h)) for _ in range(n)]
    py = [R.Signal() for _ in range(n)]

    def same():
        for a, b in zip(co, py):
            assert C.is_computed(a) == R.is_computed(b)
            assert C.is_pending(a) == R.is_pending(b)
            assert tuple(C.get_dependency_props(a)) == tuple(b.dependencies_props.nibble(i) for i in range(1, b.dependencies_props.length + 1))
            assert [x.sid for x in C.get_dependencies(a)] == [py.index(x) for x in b.dependencies]
            assert [x.sid for x in C.get_listeners(a)] == [py.index(x) for x in b.listeners]
            assert [bool(x) for x in C.get_listenmask(a)] == [bool(x) for x in b.listenmask]
            if C.is_computed(a):
                assert float(C.get_values([a])[0][0]) == float(b.value)

    for _ in range(40):
        r = rng.random()
        if r < 0.45:
            s, d = int(rng.integers(n)), int(rng.integers(n))  # s == d: self dependency; repeated pairs: duplicates
            kw = dict(weak=bool(rng.random() < 0.3), listen=bool(rng.random() < 0.8), check_computed=bool(rng.random() < 0.8),
                      intermediate=bool(rng.random() < 0.5))
            C.add_dependency(co[s], co[d], **kw)
            R.add_dependency(py[s], py[d], **kw)
        elif r < 0.85:
            s, v = int(rng.integers(n)), float(rng.integers(1, 9))
            C.set_values([co[s]], np.array([[v]]))
            R.set_value(py[s], v)
        else:
            s = int(rng.integers(n))
            if py[s].dependencies and all(R.is_computed(d) for d in py[s].dependencies):
                for d in co[s:s + 1]:
                    C.compute(d, force=True)  # family SUM: the left-to-right sum of the dependencies
                R.compute(lambda sig, deps: float(sum(x.value for x in deps)), py[s], force=True)
        same()
