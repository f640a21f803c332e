"""Fuzzer for the class AUTO sends to the LEVEL schedule: graphs wired ONLY by the default resolver (random bipartite graphs with
loops, leaves, hubs above the segment-tree threshold, factors of degree 1..3), random scripts of set_value! on input signals /
on already-computed messages, link_signal_to_variable! of random m2f, and update_marginals! on random variable subsets in random
order. Counts requests the oracle's level schedule ACCEPTS but answers differently from the sequential reference
(python tests/fuzz_bp_graphs.py [n_seeds]); tests/test_fuzz_device.py runs a slice of it on the device."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tests import models  # noqa: E402
from tests._pkg import pkg  # noqa: E402

C = pkg
cap = pkg.capi


def build(api, rng, dtype=cap.F64):
    n_var, n_fac = int(rng.integers(2, 9)), int(rng.integers(1, 12))
    g = C.BipartiteFactorGraph()
    vs = [g.add_variable(C.Variable(name="v", index=(i,))) for i in range(n_var)]
    fs = [g.add_factor(C.Factor(functional_form="f")) for _ in range(n_fac)]
    hub = int(rng.integers(0, n_var)) if rng.random() < 0.4 else -1
    for f in fs:
        k = int(rng.integers(1, min(3, n_var) + 1))
        members = set(int(v) for v in rng.choice(n_var, size=k, replace=False))
        if hub >= 0 and rng.random() < 0.8:
            members.add(hub)  # one variable with many factors: segment tree when it exceeds 5
        for v in sorted(members):
            g.add_edge(vs[v], f, C.Connection(label="e"))
    proc = C.RuleProcessor({"f": (cap.RULE_SCALE2, [])}, family=cap.FAMILY_SUM, value_dim=1)
    e = C.InferenceEngine(model_engine=g, inference_request_processor=proc, dtype=dtype, api=api)
    n = e.store.n_signals()
    inputs = [s for s in range(n) if not C.get_dependencies(C.Signal(e.store, s))]
    m2f = [C.get_connection_message_to_factor(e, v, f).sid for (v, f) in g.edges()]
    return e, vs, inputs, m2f, g.edges()


def script(rng, n_var, inputs, m2f, edges, n_ops):
    ops = []
    for _ in range(n_ops):
        r = rng.random()
        if r < 0.35 and inputs:
            k = int(rng.integers(1, len(inputs) + 1))
            ops.append(("set", [int(x) for x in rng.choice(inputs, size=k, replace=False)], rng.integers(1, 9, size=k).astype(np.float64)))
        elif r < 0.45 and m2f:  # (re)initialise some messages, as the loopy protocols do
            k = int(rng.integers(1, len(m2f) + 1))
            ops.append(("set", [int(x) for x in rng.choice(m2f, size=k, replace=False)], rng.integers(1, 9, size=k).astype(np.float64)))
        elif r < 0.52 and edges:
            v, f = edges[int(rng.integers(0, len(edges)))]
            ops.append(("link", (v, f), None))
        else:
            k = int(rng.integers(1, n_var + 1))
            ops.append(("update", [int(x) for x in rng.choice(n_var, size=k, replace=False)], None))
    return ops


def run(engine, vs, op, schedule):
    kind, arg, vals = op
    if kind == "set":
        # one by one: the signals of a random subset may depend on each other (sequential set_value! semantics)
        for s, v in zip(arg, vals):
            C.set_value(C.Signal(engine.store, s), float(v))
        return "ok"
    if kind == "link":
        v, f = arg
        C.link_signal_to_variable(C.get_variable(engine, v), C.get_connection_message_to_factor(engine, v, f))
        return "ok"
    try:
        C.update_marginals(engine, [vs[i] for i in arg], schedule=schedule)
        return "ok"
    except C.OutOfContractError:
        return "refused"
    except C.NoRuleError:
        return "norule"
    except C.CortexError as e:
        if "does not terminate" in str(e):
            return "diverges"  # the reference's own loop never returns on this request
        raise


def state(engine):
    st, vals = models.engine_state(engine)
    return st, [None if not c else float(v[0]) for (c, _, _), v in zip(st, vals)]


def one_seed(api_a, api_b, seed, sched_a, sched_b, n_ops=18):
    """Runs the same script on two engines; returns 'equal', 'refused' (engine A refused: stop) or 'differs'."""
    rng = np.random.Generator(np.random.PCG64(31000 + seed))
    build_seed = int(rng.integers(1 << 30))
    ea, vsa, inputs, m2f, edges = build(api_a, np.random.Generator(np.random.PCG64(build_seed)))
    eb, vsb, _, _, _ = build(api_b, np.random.Generator(np.random.PCG64(build_seed)))
    for op in script(rng, len(vsa), inputs, m2f, edges, n_ops):
        a = run(ea, vsa, op, sched_a)
        if a == "refused":
            return "refused"
        b = run(eb, vsb, op, sched_b)
        if b == "diverges":
            return "diverges" if a == "diverges" else "reference-diverges"
        if a != b:
            return "differs"
        if a != "ok":
            return "equal"  # both threw the same error midway
        if state(ea) != state(eb):
            return "differs"
    return "equal"


if __name__ == "__main__":
    from tests._pkg import ORACLE_LIB

    api = pkg.CApi(ORACLE_LIB, "cxo_")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    out = {"equal": 0, "refused": 0, "differs": 0, "diverges": 0, "reference-diverges": 0}
    bad = []
    for seed in range(n):
        r = one_seed(api, api, seed, "lvl", "seq")
        out[r] += 1
        if r == "differs":
            bad.append(seed)
    print(f"default-resolver graphs, level (strict) vs sequential, {n} scripts: {out}; differing seeds: {bad[:20]}")
