#!/usr/bin/env python
"""Generates the committed fixtures of tests/golden/.

Two kinds of fixture, kept apart on purpose:

1. `reference_known_answers.json` — known answers TRANSCRIBED from the reference's own test-suite (Cortex.jl v0.3.0,
   file:line given per case).  The reference cannot be executed here or on the GPU box (no Julia toolchain, the
   BipartiteFactorGraphs dependency is not vendored), so these are the only vectors that come from the reference
   itself; they pin the oracle (tests/test_golden.py, CPU) and through it the CUDA path.
2. `oracle_*.npz` — small seeded input/output vectors produced by running the CPU oracle (oracle/liboracle.so, the
   restatement of the reference's signal/engine semantics) on explicit graphs.  They are REGRESSION fixtures for the
   structured CUDA engines (chains, Potts grid, HMM, pairwise graph): `pytest -m gpu` compares the device results with
   them without needing the oracle at all, and the CPU suite re-derives them so that they cannot silently rot.

    python tests/golden/make_golden.py        # rewrites the files next to this script
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from tests import models  # noqa: E402
from tests._pkg import ORACLE_LIB, pkg  # noqa: E402

C, cap = pkg, pkg.capi


def reference_known_answers():
    return {
        "source": "ReactiveBayes/Cortex.jl v0.3.0 test-suite (transcribed; the reference cannot run in this image)",
        "gaussian_product_mean_variance": {  # test/runtests.jl:40-46
            "cite": "test/runtests.jl:40-46",
            "formula": "w = 1/v1 + 1/v2; v = 1/w; m = v * (m1/v1 + m2/v2)",
            "cases": [{"m1": 0.0, "v1": 1.0, "m2": 2.0, "v2": 1.0, "m": 1.0, "v": 0.5},
                      {"m1": 1.0, "v1": 2.0, "m2": -1.0, "v2": 0.5, "m": -0.6, "v": 0.4}],
        },
        "ssm_rules": {  # test/inference_engine_tests.jl:415-432
            "cite": "test/inference_engine_tests.jl:415-432",
            "observation": "N(y, 1.0)", "transition": "N(m, v + 1.0)",
        },
        "beta_bernoulli": {  # test/inference_engine_tests.jl:241-377
            "cite": "test/inference_engine_tests.jl:360-376",
            "n": 100, "prior": [1.0, 1.0], "posterior": "Beta(1 + sum(y), 1 + n - sum(y))",
            "product_rule": "(a1 + a2 - 1, b1 + b2 - 1)  (test/inference_engine_tests.jl:273-294)",
            "product_of_messages_nodes": 98, "executions": 199,
        },
        "tracing_iid_model": {  # test/inference_engine_tests.jl:1149-1261
            "cite": "test/inference_engine_tests.jl:1149-1261",
            "data": {"o1": 1, "o2": 2, "prior": 3}, "rule": "m2v = 2 * dependency; marginal = sum(dependencies)",
            "marginal": 9, "rounds": 2,
            "round1": [{"variant": "MessageToVariable(p, f1)", "value_after": 2},
                       {"variant": "MessageToVariable(p, f2)", "value_after": 4}],
            "round2": [{"variant": "IndividualMarginal(p)", "value_after": 9}],
        },
        "ssm_bp_properties": {  # test/inference_engine_tests.jl:477-487
            "cite": "test/inference_engine_tests.jl:477-487",
            "data": "y_i = 2 i + noise", "means": "non-negative and non-decreasing", "variances": "non-negative",
        },
        "pending_nibble_layout": {  # src/signal.jl:507-526
            "cite": "src/signal.jl:36-45, 507-526",
            "bits": {"intermediate": 1, "weak": 2, "computed": 4, "fresh": 8}, "deps_per_chunk": 16,
        },
    }


def f32(a):
    """Inputs are made fp32-representable, so that the fp32 engine sees exactly the numbers the fp64 oracle saw and the
    north-star tolerance (rel 1e-5) covers arithmetic only."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def oracle_chain(api):
    rng = np.random.Generator(np.random.PCG64(42))
    T, B = 6, 3
    q, r = f32(rng.uniform(0.5, 2.0, B)), f32(rng.uniform(0.5, 2.0, B))
    y = f32(np.cumsum(rng.standard_normal((T, B)), axis=0) + rng.standard_normal((T, B)))
    marg = np.zeros((T, B, 2))
    for b in range(B):
        e, x, yv, lik, tr = models.make_ssm_model(T, api, form="canon", q=float(q[b]), r=float(r[b]))
        models.ssm_set_data(e, yv, lik, y[:, b])
        C.update_marginals(e, x, schedule="seq")
        marg[:, b] = C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in x])
    return dict(q=q, r=r, y=y, marginals=marg)


def oracle_hmm(api):
    rng = np.random.Generator(np.random.PCG64(43))
    T, K, M, B = 7, 8, 5, 2
    A = f32(rng.dirichlet(np.ones(K), size=K))
    E = f32(rng.dirichlet(np.ones(K), size=M).T * K)
    obs = rng.integers(0, M, size=(T, B)).astype(np.uint8)
    marg = np.zeros((T, B, K))
    for b in range(B):
        e, z, yv, prior, em, tr = models.make_hmm_model(T, K, M, A, E, api)
        models.hmm_set_data(e, z, yv, prior, em, obs[:, b], K)
        C.update_marginals(e, z, schedule="seq")
        marg[:, b] = C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in z])
    return dict(A=A, E=E, obs=obs, marginals=marg)


def oracle_pairwise(api):
    rng = np.random.Generator(np.random.PCG64(44))
    n, K, n_tables, sweeps = 30, 8, 3, 2
    edges = models.chung_lu_edges(n, 60, seed=3)
    tables = f32(np.exp(rng.standard_normal((n_tables, K, K))))
    ttype = rng.integers(0, n_tables, size=len(edges))
    unary = f32(rng.dirichlet(np.ones(K), size=n))
    g = C.BipartiteFactorGraph()
    vs = [g.add_variable(C.Variable(name="v", index=(i,))) for i in range(n)]
    un = [g.add_factor(C.Factor(functional_form="unary")) for _ in range(n)]
    for i in range(n):
        g.add_edge(vs[i], un[i], C.Connection(label="out"))
    for (u, v), t in zip(edges, ttype):
        f = g.add_factor(C.Factor(functional_form=f"pair{int(t)}"))
        g.add_edge(vs[u], f, C.Connection(label="a"))
        g.add_edge(vs[v], f, C.Connection(label="b"))
    proc = C.RuleProcessor({f"pair{t}": (cap.RULE_CAT_TABLE, tables[t].ravel()) for t in range(n_tables)},
                           family=cap.FAMILY_CATEGORICAL, value_dim=K)
    e = C.InferenceEngine(model_engine=g, dependency_resolver=C.DefaultDependencyResolver(), inference_request_processor=proc, api=api)
    models.protocol_b_link(e, vs)
    models.protocol_b_init(e, vs, K)
    usig = [C.get_connection_message_to_variable(e, vs[i], un[i]) for i in range(n)]
    for _ in range(sweeps):
        models.protocol_b_sweep(e, vs, usig, unary, schedule="lvl")
    marg = C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in vs])
    return dict(edges=np.asarray(edges, dtype=np.int64), ttype=ttype.astype(np.int32), tables=tables, unary=unary, sweeps=np.int64(sweeps),
                marginals=marg)


def oracle_vmp(api):
    """The reference's two variational SSM models (test/inference_engine_tests.jl:593-809, 811-1147), small: posterior
    parameters after a few VMP iterations of the reference's call sequence on the SEQUENTIAL schedule."""
    n, iters = 12, 5
    data = models.ssm_mean_field_dataset(n, seed=21)
    m = models.make_ssm_mean_field_model(n, api)
    mf = models.ssm_mean_field_experiment(m[0], m[1], m[2], m[3], m[4], data, iters, schedule="seq")
    m = models.make_ssm_structured_model(n, api)
    st = models.ssm_structured_experiment(m[0], m[1], m[2], m[3], m[4], data, iters, schedule="seq", merged_all=False)
    return dict(data=data, iters=np.int64(iters), mf_x=mf["x"], mf_ssnoise=mf["ssnoise"], mf_obsnoise=mf["obsnoise"],
                st_x=st["x"][:, :2], st_ssnoise=st["ssnoise"][:2], st_obsnoise=st["obsnoise"][:2])


def main():
    (HERE / "reference_known_answers.json").write_text(json.dumps(reference_known_answers(), indent=1) + "\n")
    api = pkg.CApi(ORACLE_LIB, "cxo_")
    np.savez(HERE / "oracle_gauss_chain.npz", **oracle_chain(api))
    np.savez(HERE / "oracle_hmm.npz", **oracle_hmm(api))
    np.savez(HERE / "oracle_pairwise.npz", **oracle_pairwise(api))
    np.savez(HERE / "oracle_vmp.npz", **oracle_vmp(api))
    print("wrote", sorted(p.name for p in HERE.iterdir()))


if __name__ == "__main__":
    main()
