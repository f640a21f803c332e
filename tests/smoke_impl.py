"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle."""
import numpy as np

from tests._pkg import ORACLE_LIB, pkg
from tests import models


def run_smoke():
    import subprocess

    import torch

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    if not ORACLE_LIB.exists():
        subprocess.run(["make", "-C", str(ORACLE_LIB.parent)], check=True)
    C, cap = pkg, pkg.capi
    oracle = C.CApi(ORACLE_LIB, "cxo_")
    dev = C.default_api()
    l0 = dev.kernel_launches()
    # 1. structured kernel: 256 Gaussian chains x T=64 (fp32) vs the oracle's dense restatement
    B, T = 256, 64
    rng = np.random.Generator(np.random.PCG64(1234))
    q, r = rng.uniform(0.5, 2.0, B), rng.uniform(0.5, 2.0, B)
    y = (np.cumsum(rng.standard_normal((T, B)), axis=0) + rng.standard_normal((T, B))).astype(np.float32)
    ch = C.GaussianChainBatch(B, T, dtype=cap.F32)
    ch.set_noise(q, r)
    ch.set_observations(y)
    assert ch.update_marginals() == B * (6 * T - 4)
    ref = np.zeros((6, T, B, 2))
    y64 = np.ascontiguousarray(y.astype(np.float64))
    q32, r32 = q.astype(np.float32).astype(np.float64), r.astype(np.float32).astype(np.float64)  # what the fp32 engine holds
    oracle.chains_reference(B, T, q32.ctypes.data_as(cap.f64p), r32.ctypes.data_as(cap.f64p), y64.ctypes.data_as(cap.f64p),
                            ref.ctypes.data_as(cap.f64p))
    models.assert_values_close(ch.get_marginals(), ref[5], cap.F32, kind="canon")  # element-wise: precision, mean, variance at 1e-5
    # 2. generic engine: explicit Signal graph of one chain, device frontier vs oracle (levels, values)
    Tn = 16
    data = np.cumsum(rng.standard_normal(Tn))
    res = {}
    for name, api in (("o", oracle), ("d", dev)):
        e, x, yv, lik, tr = models.make_ssm_model(Tn, api, form="canon", trace=True)
        models.ssm_set_data(e, yv, lik, data)
        st = C.update_marginals(e, x, schedule="lvl")  # the level-synchronous frontier: per-level lists must be bit-exact
        res[name] = (st.updates, st.levels, models.level_trace(e),
                     C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in x]))
        # and the default schedule (device: level schedule certified against the sequential executor; oracle: the reference loop)
        models.ssm_set_data(e, yv, lik, data + 1.0)
        C.update_marginals(e, x)
        res[name] += (C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in x]),)
    assert res["o"][:3] == res["d"][:3]
    np.testing.assert_allclose(res["d"][3], res["o"][3], rtol=1e-12)
    np.testing.assert_allclose(res["d"][4], res["o"][4], rtol=1e-12)
    print(f"smoke ok: {dev.kernel_launches() - l0} CUDA kernel launches, chains kernel {ch.last_kernel_ms():.3f} ms")
