"""One entry point (src/inference_engine.jl:553-559): graphs built with cxb_graph_build and run with cxb_update_marginals reach
the closed-form kernels. A batch of random-walk chains is recognised from its wiring; once a request has been recorded
(memoised schedule) the values come from k_chain_plan (CXB_RAN_PLAN) and the engine is left EXACTLY as the generic level
schedule leaves it: values bit for bit, computed / pending flags, nibbles - so cxb_get_values, cxb_is_pending, cxb_scan keep
answering - and as the sequential reference leaves it (north-star tolerance)."""
import numpy as np
import pytest

from tests import models
from tests._pkg import pkg as C

cap = C.capi
pytestmark = pytest.mark.gpu


def _values(e):
    st = e.store
    return C.get_values([C.Signal(st, s) for s in range(st.n_signals())])


@pytest.mark.parametrize("dtype", [cap.F64, cap.F32])
@pytest.mark.parametrize("lengths,interleave", [([50], False), ([2, 3, 17, 40], False), ([24] * 9, False), ([24] * 9, True),
                                                # more than one 32-step tile / 32-chain block, both lane mappings, ragged tails, more tiles than warps
                                                ([70] * 40, True), ([70] * 40, False), ([2, 33, 64, 65, 200] * 8, False), ([161], False)])
def test_chain_batches_are_routed_to_the_plan_kernel(oracle_api, device_api, monkeypatch, dtype, lengths, interleave):
    rng = np.random.Generator(np.random.PCG64(11))
    B = len(lengths)
    q, r = rng.uniform(0.5, 2.0, B), rng.uniform(0.5, 2.0, B)

    def build(api, dt):
        return models.make_ssm_batch_model(lengths, api, dtype=dt, q=q, r=r, interleave=interleave)

    ep = build(device_api, dtype)                      # default: memo + plan
    monkeypatch.setenv("CXB_MEMO", "0")
    eg = build(device_api, dtype)                      # the generic level schedule, every time
    monkeypatch.delenv("CXB_MEMO")
    eo = build(oracle_api, cap.F64)
    ran = []
    for rep in range(5):
        rng2 = np.random.Generator(np.random.PCG64(100 + rep))  # the three engines must see the same data
        data = rng2.standard_normal(sum(lengths)).astype(np.float32 if dtype == cap.F32 else np.float64).astype(np.float64)
        for (e, xs, ys, liks, trs) in (ep, eg, eo):
            sig = [C.get_connection_message_to_factor(e, ys[b][t], liks[b][t]) for b in range(B) for t in range(lengths[b])]
            C.set_values(sig, np.stack([data, np.zeros(len(sig))], axis=1))
        ids = lambda m: [v for chain in m[1] for v in chain]  # noqa: E731
        st_p = C.update_marginals(ep[0], ids(ep))
        ran.append(C.last_schedule(ep[0]))
        st_g = C.update_marginals(eg[0], ids(eg))
        C.update_marginals(eo[0], ids(eo), schedule="seq")
        assert (st_p.updates, list(st_p.updates_by_kind)) == (st_g.updates, list(st_g.updates_by_kind))
        assert st_p.updates == sum(6 * T - 4 for T in lengths)
        np.testing.assert_array_equal(_values(ep[0]), _values(eg[0]))  # bit for bit
    assert ran[0] == cap.SCHEDULE_LEVEL and ran[-1] == cap.RAN_PLAN and ran.count(cap.RAN_PLAN) >= 3, ran
    sp, vp = models.engine_state(ep[0])
    sg, vg = models.engine_state(eg[0])
    so, vo = models.engine_state(eo[0])
    assert sp == sg == so  # is_computed / is_pending / nibbles of every signal
    np.testing.assert_array_equal(vp, vg)
    models.assert_values_close(vp, vo, dtype, kind="canon")
    # the engine keeps answering the reactive API after a planned run: nothing is pending, a scan is empty, and new evidence
    # on one chain makes exactly that chain's request productive again
    req = C.request_inference_for(ep[0], ids(ep))
    assert C.scan_inference_request(req, order="id") == []


def test_chain_plan_follows_changed_noise_parameters(oracle_api, device_api):
    """cxb_set_factor_params after the plan was built: the next planned run uses the new variances."""
    lengths = [30, 30]
    ep = models.make_ssm_batch_model(lengths, device_api, q=[1.0, 1.0], r=[1.0, 1.0])
    eo = models.make_ssm_batch_model(lengths, oracle_api, q=[0.3, 2.0], r=[1.7, 0.6])
    rng = np.random.Generator(np.random.PCG64(5))

    def run(m, sched):
        e, xs, ys, liks, trs = m
        sig = [C.get_connection_message_to_factor(e, ys[b][t], liks[b][t]) for b in range(2) for t in range(30)]
        C.set_values(sig, np.stack([data, np.zeros(60)], axis=1))
        C.update_marginals(e, [v for c in xs for v in c], schedule=sched)

    for _ in range(3):
        data = rng.standard_normal(60)
        run(ep, "auto")
    assert C.last_schedule(ep[0]) == cap.RAN_PLAN
    e, xs, ys, liks, trs = ep
    fids = np.ascontiguousarray(liks[0] + liks[1] + trs[0] + trs[1], dtype=np.int64)
    vals = np.ascontiguousarray([1.7] * 30 + [0.6] * 30 + [0.3] * 29 + [2.0] * 29, dtype=np.float64)
    e.store.check(e.api.set_factor_params(e.store.h, len(fids), fids.ctypes.data_as(cap.i64p), vals.ctypes.data_as(cap.f64p)))
    data = rng.standard_normal(60)
    run(ep, "auto")
    assert C.last_schedule(ep[0]) == cap.RAN_PLAN
    run(eo, "seq")
    got = C.get_values([C.get_variable_marginal(C.get_variable(ep[0], v)) for c in ep[1] for v in c])
    want = C.get_values([C.get_variable_marginal(C.get_variable(eo[0], v)) for c in eo[1] for v in c])
    models.assert_values_close(got, want, cap.F64, kind="canon")


@pytest.mark.parametrize("which", ["oracle", "device"])
def test_prepared_lists_and_requests(oracle_api, device_api, which):
    """cxb_prepare_signals / cxb_set_values_prepared / cxb_prepare_request / cxb_update_marginals_prepared leave exactly what
    the per-call forms leave (host values; on the device also values taken from a device buffer)."""
    import torch

    api = oracle_api if which == "oracle" else device_api
    lengths = [20] * 6
    a = models.make_ssm_batch_model(lengths, api)
    b = models.make_ssm_batch_model(lengths, api)
    sig = lambda m: [C.get_connection_message_to_factor(m[0], m[2][c][t], m[3][c][t]) for c in range(6) for t in range(20)]  # noqa: E731
    ids = lambda m: [v for chain in m[1] for v in chain]  # noqa: E731
    plist, preq = C.prepare_signals(b[0], sig(b)), C.prepare_request(b[0], ids(b))
    rng = np.random.Generator(np.random.PCG64(9))
    for rep in range(4):
        vals = np.stack([rng.standard_normal(120), np.zeros(120)], axis=1)
        C.set_values(sig(a), vals)
        C.update_marginals(a[0], ids(a))
        if which == "device" and rep % 2 == 1:  # from device memory
            dev = torch.tensor(vals, dtype=torch.float64, device="cuda")
            torch.cuda.synchronize()
            C.set_values_prepared(plist, None, device_pointer=dev.data_ptr())
        else:
            C.set_values_prepared(plist, vals)
        C.update_marginals(b[0], preq)
    sa, va = models.engine_state(a[0])
    sb, vb = models.engine_state(b[0])
    assert sa == sb
    np.testing.assert_array_equal(va, vb)
    marg = C.prepare_signals(b[0], [C.get_variable_marginal(C.get_variable(b[0], v)) for v in ids(b)])
    np.testing.assert_array_equal(C.get_values_prepared(marg), C.get_values([C.get_variable_marginal(C.get_variable(a[0], v)) for v in ids(a)]))
    if which == "device":
        assert C.last_schedule(b[0]) == cap.RAN_PLAN
    with pytest.raises(ValueError):
        C.prepare_signals(b[0], sig(b)[:3] + sig(b)[:1])  # a repeated signal
