"""Memoised level schedules (cxb_update_marginals, CXB_RAN_REPLAY): a request that arrives with the same ids and a flag
state bit-identical to a recorded run replays the recorded levels. The replay must leave EXACTLY what the full level
schedule leaves - values bit for bit, computed / pending flags, nibbles - and what the sequential reference leaves."""
import numpy as np
import pytest

from tests import models
from tests._pkg import pkg as C

cap = C.capi
pytestmark = pytest.mark.gpu


def _same(e1, e2, exact=True):
    s1, v1 = models.engine_state(e1)
    s2, v2 = models.engine_state(e2)
    assert s1 == s2
    if exact:
        np.testing.assert_array_equal(v1, v2)
    else:
        np.testing.assert_allclose(v1, v2, rtol=1e-12, atol=0)


def _quiet_state(e):
    """Values only (engine_state evaluates is_pending on every signal, which changes the cached flags and with them the
    memo key: the replay tests compare the full state at the END only)."""
    st = e.store
    return C.get_values([C.Signal(st, s) for s in range(st.n_signals())])


@pytest.mark.parametrize("resident", ["1", "0"])
def test_chain_requests_are_replayed(oracle_api, device_api, monkeypatch, resident):
    monkeypatch.setenv("CXB_ENGINE_RESIDENT", resident)
    monkeypatch.setenv("CXB_PLAN", "0")  # the recorded levels themselves (a chain would otherwise go to k_chain_plan: tests/test_routing.py)
    T = 40
    rng = np.random.Generator(np.random.PCG64(3))
    em = models.make_ssm_model(T, device_api, form="canon", q=0.7, r=1.3)
    monkeypatch.setenv("CXB_MEMO", "0")
    en = models.make_ssm_model(T, device_api, form="canon", q=0.7, r=1.3)  # same engine without memoisation
    monkeypatch.delenv("CXB_MEMO")
    eo = models.make_ssm_model(T, oracle_api, form="canon", q=0.7, r=1.3)
    ran = []
    for rep in range(6):
        data = np.cumsum(rng.standard_normal(T))
        for (e, x, y, lik, tr) in (em, en, eo):
            models.ssm_set_data(e, y, lik, data)
        st_m = C.update_marginals(em[0], em[1])
        ran.append(C.last_schedule(em[0]))
        st_n = C.update_marginals(en[0], en[1])
        assert C.last_schedule(en[0]) == cap.SCHEDULE_LEVEL
        C.update_marginals(eo[0], eo[1], schedule="seq")
        assert (st_m.updates, st_m.levels, list(st_m.updates_by_kind)) == (st_n.updates, st_n.levels, list(st_n.updates_by_kind))
        np.testing.assert_array_equal(_quiet_state(em[0]), _quiet_state(en[0]))
    assert ran[0] == cap.SCHEDULE_LEVEL and ran[-1] == cap.RAN_REPLAY and ran.count(cap.RAN_REPLAY) >= 4, ran
    _same(em[0], en[0])
    _same(em[0], eo[0], exact=False)


@pytest.mark.parametrize("resident", ["1", "0"])
@pytest.mark.parametrize("family", ["grid", "powerlaw"])
def test_protocol_b_sweeps_are_replayed(oracle_api, device_api, monkeypatch, resident, family):
    monkeypatch.setenv("CXB_ENGINE_RESIDENT", resident)
    K = 8
    rng = np.random.Generator(np.random.PCG64(4))

    def build(api):
        if family == "grid":
            e, pix, un, pair = models.make_grid_model(4, 5, K, 0.7, api)
            vs = [v for row in pix for v in row]
            usig = [C.get_connection_message_to_variable(e, pix[i][j], un[i][j]) for i in range(4) for j in range(5)]
        else:
            e, vs, un, pair, _unary, tables, ttype = models.make_powerlaw_model(60, 150, K, api)
            usig = [C.get_connection_message_to_variable(e, vs[i], un[i]) for i in range(len(vs))]
        models.protocol_b_init(e, vs, K)
        return e, vs, usig

    em = build(device_api)
    monkeypatch.setenv("CXB_MEMO", "0")
    en = build(device_api)
    monkeypatch.delenv("CXB_MEMO")
    eo = build(oracle_api)
    ran = []
    for sweep in range(6):
        unary = rng.dirichlet(np.ones(K), size=len(em[1]))
        models.protocol_b_sweep(em[0], em[1], em[2], unary, schedule="auto")
        ran.append(C.last_schedule(em[0]))
        models.protocol_b_sweep(en[0], en[1], en[2], unary, schedule="auto")
        models.protocol_b_sweep(eo[0], eo[1], eo[2], unary, schedule="seq")
        np.testing.assert_array_equal(_quiet_state(em[0]), _quiet_state(en[0]))
    assert ran[0] == cap.SCHEDULE_LEVEL and ran[-1] == cap.RAN_REPLAY and ran.count(cap.RAN_REPLAY) >= 3, ran
    _same(em[0], en[0])
    so, vo = models.engine_state(eo[0])
    sm, vm = models.engine_state(em[0])
    assert so == sm
    models.assert_values_close(vm, vo, cap.F64, kind="prob")


def test_a_different_flag_state_is_not_replayed(device_api):
    """Same ids, different flag state (one observation missing): the memo must not be used."""
    T = 12
    e, x, y, lik, tr = models.make_ssm_model(T, device_api, form="canon")
    data = np.arange(T, dtype=np.float64)
    for _ in range(3):
        models.ssm_set_data(e, y, lik, data)
        C.update_marginals(e, x)
    assert C.last_schedule(e) in (cap.RAN_REPLAY, cap.RAN_PLAN)
    sig = [C.get_connection_message_to_factor(e, y[i], lik[i]) for i in range(T)]
    C.set_values(sig[1:], np.stack([data[1:], np.zeros(T - 1)], axis=1))  # y_0 not refreshed
    st = C.update_marginals(e, x)
    assert C.last_schedule(e) not in (cap.RAN_REPLAY, cap.RAN_PLAN)
    assert st.updates < 6 * T - 4  # an incomplete request, run by the full schedule
