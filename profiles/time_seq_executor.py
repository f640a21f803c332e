import sys, time, numpy as np
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
from tests import models
from tests._pkg import pkg as C
cap = C.capi
api = C.default_api()
for T in (1000, 4000):
    e, x, y, lik, tr = models.make_ssm_model(T, api, form="canon")
    data = np.cumsum(np.random.default_rng(0).standard_normal(T))
    for sched in ("seq", "lvl"):
        ts = []
        for rep in range(3):
            models.ssm_set_data(e, y, lik, data + rep)
            t0 = time.perf_counter(); st = C.update_marginals(e, x, schedule=sched); ts.append(time.perf_counter() - t0)
        print("T", T, sched, "updates", st.updates, "ms", [round(1e3 * t, 2) for t in ts], "us per update", round(1e6 * min(ts) / st.updates, 3), flush=True)
