"""Exploratory (not collected): what the strict level-schedule rules cost. For every random script, at the FIRST request the
strict schedule (CXO_STRICT_FRESHNESS=1) refuses, the same request is given to the default level schedule and to the
sequential one: was it refused by the default rules as well, and if not, would the default answer have been equal?
Round 1, 1,500 scripts: strong listening dependencies 986 refused = 782 by the default rules too + 14 necessary + 190
unnecessary; 10 % non-listening: 1,126 = 445 + 156 necessary + 525 unnecessary. Run: python tests/fuzz_strict_cost.py 1500"""
import sys, os
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import numpy as np
from tests import fuzz_schedules as fz
from tests._pkg import ORACLE_LIB, pkg
api = pkg.CApi(str(ORACLE_LIB), "cxo_")
pw = float(os.environ.get("FZ_WEAK", "0")); pl = float(os.environ.get("FZ_LISTEN", "1"))
n_seeds = int(sys.argv[1])
stat = dict(strict_refused=0, also_default_refused=0, default_equal=0, default_differs=0)
for seed in range(n_seeds):
    rng = np.random.Generator(np.random.PCG64(9000 + seed))
    n_var, n_fac = int(rng.integers(2, 8)), int(rng.integers(1, 8))
    dep_p = float(rng.uniform(0.3, 0.9))
    build_seed = int(rng.integers(1 << 30))
    eng = []
    for strict in ("1", "0", "0"):
        os.environ["CXO_STRICT_FRESHNESS"] = strict
        e, vs, inputs = fz._build(api, np.random.Generator(np.random.PCG64(build_seed)), n_var, n_fac, dep_p, p_weak=pw, p_listen=pl)
        eng.append((e, vs))
    for op in fz._script(rng, n_var, inputs, 20):
        a = fz._run(eng[0][0], eng[0][1], op, "lvl")
        if a != "ok":
            if a == "refused":
                stat["strict_refused"] += 1
                b = fz._run(eng[1][0], eng[1][1], op, "lvl")
                if b != "ok":
                    stat["also_default_refused"] += 1
                else:
                    c = fz._run(eng[2][0], eng[2][1], op, "seq")
                    same = c == "ok" and fz._state(eng[1][0]) == fz._state(eng[2][0])
                    stat["default_equal" if same else "default_differs"] += 1
            break
        fz._run(eng[1][0], eng[1][1], op, "lvl"); fz._run(eng[2][0], eng[2][1], op, "seq")
        if fz._state(eng[0][0]) != fz._state(eng[2][0]): break
print(f"weak={pw} listen={pl}:", stat)
