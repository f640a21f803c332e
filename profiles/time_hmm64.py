"""Kernel time of the K = 64 HMM batch (config 3): the tensor-core kernel (default) against the FFMA2 register kernel
(CXB_HMM64_TC=0), same inputs, results compared. Usage: python profiles/time_hmm64.py [T] [B]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

C = load_package()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K, M = 64, 32
rng = np.random.Generator(np.random.PCG64(3))
A = rng.dirichlet(np.ones(K), size=K)
E = rng.dirichlet(np.ones(K), size=M).T * K
obs = rng.integers(0, M, size=(T, B)).astype(np.uint8)
res = {}
for name, env in (("tensor-core, both recursions per launch", "1"), ("FFMA2 registers, pass after pass", "0")):
    os.environ["CXB_HMM64_TC"] = env
    hm = C.HmmBatch(B, T, K, M, dtype=C.capi.F32)
    hm.set_tables(A, E)
    hm.set_observations(obs)
    ms = []
    for _ in range(3):
        hm.update_marginals()
        ms.append(hm.last_kernel_ms())
    sl = slice(max(0, T // 2 - 3), T // 2 + 3)
    res[env] = (hm.get_forward(T - 4, T), hm.get_marginals(0, 4), hm.get_marginals(sl.start, sl.stop), hm.get_marginals(T - 4, T))
    best = min(ms)
    print(f"{name:44s} T={T} B={B}: {best:9.3f} ms = {best * 1e6 / T * 1.9 / 2:7.1f} cycles per step and pass at 1.9 GHz; "
          f"{B * T * 770 / best * 1e-6:7.1f} GB/s algorithmic", flush=True)
    del hm
for a, b in zip(res["1"], res["0"]):
    err = np.max(np.abs(a - b) / (np.abs(b) + 1e-8))
    print("max rel diff tensor-core vs FFMA2:", err)
