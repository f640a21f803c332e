import sys, numpy as np
sys.path.insert(0, '/root/repo')
from tests._pkg import pkg as C
cap = C.capi
K, beta, sweeps = 16, 0.7, int(sys.argv[2]) if len(sys.argv) > 2 else 2
for N in [int(x) for x in sys.argv[1].split(',')]:
    rng = np.random.Generator(np.random.PCG64(1))
    e = rng.standard_exponential((N, N, K), dtype=np.float32)
    unary = e / e.sum(axis=-1, keepdims=True)
    full = C.PottsGrid(N, N, K, beta, dtype=cap.F32)
    full.set_unary(unary); full.reset_messages()
    for _ in range(sweeps): full.sweep()
    mf = full.get_marginals(); del full
    cut = N // 2
    shards = [C.PottsGrid(cut, N, K, beta, dtype=cap.F32, has_upper=i > 0, has_lower=i < 1) for i in range(2)]
    for i, sh in enumerate(shards):
        sh.set_unary(unary[i * cut:(i + 1) * cut]); sh.reset_messages()
    shards[0].p2p_connect_local(1, shards[1]); shards[1].p2p_connect_local(0, shards[0])
    for _ in range(sweeps):
        for sh in shards: sh.sweep()
    for i, sh in enumerate(shards):
        g = sh.get_marginals()
        d = (g != mf[i * cut:(i + 1) * cut])
        rows = np.flatnonzero(d.any(axis=(1, 2)))
        print(N, "shard", i, "mismatching rows:", len(rows), rows[:8], "cols of first row:", np.flatnonzero(d[rows[0]].any(axis=-1))[:10] if len(rows) else "", "max abs", float(np.abs(g - mf[i*cut:(i+1)*cut]).max()), flush=True)
