"""Two row shards of the 8192-wide Potts grid on cuda:0 and cuda:1 in ONE process (peer access, cxb_grid_p2p_connect_local):
the same fused halo path the multi-process bench uses over CUDA IPC. Run under
    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:k_potts_sweep --csv python profiles/nvlink_halo_check.py
to read the NVLink bytes of each sweep kernel: the cut-edge messages of one boundary are 8192 x 16 x 4 B = 512 KiB per direction."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
cap = pkg.capi
N, K, rows, sweeps = 8192, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 4
rng = np.random.Generator(np.random.PCG64(1))
shards = [pkg.PottsGrid(rows, N, K, 0.7, dtype=cap.F32, device=d, has_upper=d > 0, has_lower=d < 1) for d in range(2)]
for sh in shards:
    e = rng.standard_exponential((rows, N, K), dtype=np.float32)
    sh.set_unary(e / e.sum(axis=-1, keepdims=True))
    sh.reset_messages()
shards[0].p2p_connect_local(1, shards[1])
shards[1].p2p_connect_local(0, shards[0])
for _ in range(sweeps):
    for sh in shards:
        sh.sweep()
for sh in shards:
    sh.sync()
print("halo check:", sweeps, "sweeps on 2 devices, marginal checksum", float(shards[0].get_marginals().sum()), float(shards[1].get_marginals().sum()))
