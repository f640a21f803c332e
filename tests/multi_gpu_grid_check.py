"""Multi-process check of the fused (peer-memory) halo exchange on real GPUs, run under torchrun on an N-GPU box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_grid_check.py
Every rank owns a row shard; the shards are connected with CUDA IPC handles (cxb_grid_p2p_*); after a few sweeps the
gathered marginals must equal, bit for bit, those of the same grid swept as ONE shard on rank 0. (Not collected by
pytest: the round-end `-m gpu` box has one GPU; the single-GPU variant is tests/test_device_parity.py.)"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402


def main():
    pkg = entry.load_package()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W, K, beta, sweeps = 64 * world + 3, 40, 16, 0.7, 6
    unary = np.random.Generator(np.random.PCG64(11)).dirichlet(np.ones(K), size=(H, W)).astype(np.float32)
    row0, rows, up, down = pkg.row_shard(H, world, rank)
    gr = pkg.PottsGrid(rows, W, K, beta, device=local, has_upper=up, has_lower=down)
    gr.set_unary(unary[row0:row0 + rows])
    gr.reset_messages()
    assert pkg.connect_row_neighbours(dist, gr, rank, world), "CUDA IPC / peer access unavailable"
    gr.sync()
    dist.barrier()
    for _ in range(sweeps):
        gr.sweep()
    mine = torch.from_numpy(gr.get_marginals()).cuda()
    parts = [torch.empty((pkg.row_shard(H, world, r)[1], W, K), dtype=torch.float32, device="cuda") for r in range(world)]
    dist.all_gather(parts, mine)
    if rank == 0:
        full = pkg.PottsGrid(H, W, K, beta, device=local)
        full.set_unary(unary)
        full.reset_messages()
        for _ in range(sweeps):
            full.sweep()
        got = torch.cat(parts, dim=0).cpu().numpy()
        assert np.array_equal(got, full.get_marginals()), "fused halo exchange differs from the single-shard sweep"
        print(f"multi_gpu_grid_check: {world} shards == 1 shard, bit for bit, after {sweeps} sweeps")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
