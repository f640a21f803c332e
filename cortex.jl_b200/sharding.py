"""Multi-GPU plumbing (one process per GPU, torch.distributed): how the hot path shards.

* batches of independent chains (configs 2, 3): `batch_shard` — no data-path collective;
* row-sharded grid (config 4): `row_shard` + `HaloExchanger` — per sweep every rank sends the m2f messages of its
  first/last row to the row-neighbour ranks and receives theirs (8192*K*4 B = 512 KiB per direction at config 4).
The exchange is expressed on torch tensors, so the same code runs over NCCL (tensors wrapping the library's device
halo buffers) and over gloo on CPU (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

from typing import Optional, Tuple


def batch_shard(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of `n_items` independent units: (first, count) of this rank (remainder to the low ranks)."""
    base, rem = divmod(n_items, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def row_shard(n_rows: int, world: int, rank: int) -> Tuple[int, int, bool, bool]:
    """(row0, rows, has_upper_neighbour, has_lower_neighbour) of this rank's row block."""
    row0, rows = batch_shard(n_rows, world, rank)
    if rows == 0:
        raise ValueError("more ranks than grid rows")
    return row0, rows, rank > 0, rank < world - 1


class HaloExchanger:
    """Per-sweep halo exchange between row-neighbour ranks. `dist` is torch.distributed (any backend)."""

    def __init__(self, dist, rank: int, world: int):
        self.dist, self.rank, self.world = dist, rank, world

    def exchange(self, send_up, send_down, recv_up, recv_down):
        """send_up -> rank-1 (lands in its recv_down), send_down -> rank+1 (lands in its recv_up).
        Any of the tensors may be None at the outer borders. Returns when the receives have completed
        (stream-ordered for NCCL)."""
        d, ops = self.dist, []
        if self.rank > 0:
            ops.append(d.P2POp(d.isend, send_up, self.rank - 1))
            ops.append(d.P2POp(d.irecv, recv_up, self.rank - 1))
        if self.rank < self.world - 1:
            ops.append(d.P2POp(d.isend, send_down, self.rank + 1))
            ops.append(d.P2POp(d.irecv, recv_down, self.rank + 1))
        if not ops:
            return
        for w in d.batch_isend_irecv(ops):
            w.wait()


def connect_row_neighbours(dist, grid, rank: int, world: int) -> bool:
    """Fused halo exchange (cxb_grid_p2p_*): every rank publishes the CUDA IPC handles of its halo buffer and sweep
    counters, opens its row neighbours' and from then on `grid.sweep()` stores the cut-edge messages straight into the
    neighbour GPU's memory over NVLink — no separate exchange call. Returns False (and connects nothing) when any rank
    cannot open a neighbour's handle; the caller then keeps using `HaloExchanger` (NCCL)."""
    if world == 1:
        return True
    handles = [None] * world
    dist.all_gather_object(handles, grid.p2p_export())
    ok = 1
    try:
        if rank > 0:
            grid.p2p_connect_ipc(0, handles[rank - 1])
        if rank < world - 1:
            grid.p2p_connect_ipc(1, handles[rank + 1])
    except Exception:  # no peer access between the two GPUs
        ok = 0
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    return all(oks)


def device_tensor(ptr: int, n_elems: int, device_index: int, typestr: str = "<f4"):
    """Zero-copy torch view of a device buffer owned by the CUDA library (for NCCL)."""
    import torch

    class _Arr:
        pass

    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (int(n_elems),), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(a, device=f"cuda:{device_index}")
