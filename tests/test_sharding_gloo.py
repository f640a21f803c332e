"""N>1 host logic on CPU: world_size-2 and -3 gloo runs of the row-sharded Potts grid.
Each rank owns a row block as an explicit oracle graph with a ghost row for the neighbour shard; every sweep
the cut-edge m2f messages travel through `HaloExchanger` (the same code bench.py drives over NCCL). The sharded
result must be bit-identical to the single-graph result."""
import os
import socket

import numpy as np
import pytest

from tests._pkg import pkg
from tests import models

C = pkg
cap = pkg.capi
H, W, K, BETA, SWEEPS = 7, 5, 4, 0.7, 4


def _unary():
    return np.random.Generator(np.random.PCG64(42)).dirichlet(np.ones(K), size=(H, W))


def _run_full(api):
    e, pix, un, pair = models.make_grid_model(H, W, K, BETA, api)
    vids = [v for row in pix for v in row]
    models.protocol_b_init(e, vids, K)
    usig = [C.get_connection_message_to_variable(e, pix[i][j], un[i][j]) for i in range(H) for j in range(W)]
    for _ in range(SWEEPS):
        models.protocol_b_sweep(e, vids, usig, _unary().reshape(-1, K), schedule="seq")
    return C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in vids]).reshape(H, W, K)


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests._pkg import ORACLE_LIB

        api = C.CApi(ORACLE_LIB, "cxo_")
        row0, rows, has_up, has_down = C.row_shard(H, world, rank)
        e, pix, un, pair = models.make_grid_model(rows, W, K, BETA, api, ghost_top=has_up, ghost_bottom=has_down)
        real = pix[int(has_up): int(has_up) + rows]
        unr = un[int(has_up): int(has_up) + rows]
        vids = [v for row in real for v in row]
        models.protocol_b_init(e, [v for row in pix for v in row], K)
        usig = [C.get_connection_message_to_variable(e, real[i][j], unr[i][j]) for i in range(rows) for j in range(W)]
        unary = _unary()[row0:row0 + rows].reshape(-1, K)
        fac = {}
        for f, a, b in pair:
            fac[(a, b)] = f
            fac[(b, a)] = f
        ex = C.HaloExchanger(dist, rank, world)

        def cut(ghost_row, real_row):  # (signals we send, signals we receive into) across one cut
            send = [C.get_connection_message_to_factor(e, r, fac[(r, g)]) for r, g in zip(real_row, ghost_row)]
            recv = [C.get_connection_message_to_factor(e, g, fac[(r, g)]) for r, g in zip(real_row, ghost_row)]
            return send, recv

        up = cut(pix[0], real[0]) if has_up else None
        down = cut(pix[-1], real[-1]) if has_down else None
        for _ in range(SWEEPS):
            models.protocol_b_sweep(e, vids, usig, unary, schedule="lvl")
            s_up = torch.from_numpy(C.get_values(up[0]).copy()) if up else None
            s_dn = torch.from_numpy(C.get_values(down[0]).copy()) if down else None
            r_up = torch.empty((W, K), dtype=torch.float64) if up else None
            r_dn = torch.empty((W, K), dtype=torch.float64) if down else None
            ex.exchange(s_up, s_dn, r_up, r_dn)
            if up:
                C.set_values(up[1], r_up.numpy())
            if down:
                C.set_values(down[1], r_dn.numpy())
        out[rank] = (row0, C.get_values([C.get_variable_marginal(C.get_variable(e, v)) for v in vids]).reshape(rows, W, K))
    finally:
        dist.destroy_process_group()


def test_row_and_batch_shard_cover_everything():
    for n, world in ((8192, 8), (7, 2), (10, 3), (65536, 8), (5, 5)):
        blocks = [C.batch_shard(n, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        rs = [C.row_shard(n, world, r) for r in range(world)]
        assert [x[2] for x in rs] == [r > 0 for r in range(world)] and [x[3] for x in rs] == [r < world - 1 for r in range(world)]
    with pytest.raises(ValueError):
        C.row_shard(2, 3, 2)


@pytest.mark.parametrize("world", [2, 3])  # 3 ranks: the middle shard exchanges with both neighbours in the same sweep
def test_gloo_halo_exchange_matches_single_graph(oracle_api, world):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert sorted(out.keys()) == list(range(world))
    got = np.concatenate([out[r][1] for r in sorted(out.keys())], axis=0)
    assert np.array_equal(got, _run_full(oracle_api))
