"""world_size-2 gloo test (CPU) of the multi-process path: global-id sharding of the synthetic actions is invariant to
the number of ranks, timing is reduced with MAX and episode statistics with SUM - the only collectives bench.py uses (through gym_cloth_b200/dist.py)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_per_rank, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from gym_cloth_b200.dist import episode_stats, reduce_max, reduce_sum, shard_bounds
    lo, hi = shard_bounds(n_per_rank, rank)
    raw, pick = bench.draw_actions(11, 2, lo, hi)
    a = torch.cat([torch.from_numpy(raw), torch.from_numpy(pick).double()[:, None]], 1)
    gathered = [torch.zeros_like(a) for _ in range(world)]
    dist.all_gather(gathered, a)
    raw_f, pick_f = bench.draw_actions(11, 2, 0, n_per_rank * world)
    full = torch.cat([torch.from_numpy(raw_f), torch.from_numpy(pick_f).double()[:, None]], 1)
    ok = torch.equal(torch.cat(gathered), full) and reduce_sum([rank + 1, 5], "cpu") == [3.0, 10.0]
    tmax = reduce_max([10.0 + rank, 3.0 - rank], "cpu")
    cov = torch.full((n_per_rank,), 0.25 * (rank + 1), dtype=torch.float64)
    st = episode_stats(cov, torch.ones(n_per_rank) * rank, "cpu")
    q.put((rank, ok, tmax, st))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, n = 2, 1536
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
    for rank, ok, tmax, st in res:
        assert ok
        assert tmax == [11.0, 3.0]
        assert st["n_env"] == 2 * n and st["n_done"] == n
        assert abs(st["mean_coverage"] - 0.375) < 1e-12 and abs(st["std_coverage"] - 0.125) < 1e-12
