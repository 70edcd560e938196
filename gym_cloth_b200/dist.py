"""Multi-GPU plumbing: environments shard by contiguous global index, nothing is exchanged on the step path.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for timing reductions and the optional
episode-statistics reduction at reporting points."""
import torch
import torch.distributed as dist


def shard_bounds(envs_per_rank, rank):
    """Global env ids [lo, hi) owned by `rank` (weak scaling: every rank owns envs_per_rank environments)."""
    return rank * envs_per_rank, (rank + 1) * envs_per_rank


def reduce_max(values, device):
    """Max over ranks of a list of floats (timings are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def reduce_sum(values, device):
    """Sum over ranks of a list of counts."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def episode_stats(coverage, done, device):
    """Sufficient statistics of per-env results summed over ranks: (sum coverage, sum coverage^2, n, n_done)."""
    c = coverage.double()
    t = torch.stack([c.sum(), (c * c).sum(), torch.tensor(float(c.numel()), dtype=torch.float64, device=c.device),
                     done.double().sum()]).to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    s, s2, n, nd = t.tolist()
    mean = s / n
    return {"mean_coverage": mean, "std_coverage": max(s2 / n - mean * mean, 0.0) ** 0.5, "n_env": int(n), "n_done": int(nd)}
