"""Reset throughput of BatchedClothEnv: wall time of seed() + reset() for n environments against the
device time of the same reset (CUDA events around the whole reset minus host gaps are not separable, so the device share is
taken from the kernels' own clocks: substeps run x measured substeps/s), plus a cProfile of the host side.
Usage: python scripts/reset_bench.py [n_env] [tier] [--profile]"""
import cProfile, io, json, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_cloth_b200 import cfg_path
from gym_cloth_b200.envs import BatchedClothEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tier = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t0 = time.perf_counter()
env = BatchedClothEnv(cfg_path(tier), n, dtype="f32", seed=1337)
env.time_resets = True
torch.cuda.synchronize()
t_make = time.perf_counter() - t0
pr = cProfile.Profile() if "--profile" in sys.argv else None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
if pr:
    pr.enable()
env.reset()
if pr:
    pr.disable()
e1.record(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
dev_busy = getattr(env, "reset_device_ms", None)
out = {"n_env": n, "tier": tier, "construct_and_seed_s": t_make, "reset_wall_s": wall, "resets_per_s": n / wall,
       "reset_device_busy_s": None if dev_busy is None else dev_busy / 1e3,
       "wall_over_device": None if not dev_busy else wall / (dev_busy / 1e3),
       "mean_start_coverage": float(env.start_coverage.mean().item())}
print(json.dumps(out))
if pr:
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:6000])
