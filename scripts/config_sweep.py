"""Measure the other BASELINE.json configs on one GPU (not bench lines; numbers for DESIGN.md / profiles):
  config 3: 16384 envs from tier-2 and tier-3 initial states (8192 each), random pull actions, reference-order f32
  config 4: 64x64 cloths, graph-coloured mode, relax_iters 1 and 2, thickness 0.006, flat start, random actions
  config 5: 65536 tier-1 envs on one GPU
Usage: python scripts/config_sweep.py [config3] [config5] > profiles/rXX_configs.json"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gym_cloth_b200 import cfg_path
from gym_cloth_b200.envs import BatchedClothEnv


def run(tier, n, steps=2, seed=11):
    env = BatchedClothEnv(cfg_path(tier), n, dtype="f32", seed=seed)
    t0 = time.perf_counter(); env.reset(); torch.cuda.synchronize(); reset_s = time.perf_counter() - t0
    c = env.cloth
    start_cov = float(env.start_coverage.mean().item())
    acts = [torch.from_numpy(bench.actions_for_step(seed, t, 0, n)).to(c.device, torch.float32) for t in range(steps + 1)]
    env.step(acts[0]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sub = 0
    e0.record()
    for t in range(steps):
        env.step(acts[1 + t]); sub = sub + c.sim_steps.sum()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"tier": tier, "n_env": n, "reset_seconds": reset_s, "mean_start_coverage": start_cov, "steps": steps,
            "env_steps_per_s": n * steps / (ms * 1e-3), "substeps_per_s": float(sub.item()) / (ms * 1e-3), "ms_per_step": ms / steps,
            "nograb_frac": float(((c.flags & 4) != 0).float().mean().item()), "tear_frac": float(((c.flags & 1) != 0).float().mean().item())}


def run_w64(n, relax, steps=2, seed=11):
    from gym_cloth_b200 import lib as L
    from gym_cloth_b200.batched import BatchedCloth
    P = L.default_params()
    P.num_width_points = P.num_height_points = 64; P.thickness = 0.006; P.reserved0 = relax
    bc = BatchedCloth(P, n, dtype=torch.float32, mode=1)
    acts = [torch.from_numpy(bench.actions_for_step(seed, t, 0, n)).to(bc.device, torch.float32) for t in range(steps + 1)]
    bc.step_actions(acts[0]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sub = 0
    e0.record()
    for t in range(steps):
        bc.step_actions(acts[1 + t]); sub = sub + bc.sim_steps.sum()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"grid": "64x64", "mode": "coloured", "relax_iters": relax, "thickness": 0.006, "n_env": n, "steps": steps,
            "env_steps_per_s": n * steps / (ms * 1e-3), "substeps_per_s": float(sub.item()) / (ms * 1e-3), "ms_per_step": ms / steps,
            "point_updates_per_s": 4096.0 * float(sub.item()) / (ms * 1e-3),
            "nograb_frac": float(((bc.flags & 4) != 0).float().mean().item()), "tear_frac": float(((bc.flags & 1) != 0).float().mean().item())}


if __name__ == "__main__":
    what = sys.argv[1:] or ["config3", "config4", "config5"]
    out = {}
    if "config3" in what:
        out["config3_tier2_8192"] = run(2, 8192)
        out["config3_tier3_8192"] = run(3, 8192)
    if "config4" in what:
        out["config4_64x64_coloured_relax1_1184"] = run_w64(1184, 1)
        out["config4_64x64_coloured_relax2_1184"] = run_w64(1184, 2)
    if "config5" in what:
        out["config5_tier1_65536_one_gpu"] = run(1, 65536)
    print(json.dumps(out, indent=1))
