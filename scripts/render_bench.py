"""Throughput of the image observations (not the contract bench).  Usage: python scripts/render_bench.py [n_env]
Prints one JSON object: images/s of the colour, depth and RGB-D observations over n_env crumpled tier-1 cloths,
device-resident (CUDA events), with the numpy checker's time per image beside it."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gym_cloth_b200 import cfg_path
from gym_cloth_b200.envs import BatchedClothEnv
from gym_cloth_b200.render import ClothRenderer


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    env = BatchedClothEnv(cfg_path(1), n, dtype="f32", seed=3)
    env.reset()
    env.step(torch.from_numpy(bench.actions_for_step(3, 0, 0, n)).to(env.device, torch.float32))
    pos = env.cloth.pos
    r = ClothRenderer(env.P, n, device=env.device)
    out = {"n_env": n, "image": "224x224", "cloth": "25x25, 1152 triangles, crumpled tier-1 states"}
    rgb = torch.empty(n, 224, 224, 3, dtype=torch.uint8, device=env.device)
    for name, fn in (("colour_2x2_samples", lambda: r.rgb_raw(pos, rgb)), ("depth_with_bilateral", lambda: r.depth(pos, out=rgb)),
                     ("rgbd", lambda: r.rgbd(pos))):
        ms = timed(fn)
        out[name] = {"ms": ms, "images_per_s": n / ms * 1e3}
    # output bytes only (the cloth state is 10 KB per image): HBM write rate of the colour pass
    out["colour_2x2_samples"]["hbm_write_gb_per_s"] = n * 224 * 224 * 3 / (out["colour_2x2_samples"]["ms"] * 1e-3) / 1e9
    from oracle import render_oracle as ro
    p0 = pos[0, :, :3].cpu().numpy()
    t = time.perf_counter(); ro.render_rgb(p0, 25); ro.post_depth(ro.render_depth_raw(p0, 25)); out["numpy_checker_s_per_rgbd"] = time.perf_counter() - t
    out["reference"] = "one Blender process per image plus time.sleep(1) (cloth_env.py:255-283): < 1 image/s per env"
    print(json.dumps(out, indent=1))
