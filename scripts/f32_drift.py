"""How far does the f32 production build drift from the bit-exact f64 build over one pull action, and is the drift
rounding + chaos or something the f32-only code paths add?

Three builds run the SAME schedule from the SAME start states (tier-1 reset pool, f64 states rounded to f32 for the
float builds):  f64 (bit-exact with the reference), f32 (production: rsqrt / squared-distance forms, -ftz, approximate
div/sqrt) and f32ieee (study variant libclothb200_f32ieee.so: the reference's expressions in float, IEEE div/sqrt, no
flush-to-zero).  Schedule = SURVEY.md App. E-2's: grip a cloth point, lift 50, rest 80, pull `iters_pull` substeps,
grip rest 300, release, rest 1000 - driven substep by substep through Gripper.adjust / Cloth.update so that the states
can be compared at fixed horizons.

    python scripts/f32_drift.py [n_env] [out.json]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_cloth_b200 import cfg_path, lib as L          # noqa: E402
from gym_cloth_b200.batched import BatchedCloth        # noqa: E402

HORIZONS = (1, 10, 50, 130, 230, 430, 530, 730, 1000, 1730)


def start_states(n, seed=1337):
    """Tier-1 reset states (cloth_env.py:843-891) from the f64 build."""
    from gym_cloth_b200.envs import BatchedClothEnv
    env = BatchedClothEnv(cfg_path(1), n, dtype="f64", seed=seed)
    env.reset()
    return env.cloth.pos.clone(), env.cloth.prev.clone()


def run_schedule(bc, grip_xy, delta, horizons=HORIZONS):
    """One pull action, substep by substep (cloth_env.py:352-367, 495-515); returns {horizon: (pos, coverage)} and the
    grabbed masks."""
    P = bc.P
    plan = bc.decode_host(np.array([[0.0, 0.0, delta[0], delta[1]]]))[0]
    ip, dxr, dyr = plan.iters_pull, plan.dxr, plan.dyr
    e0 = int(P.iters_up); e1 = e0 + int(P.iters_up_rest); e2 = e1 + ip; e3 = e2 + int(P.iters_grip_rest)
    total = e3 + int(P.iters_rest)
    bc.grab_top(grip_xy)
    grabbed = bc.grab_mask.clone(); ngrab = bc.n_grabbed.clone()
    out = {}
    marks = sorted(set(h for h in horizons if h <= total) | {total})
    i = 0
    released = False
    while i < total:
        if i < e0:
            bc.adjust(0.0, 0.0, 0.0025); bc.update(1); i += 1
        elif i < e1:
            stop = min([m for m in marks if m > i] + [e1]); stop = min(stop, e1)
            bc.update(stop - i); i = stop
        elif i < e2:
            bc.adjust(dxr, dyr, 0.0); bc.update(1); i += 1
        elif i < e3:
            stop = min([m for m in marks if m > i] + [e3]); stop = min(stop, e3)
            bc.update(stop - i); i = stop
        else:
            if not released:
                bc.release(); released = True
            stop = min([m for m in marks if m > i] + [total])
            bc.update(stop - i); i = stop
        if i in marks:
            bc.measure()
            torch.cuda.synchronize()
            out[i] = (bc.pos[:, :, :3].double().clone(), bc.coverage.clone(), bc.flags.clone())
    return out, grabbed, ngrab, {"iters_pull": ip, "total": total}


def compare(a, b):
    """per-environment max / mean |dpos| and |dcoverage| between two builds' checkpoints"""
    d = (a[0] - b[0]).abs()
    return d.amax(dim=(1, 2)).cpu().numpy(), d.mean(dim=(1, 2)).cpu().numpy(), (a[1] - b[1]).abs().cpu().numpy()


def drift_study(n=512, seed=1337, delta=(0.36, -0.48), variants=("f32", "f32ieee"), horizons=HORIZONS):
    pos0, prev0 = start_states(n, seed)
    N = pos0.shape[1]
    # grip a point of each cloth (its f64 start state) so that every environment does work
    idx = (37 * torch.arange(n, device=pos0.device) + 11) % N
    xy = pos0[torch.arange(n), idx, :2].cpu().numpy()
    P = L.default_params()
    runs = {}
    for name in ("f64",) + tuple(variants):
        dt = torch.float64 if name == "f64" else torch.float32
        bc = BatchedCloth(P, n, dtype=dt, variant="f32ieee" if name == "f32ieee" else None)
        bc.pos.copy_(pos0.to(dt)); bc.prev.copy_(prev0.to(dt))
        runs[name] = run_schedule(bc, xy, delta, horizons)
    ref = runs["f64"]
    rows = []
    q = lambda x: [float(v) for v in np.percentile(x, [50, 90, 99, 100])]
    for h in sorted(ref[0]):
        row = {"substeps": h}
        for name in variants:
            mx, mn, dc = compare(runs[name][0][h], ref[0][h])
            row[name] = {"max_abs_dpos_p50_p90_p99_max": q(mx), "mean_abs_dpos_p50_p90_p99_max": q(mn), "abs_dcov_p50_p90_p99_max": q(dc)}
        if "f32" in variants and "f32ieee" in variants:
            mx, mn, dc = compare(runs["f32"][0][h], runs["f32ieee"][0][h])
            row["f32_vs_f32ieee"] = {"max_abs_dpos_p50_p90_p99_max": q(mx), "mean_abs_dpos_p50_p90_p99_max": q(mn), "abs_dcov_p50_p90_p99_max": q(dc)}
        row["mean_coverage"] = {k: float(runs[k][0][h][1].mean().item()) for k in runs}
        row["tear_or_bad"] = {k: int((runs[k][0][h][2] & (L.FLAG_TEAR | L.FLAG_BADSTATE) != 0).sum().item()) for k in runs}
        rows.append(row)
    same_grab = {k: bool(torch.equal(runs[k][1], ref[1]) and torch.equal(runs[k][2], ref[2])) for k in variants}
    return {"n_env": n, "seed": seed, "delta": list(delta), "schedule": ref[3], "grabbed_sets_equal_to_f64": same_grab,
            "grabbed_points_mean": float(ref[2].double().mean().item()), "rows": rows}


def markdown(res):
    out = ["| substeps | build | max\\|dpos\\| p50 / p99 / worst | mean\\|dpos\\| p50 / p99 | \\|dcov\\| p50 / p99 / worst | mean coverage f64 / build |",
           "|---|---|---|---|---|---|"]
    for r in res["rows"]:
        for name in [k for k in ("f32", "f32ieee", "f32_vs_f32ieee") if k in r]:
            v = r[name]
            mx, mn, dc = v["max_abs_dpos_p50_p90_p99_max"], v["mean_abs_dpos_p50_p90_p99_max"], v["abs_dcov_p50_p90_p99_max"]
            cov = r["mean_coverage"]
            out.append("| %d | %s | %.1e / %.1e / %.1e | %.1e / %.1e | %.1e / %.1e / %.1e | %.5f / %s |" % (
                r["substeps"], name, mx[0], mx[2], mx[3], mn[0], mn[2], dc[0], dc[2], dc[3], cov["f64"],
                "%.5f" % cov[name] if name in cov else "-"))
    return "\n".join(out)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    res = drift_study(n)
    print(markdown(res))
    print("grabbed sets equal to f64:", res["grabbed_sets_equal_to_f64"], "| schedule", res["schedule"])
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as fh:
            json.dump(res, fh, indent=1)
