"""Drive the reference's own compiled physics (oracle/_ref) through one ClothEnv.step (TEST/BASELINE
INFRASTRUCTURE; used by bench.py's cpu_baseline / --impl reference legs and by tests only).

gym_cloth/envs/cloth_env.py is pure Python and lives only in /root/reference, which does not exist on
the GPU box, so the ~40 lines of its step loop (cloth_env.py:401-515) are restated here around the
reference's real `Cloth.update()` / `Gripper` objects, which is where > 99 % of the time goes
(SURVEY.md §3.3).  tests/test_ref_driver.py checks this driver against fixtures recorded from the
reference ClothEnv itself.
"""
import math
import multiprocessing as mp
import os
import time

import numpy as np

_CFG = {"cloth": {"num_width_points": 25, "num_height_points": 25, "width": 1, "height": 1, "density": 200.0,
                  "ks": 10000.0, "damping": 2.0, "thickness": 0.02, "plane_friction": 1.0, "tear_thresh": 2.0,
                  "pin_cond": "y=0", "color_pts": "None"},
        "frames_per_sec": 30, "simulation_steps": 30, "seed": 1, "init": {"type": "tier1"}}
_ENV = {"iters_up": 50, "iters_up_rest": 80, "iters_grip_rest": 300, "iters_rest": 1000, "reduce_factor": 0.002,
        "grip_radius": 0.003}


def make_ref_cloth(pos=None, prev=None, rest=None):
    """A reference Cloth + Gripper pair, optionally loaded with a given state."""
    from oracle.ref_loader import load_physics
    Cloth, Gripper, _ = load_physics()
    c = Cloth(params=_CFG, render=False, random_state=np.random.RandomState(0))
    if pos is not None:
        for p, x, q in zip(c.pts, np.asarray(pos, np.float64).tolist(), np.asarray(prev, np.float64).tolist()):
            p.x, p.y, p.z = x
            p.px, p.py, p.pz = q
    if rest is not None:
        for s, r in zip(c.springs, np.asarray(rest, np.float64).tolist()):
            s.rest_length = r
    g = Gripper(c, _ENV["grip_radius"], _CFG["cloth"]["height"], _CFG["cloth"]["thickness"])
    return c, g


def ref_step(c, g, action):
    """cloth_env.py:401-515 (clip_act_space, delta_actions) on reference objects. Returns #updates."""
    x, y, dx, dy = (max(min(float(v), 1.0), -1.0) for v in action)
    x = (x / 2.0) + 0.5
    y = (y / 2.0) + 0.5
    g.grab_top(x, y)
    total = np.sqrt(dx ** 2 + dy ** 2)
    xr = dx / (total + 1e-5) * _ENV["reduce_factor"]
    yr = dy / (total + 1e-5) * _ENV["reduce_factor"]
    ii, cur = 0, 0
    while True:
        cur += np.sqrt(xr ** 2 + yr ** 2)
        if cur >= total:
            break
        ii += 1
    iu, iur, igr, ir = _ENV["iters_up"], _ENV["iters_up_rest"], _ENV["iters_grip_rest"], _ENV["iters_rest"]
    iterations = iu + iur + ii + igr + ir
    if len(g.grabbed_pts) == 0:
        iterations = 0
    i = 0
    n = 0
    while i < iterations:
        if i < iu:
            g.adjust(x=0.0, y=0.0, z=0.0025)
        elif i < iu + iur:
            pass
        elif i < iu + iur + ii:
            g.adjust(x=xr, y=yr, z=0.0)
        elif i < iu + iur + ii + igr:
            pass
        else:
            g.release()
        c.update()
        n += 1
        if c.have_tear:
            break
        i += 1
    return n


def ref_coverage(c):
    from scipy.spatial import ConvexHull
    pts = np.array([[min(max(p.x, 0), 1), min(max(p.y, 0), 1)] for p in c.pts])
    try:
        return ConvexHull(pts).volume
    except Exception:
        return 0


def _oob(pos):
    """ClothEnv._out_of_bounds (cloth_env.py:1020-1045): bounds (1,1,1), slack 0.25."""
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    return bool(((x >= 1.25) | (x < -0.25) | (y >= 1.25) | (y < -0.25) | (z >= 1) | (z < 0)).any())


def _worker(args):
    kind, pos, prev, action = args
    t0 = time.perf_counter()
    if kind == "reference":
        c, g = make_ref_cloth(pos, prev)
        n = ref_step(c, g, action)
        cov = ref_coverage(c)
        t1 = time.perf_counter()
        new_pos = np.array([[p.x, p.y, p.z] for p in c.pts]); new_prev = np.array([[p.px, p.py, p.pz] for p in c.pts])
        tear = bool(c.cloth_have_tear)
    else:
        from oracle.oracle import OracleCloth
        o = OracleCloth()
        o.set_state(pos, prev, np.zeros(len(pos), np.uint8))
        n, _, _ = o.step_action(np.asarray(action, np.float64))
        cov = o.coverage()
        t1 = time.perf_counter()
        st = o.get_state()
        new_pos, new_prev = st[0], st[1]
        tear = bool(o.tear)
    return n, cov, t1 - t0, new_pos, new_prev, tear, _oob(new_pos)


def cpu_env_steps(kind, states, actions, cores=None):
    """Run len(actions) env.step calls on `cores` host processes (one ClothEnv each, like the reference's
    'one CPU per analytic.py process' model, analysis/README.md:9-11).  states: list of (pos, prev).
    Returns dict(value env-steps/s, substeps/s, seconds, cores, n, and per call: coverage, the new (pos, prev), tear, oob)."""
    cores = cores or os.cpu_count() or 1
    jobs = [(kind, s[0], s[1], a) for s, a in zip(states, actions)]
    cores = max(1, min(cores, len(jobs)))
    t0 = time.perf_counter()
    if cores == 1:
        res = [_worker(j) for j in jobs]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            res = pool.map(_worker, jobs, chunksize=1)
    dt = time.perf_counter() - t0
    sub = sum(r[0] for r in res)
    return {"value": len(jobs) / dt, "substeps_per_s": sub / dt, "seconds": dt, "cores": cores, "n": len(jobs),
            "substeps": sub, "coverage": [r[1] for r in res], "states": [(r[3], r[4]) for r in res],
            "tear": [r[5] for r in res], "oob": [r[6] for r in res], "n_updates": [r[0] for r in res]}


# ------------------------------------------------------------------ episodes: one process per environment, K steps each
def _episode_worker(kind, i, pool_pos, pool_prev, start, raw, pick, choice, W, max_actions, barrier, q):
    """Environment i of bench.py's workload on one host core: W untimed + K timed env.step calls.  `raw[t]` is the drawn
    action, `pick[t]` the mesh point its grip is aimed at (None: use raw as it is), `choice[t]` the pool state it restarts
    from when the step ends its episode (ClothEnv._terminal, cloth_env.py:682-715: tear, out of bounds, coverage > 0.92,
    max_actions)."""
    pos, prev = pool_pos[start].copy(), pool_prev[start].copy()
    steps = 0
    rec = []
    t0 = None
    for t in range(len(raw)):
        if t == W:
            barrier.wait()
            t0 = time.perf_counter()
        a = np.array(raw[t], np.float64)
        if pick is not None:
            a[0] = (pos[pick[t], 0] - 0.5) * 2; a[1] = (pos[pick[t], 1] - 0.5) * 2
        r = _worker((kind, pos, prev, a))
        n, cov, dt, pos, prev, tear, oob = r
        steps += 1
        done = tear or oob or cov > 0.92 or steps >= max_actions
        if t >= W:
            rec.append((n, cov, dt, int(done), int(n == 0)))
        last = pos
        if done:
            pos, prev = pool_pos[choice[t]].copy(), pool_prev[choice[t]].copy()
            steps = 0
    t1 = time.perf_counter()
    q.put((i, t0, t1, rec, last if len(raw) == 1 else None))


def cpu_env_episodes(kind, pool_pos, pool_prev, starts, raw, pick, choice, warmup, cores=None, max_actions=10):
    """len(starts) environments, one host process each (the reference's own scaling model: one CPU per analytic.py
    process, analysis/README.md:9-11), each doing `warmup` untimed and raw.shape[0]-warmup timed env.step calls with
    no synchronisation between environments.  The clock runs from the moment every process has finished its warm-up to
    the moment the last one finishes.  raw [T, n, 4], pick [T, n] or None, choice [T, n]."""
    n = len(starts)
    cores = cores or os.cpu_count() or 1
    assert n <= cores, "one environment per core"
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(n)
    q = ctx.Queue()
    procs = [ctx.Process(target=_episode_worker, args=(kind, i, pool_pos, pool_prev, starts[i], raw[:, i], None if pick is None else pick[:, i],
                                                       choice[:, i], warmup, max_actions, barrier, q)) for i in range(n)]
    for p in procs:
        p.start()
    res = sorted(q.get() for _ in procs)
    for p in procs:
        p.join()
    t0 = min(r[1] for r in res); t1 = max(r[2] for r in res)
    K = raw.shape[0] - warmup
    rec = [r[3] for r in res]
    sub = sum(x[0] for e in rec for x in e)
    busy = sum(x[2] for e in rec for x in e)
    return {"value": n * K / (t1 - t0), "substeps_per_s": sub / (t1 - t0), "seconds": t1 - t0, "cores": n, "n": n * K, "steps": K,
            "substeps": sub, "core_busy_frac": busy / (n * (t1 - t0)), "coverage": [[x[1] for x in e] for e in rec],
            "n_updates": [[x[0] for x in e] for e in rec], "done_frac": float(np.mean([x[3] for e in rec for x in e])),
            "nograb_frac": float(np.mean([x[4] for e in rec for x in e])), "final_pos": [r[4] for r in res]}
