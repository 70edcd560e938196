#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE REFERENCE ITSELF.

Runs only in the build container: it imports the reference's compiled Cython physics
(oracle/_ref, built by oracle/build_ref.py from /root/reference/gym_cloth/physics/*.pyx)
and the reference's own pure-Python `ClothEnv` straight from /root/reference.  The
reference has no tests or golden vectors of its own, so these fixtures
are what pins the oracle (oracle/cloth_oracle.c) and, through it, the CUDA path.

    python tests/golden/make_golden.py all          # everything, in parallel (~3 min)
    python tests/golden/make_golden.py kat | phases | decode | tear
    python tests/golden/make_golden.py env --tier 1 --seed 1337 --actions 3

Fixtures (float64 unless noted; N=625):
  kat_appendix_d.npz   RNG-free schedule of SURVEY.md App. D: pos/prev at substeps 1,50,230,1530
  phases.npz           one Cloth.update() from a crumpled+gripped state, dumped after every phase
  decode.npz           action -> (grip x,y, per-substep delta, iters_pull, iterations) as the env computes them
  tear.npz             an action that tears the cloth (loop break, sticky flag)
  env_t{1,2,3}_s*.npz  reset() + K step(action) through the reference ClothEnv: states, reward,
                       done, info, grabbed index lists, iters_pull, and the np_random draw log
  bench_pool_t1.npz    64 tier-1 reset states (env seeds 1337..1400) from the reference's own reset(): start states of
                       bench.py's reference arm / cpu_baseline (not part of `all`: ~5 min)
  state_t1_s1337.pkl   the file the reference's ClothEnv.save_state wrote ({"pts": [Point], "springs": [Spring]})
  state_t1_s1337.npz   that state as arrays + reset()/step() of a second reference env started from the file
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"


def _cfg(tier):
    with open(os.path.join(REFERENCE, "cfg", "t%d_rgbd.yaml" % tier)) as fh:
        return yaml.safe_load(fh)


def _state(cloth):
    pos = np.array([[p.x, p.y, p.z] for p in cloth.pts])
    prev = np.array([[p.px, p.py, p.pz] for p in cloth.pts])
    pin = np.array([bool(p.pinned) for p in cloth.pts], np.uint8)
    return pos, prev, pin


def _idx(cloth, pts):
    lut = {id(p): i for i, p in enumerate(cloth.pts)}
    return np.array([lut[id(p)] for p in pts], np.int32)


def gen_kat(out):
    from oracle.ref_loader import load_physics
    from scipy.spatial import ConvexHull
    Cloth, Gripper, _ = load_physics()
    cfg = _cfg(1)
    c = Cloth(params=cfg, render=False, random_state=np.random.RandomState(0))
    g = Gripper(c, 0.003, 1, 0.02)
    g.grab_top(0.5, 0.5)
    d = {"grabbed": _idx(c, g.grabbed_pts)}
    n = 0

    def run(k, adj=None):
        nonlocal n
        for _ in range(k):
            if adj is not None:
                g.adjust(*adj)
            c.update()
            n += 1
            if n in (1, 50, 230, 530, 1530):
                pos, prev, pin = _state(c)
                d["pos_%d" % n] = pos; d["prev_%d" % n] = prev; d["pin_%d" % n] = pin

    run(50, (0, 0, 0.0025)); run(80); run(100, (0.002 * 0.6, 0.002 * 0.8, 0)); run(300)
    g.release(); run(1000)
    pts = np.array([[min(max(p.x, 0), 1), min(max(p.y, 0), 1)] for p in c.pts])
    d["coverage"] = ConvexHull(pts).volume
    d["tear"] = bool(c.have_tear)
    np.savez_compressed(os.path.join(out, "kat_appendix_d.npz"), **d)
    print("kat: coverage", d["coverage"])


def gen_phases(out):
    """One update split into the reference's own phase methods (cloth.pyx:188-207)."""
    from oracle.ref_loader import load_physics
    Cloth, Gripper, _ = load_physics()
    cfg = _cfg(1)
    c = Cloth(params=cfg, render=False, random_state=np.random.RandomState(0))
    g = Gripper(c, 0.003, 1, 0.02)
    # crumple: corner pull across the cloth, release, then grip again mid-pull so that
    # pinned points, collisions, plane reverts and stretched springs are all present.
    g.grab_top(0.95, 0.95)
    for _ in range(50): g.adjust(0, 0, 0.0025); c.update()
    for _ in range(30): c.update()
    for _ in range(280): g.adjust(-0.0014142, -0.0014142, 0.0); c.update()
    g.release()
    for _ in range(120): c.update()
    g.grab_top(0.55, 0.55)
    for _ in range(40): g.adjust(0, 0, 0.0025); c.update()
    for _ in range(25): g.adjust(0.0016, -0.0012, 0.0); c.update()
    d = {}
    d["pos_0"], d["prev_0"], d["pin_0"] = _state(c)
    d["grabbed"] = _idx(c, g.grabbed_pts)
    g.adjust(0.0016, -0.0012, 0.0)
    d["pos_adjust"], d["prev_adjust"], _ = _state(c)
    cp = cfg["cloth"]
    mass = cp["density"] / c.width / c.height
    c._reset_gravity(mass * c.gravity)
    c._hookes(cp["ks"])
    d["force"] = np.array([[p.fx, p.fy, p.fz] for p in c.pts])
    c._verlet(mass, 1.0 / cfg["frames_per_sec"] / cfg["simulation_steps"], cp["damping"])
    d["pos_verlet"], d["prev_verlet"], _ = _state(c)
    c.build_spatial_map()
    keys = sorted(c.map.keys())
    d["map_keys"] = np.array(keys, np.int64)
    d["map_sizes"] = np.array([len(c.map[k]) for k in keys], np.int32)
    d["map_members"] = np.concatenate([_idx(c, c.map[k]) for k in keys])
    for pt in c.pts:
        c.self_collide(pt, cfg["simulation_steps"], cp["thickness"])
    d["pos_collide"], _, _ = _state(c)
    for pt in c.pts:
        c._handle_plane_collision(pt, cp["plane_friction"], 0.0001)
    d["pos_plane"], _, _ = _state(c)
    c._limit_spring_changes(cp["tear_thresh"])
    d["pos_limit"], d["prev_limit"], _ = _state(c)
    d["tear"] = bool(c.have_tear)
    d["n_collide_moved"] = int((d["pos_collide"] != d["pos_verlet"]).any(1).sum())
    d["n_plane_moved"] = int((d["pos_plane"] != d["pos_collide"]).any(1).sum())
    d["n_limit_moved"] = int((d["pos_limit"] != d["pos_plane"]).any(1).sum())
    np.savez_compressed(os.path.join(out, "phases.npz"), **d)
    print("phases: moved collide/plane/limit", d["n_collide_moved"], d["n_plane_moved"], d["n_limit_moved"],
          "buckets", len(keys), "max", d["map_sizes"].max())


class _Recorder(object):
    """Proxy around env.np_random that logs every draw the reference makes."""

    def __init__(self, rng):
        self._rng = rng
        self.log = []

    def __getattr__(self, name):
        f = getattr(self._rng, name)
        if not callable(f):
            return f

        def wrapped(*a, **k):
            v = f(*a, **k)
            if np.ndim(v) == 0:
                self.log.append([name, [float(x) if not isinstance(x, int) else x for x in a],
                                 {kk: (float(vv) if np.ndim(vv) == 0 else list(np.shape(vv))) for kk, vv in k.items()},
                                 float(v)])
            else:
                self.log.append([name, [], {"shape": list(np.shape(v))}, None])
            return v
        return wrapped


def _make_env(tier, seed, tmp):
    from oracle.ref_loader import load_env
    ClothEnv = load_env(REFERENCE)
    cfg = _cfg(tier)
    cfg["env"]["obs_type"] = "1d"
    cfg["init"]["render_opengl"] = False
    cfg["log"]["level"] = "warning"
    cfg["log"]["file"] = os.path.join(tmp, "ref.log")
    path = os.path.join(tmp, "t%d.yaml" % tier)
    with open(path, "w") as fh:
        yaml.safe_dump(cfg, fh)
    env = ClothEnv(path)
    env.seed(seed)
    env._wd = env._hd = 224                # reference bug workaround, SURVEY.md App. B-1
    return env


def gen_env(out, tier, seed, n_actions, action_seed=None):
    tmp = tempfile.mkdtemp(prefix="golden_env_")
    env = _make_env(tier, seed, tmp)
    rec = _Recorder(env.np_random)
    env.np_random = rec
    t0 = time.time()
    hooks = {"grab_xy": [], "pull": None, "grabbed": []}

    def install_hooks():
        grip = env.gripper
        orig_grab = grip.grab_top

        def grab_top(x, y):
            orig_grab(x, y)
            hooks["grab_xy"].append((x, y))
            hooks["grabbed"].append(_idx(env.cloth, grip.grabbed_pts))
        grip.grab_top = grab_top

    orig_pull = env._pull

    def pull(i, iters_pull, xr, yr):
        hooks["pull"] = (iters_pull, xr, yr)
        return orig_pull(i, iters_pull, xr, yr)
    env._pull = pull

    # reset() builds Cloth and Gripper, then runs the tier's init actions through step();
    # record those init actions too by wrapping step
    init_actions = []
    orig_step = env.step

    def step(action, initialize=False):
        if initialize:
            init_actions.append([float(a) for a in action])
        return orig_step(action, initialize=initialize)
    env.step = step
    obs0 = env.reset()
    d = {"tier": tier, "seed": seed, "init_side": bool(env.cloth.init_side),
         "init_actions": np.array(init_actions, np.float64).reshape(-1, 4),
         "reset_iters_up": float(env.iters_up)}
    d["pos_reset"], d["prev_reset"], d["pin_reset"] = _state(env.cloth)
    d["obs_reset"] = np.asarray(obs0)
    d["start_coverage"] = env._start_coverage
    d["start_variance_inv"] = env._start_variance_inv
    d["tear_reset"] = bool(env.cloth.have_tear)
    d["rest"] = np.array([sp.rest_length for sp in env.cloth.springs])
    print("env t%d s%d: reset %.1fs start_cov %.4f" % (tier, seed, time.time() - t0, env._start_coverage))
    install_hooks()
    rng = np.random.RandomState(seed + 7919 if action_seed is None else action_seed)
    acts, rews, dones, infos = [], [], [], []
    for t in range(n_actions):
        a = tuple(float(v) for v in rng.uniform(-1, 1, size=4))
        if t % 2 == 0:
            # aim at an actual cloth point (clip space, cloth_env.py:1011-1015) so that something is gripped
            pt = env.cloth.pts[int(rng.randint(len(env.cloth.pts)))]
            a = (float((pt.x - 0.5) * 2), float((pt.y - 0.5) * 2), float(a[2] * 0.6), float(a[3] * 0.6))
        hooks["pull"] = None
        ng0 = len(hooks["grabbed"])
        obs, rew, done, info = env.step(a)
        acts.append(a); rews.append(float(rew)); dones.append(bool(done))
        infos.append([info["num_steps"], info["num_sim_steps"], info["actual_coverage"],
                      info["variance_inv"], float(info["have_tear"]), float(info["out_of_bounds"])])
        d["pos_a%d" % t], d["prev_a%d" % t], d["pin_a%d" % t] = _state(env.cloth)
        d["grabbed_a%d" % t] = hooks["grabbed"][ng0] if len(hooks["grabbed"]) > ng0 else np.zeros(0, np.int32)
        d["grab_xy_a%d" % t] = np.array(hooks["grab_xy"][-1])
        d["pull_a%d" % t] = np.array(hooks["pull"] if hooks["pull"] is not None else (-1, 0.0, 0.0), np.float64)
        print("  a%d %s -> rew %.4f done %s sim %d cov %.4f grabbed %d (%.1fs)" % (
            t, np.round(a, 3), rew, done, info["num_sim_steps"], info["actual_coverage"],
            len(d["grabbed_a%d" % t]), time.time() - t0))
        if info["have_tear"]:
            break
    d["actions"] = np.array(acts); d["rewards"] = np.array(rews); d["dones"] = np.array(dones)
    d["infos"] = np.array(infos)
    d["rng_log"] = json.dumps(rec.log)
    np.savez_compressed(os.path.join(out, "env_t%d_s%d.npz" % (tier, seed)), **d)


def gen_policy(out, tier=1, seed=1337, episodes=2, kind="oracle", max_t=None):
    """BASELINE config #1: the reference's own OracleCornerPolicy (examples/analytic.py:70-155) driving the reference
    ClothEnv, tier 1.  Records every action the policy chose and what the env returned.
    kind='highest': HighestPointPolicy (analytic.py:716-808), which draws from the global np.random (seeded here)."""
    sys.path.insert(0, os.path.join(REFERENCE, "examples"))
    tmp = tempfile.mkdtemp(prefix="golden_pol_")
    env = _make_env(tier, seed, tmp)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        import analytic
    policy = {"oracle": analytic.OracleCornerPolicy, "highest": analytic.HighestPointPolicy, "wrinkle": analytic.WrinklesPolicy}[kind]()
    policy.set_env_cfg(env, env.cfg)
    d = {"tier": tier, "seed": seed}
    np.random.seed(seed)
    ep_len = []
    k = 0
    for ep in range(episodes):
        obs = env.reset()
        d["pos_reset_e%d" % ep] = _state(env.cloth)[0]
        d["start_coverage_e%d" % ep] = env._start_coverage
        done = False; t = 0
        while not done and (max_t is None or t < max_t):
            with contextlib.redirect_stdout(io.StringIO()):
                a = policy.get_action(obs, t)
            obs, rew, done, info = env.step(a)
            d["action_%d" % k] = np.array([float(v) for v in a]); d["pos_%d" % k] = _state(env.cloth)[0]
            d["result_%d" % k] = np.array([rew, float(done), info["actual_coverage"], info["num_sim_steps"]])
            print("policy ep %d t %d: action %s cov %.4f rew %.3f done %s" % (ep, t, np.round(a, 3), info["actual_coverage"], rew, done))
            k += 1; t += 1
        ep_len.append(t)
    d["episode_lengths"] = np.array(ep_len)
    np.savez_compressed(os.path.join(out, "policy_%s_t%d_s%d.npz" % (kind, tier, seed)), **d)


def gen_policy_reveal(out, tier=3, seed=1337):
    """OracleCornerRevealPolicy (analytic.py:217-358) on reference states with the occlusion vectors Blender would
    supply set by hand (the policy only reads env._occlusion_vec): every one of the 16 vectors on two states."""
    sys.path.insert(0, os.path.join(REFERENCE, "examples"))
    tmp = tempfile.mkdtemp(prefix="golden_rev_")
    env = _make_env(tier, seed, tmp)
    import contextlib, io, itertools
    with contextlib.redirect_stdout(io.StringIO()):
        import analytic
    policy = analytic.OracleCornerRevealPolicy()
    policy.set_env_cfg(env, env.cfg)
    np.random.seed(seed)
    env.reset()
    d = {"tier": tier, "seed": seed}
    states, occs, acts = [], [], []
    for rep in range(2):
        pos = _state(env.cloth)[0]
        for occ in itertools.product([False, True], repeat=4):
            env._occlusion_vec = list(occ)
            with contextlib.redirect_stdout(io.StringIO()):
                a = policy.get_action(None, 0)
            states.append(pos); occs.append(occ); acts.append([float(v) for v in a])
        env._occlusion_vec = [True, True, True, True]
        with contextlib.redirect_stdout(io.StringIO()):
            a = policy.get_action(None, 0)
        env.step(a)
    d["pos"] = np.array(states); d["occlusion"] = np.array(occs); d["actions"] = np.array(acts)
    np.savez_compressed(os.path.join(out, "policy_reveal_t%d_s%d.npz" % (tier, seed)), **d)
    print("reveal policy: %d (state, occlusion) -> action samples" % len(acts))


def gen_state(out, tier=1, seed=1337):
    """save_state / start_state_path (cloth_env.py:343-350, 120-124, 736-741): the reference env pickles its
    {"pts", "springs"} after one action; a second reference env starts from that file, resets and steps."""
    import pickle
    tmp = tempfile.mkdtemp(prefix="golden_state_")
    env = _make_env(tier, seed, tmp)
    env.reset()
    rng = np.random.RandomState(seed + 17)
    pt = env.cloth.pts[int(rng.randint(len(env.cloth.pts)))]
    a1 = (float((pt.x - 0.5) * 2), float((pt.y - 0.5) * 2), 0.31, -0.22)
    env.step(a1)
    pkl = os.path.join(out, "state_t%d_s%d.pkl" % (tier, seed))
    env.save_state(pkl)
    d = {"tier": tier, "seed": seed}
    d["pos_saved"], d["prev_saved"], d["pin_saved"] = _state(env.cloth)
    d["orig_saved"] = np.array([[p.orig_x, p.orig_y, p.orig_z] for p in env.cloth.pts])
    d["rest_saved"] = np.array([sp.rest_length for sp in env.cloth.springs])
    # a fresh reference env started from the file
    from oracle.ref_loader import load_env
    ClothEnv = load_env(REFERENCE)
    env2 = ClothEnv(os.path.join(tmp, "t%d.yaml" % tier), start_state_path=pkl)
    env2.seed(seed + 1)
    env2._wd = env2._hd = 224
    obs0 = env2.reset()
    d["pos_reset"], d["prev_reset"], d["pin_reset"] = _state(env2.cloth)
    d["obs_reset"] = np.asarray(obs0); d["start_coverage"] = env2._start_coverage
    d["init_side"] = bool(env2.cloth.init_side)
    pt = env2.cloth.pts[int(rng.randint(len(env2.cloth.pts)))]
    a2 = (float((pt.x - 0.5) * 2), float((pt.y - 0.5) * 2), -0.28, 0.35)
    obs, rew, done, info = env2.step(a2)
    d["action"] = np.array(a2); d["reward"] = rew; d["done"] = done
    d["info"] = np.array([info["num_steps"], info["num_sim_steps"], info["actual_coverage"], info["variance_inv"],
                          float(info["have_tear"]), float(info["out_of_bounds"])])
    d["pos_a0"], d["prev_a0"], d["pin_a0"] = _state(env2.cloth)
    # second episode from the same file: the reference deep-copies the start state at every reset
    obs1 = env2.reset()
    d["pos_reset2"], _, _ = _state(env2.cloth)
    np.savez_compressed(os.path.join(out, "state_t%d_s%d.npz" % (tier, seed)), **d)
    print("state: saved %s (%d bytes), step from it: rew %.4f cov %.4f sim %d" % (
        pkl, os.path.getsize(pkl), rew, info["actual_coverage"], info["num_sim_steps"]))


def _bench_pool_one(args):
    seed, tmpdir = args
    env = _make_env(1, seed, tempfile.mkdtemp(prefix="golden_pool_", dir=tmpdir))
    env.reset()
    pos, prev, _ = _state(env.cloth)
    return seed, pos, prev, env._start_coverage


def gen_bench_pool(out, base_seed=1337, n=64):
    """Start states of bench.py's reference arm: the reference's own tier-1 reset() (cloth_env.py:717-891) for env seeds
    base_seed .. base_seed+n-1 - the seeds of the first n environments of the GPU arm, whose f64 reset is bit-equal."""
    import multiprocessing as mp
    tmp = tempfile.mkdtemp(prefix="golden_pool_")
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(_bench_pool_one, [(base_seed + i, tmp) for i in range(n)], chunksize=1)
    res.sort()
    np.savez_compressed(os.path.join(out, "bench_pool_t1.npz"), seeds=np.array([r[0] for r in res]),
                        pos=np.stack([r[1] for r in res]), prev=np.stack([r[2] for r in res]),
                        start_coverage=np.array([r[3] for r in res]))
    print("bench pool: %d tier-1 reset states, mean start coverage %.4f" % (n, np.mean([r[3] for r in res])))


def gen_decode(out):
    """Action decode exactly as ClothEnv.step computes it (cloth_env.py:401-475), captured by
    hooking gripper.grab_top and _pull with the physics update stubbed out."""
    tmp = tempfile.mkdtemp(prefix="golden_dec_")
    env = _make_env(1, 3, tmp)
    from oracle.ref_loader import load_physics
    Cloth, Gripper, _ = load_physics()
    env.cloth = Cloth(params=env.cfg, render=False, random_state=np.random.RandomState(0))
    env.gripper = Gripper(env.cloth, env.grip_radius, env.cfg["cloth"]["height"], env.cfg["cloth"]["thickness"])
    env.num_steps = env.num_sim_steps = 0
    env.have_tear = False
    env.cloth.update = lambda: None
    cap = {}

    def grab_top(x, y):
        cap["xy"] = (x, y)
        env.gripper.grabbed_pts = [env.cloth.pts[0]]   # pretend something is gripped
    env.gripper.grab_top = grab_top
    orig_pull = env._pull

    def pull(i, iters_pull, xr, yr):
        cap["pull"] = (iters_pull, xr, yr)
        cap["n"] = cap.get("n", 0) + 1
        env.gripper.grabbed_pts = []                   # adjust/release become no-ops
    env._pull = pull
    rng = np.random.RandomState(12345)
    acts = [tuple(float(v) for v in rng.uniform(-1, 1, size=4)) for _ in range(600)]
    acts += [(0.0, 0.0, 0.0, 0.0), (1.0, -1.0, 1.0, 1.0), (-1.0, 1.0, -1.0, 0.0), (0.3, 0.2, 0.0, 1e-3),
             (0.5, 0.5, 1e-9, 0.0), (0.1, -0.7, 0.002, 0.0), (0.1, -0.7, 0.0019, 0.0), (0.1, -0.7, 0.6, 0.8),
             (1.5, -2.0, 0.25, -0.5), (0.25, 0.75, -0.0625, 0.5)]
    rows = []
    for a in acts:
        cap.clear()
        env.step(a, initialize=True)
        ip, xr, yr = cap.get("pull", (0, 0.0, 0.0)) if "pull" in cap else (None, None, None)
        if ip is None:
            # iterations == 0 cannot happen here (a point is always "gripped"), so _pull ran at least once
            raise RuntimeError("no pull captured")
        rows.append([a[0], a[1], a[2], a[3], cap["xy"][0], cap["xy"][1], xr, yr, ip, cap["n"]])
    np.savez_compressed(os.path.join(out, "decode.npz"), table=np.array(rows, np.float64))
    print("decode: %d actions, iters_pull range %d..%d" % (len(rows), min(r[8] for r in rows), max(r[8] for r in rows)))


def gen_tear(out):
    """An action that tears a tier-1 cloth: a long fast pull (reduce_factor raised so each
    substep moves the gripped points by 0.02) through the reference ClothEnv loop."""
    tmp = tempfile.mkdtemp(prefix="golden_tear_")
    env = _make_env(1, 5, tmp)
    from oracle.ref_loader import load_physics
    Cloth, Gripper, _ = load_physics()
    env.cloth = Cloth(params=env.cfg, render=False, random_state=np.random.RandomState(0))
    env.gripper = Gripper(env.cloth, env.grip_radius, env.cfg["cloth"]["height"], env.cfg["cloth"]["thickness"])
    env.num_steps = env.num_sim_steps = 0
    env.have_tear = False
    env._prev_reward = env._compute_coverage()
    env._start_coverage = env._prev_reward
    env._start_variance_inv = env._compute_variance()
    env.reduce_factor = 0.02
    d = {"reduce_factor": 0.02}
    a = (0.0, 0.0, 0.9, 0.9)
    obs, rew, done, info = env.step(a)
    d["action"] = np.array(a)
    d["pos"], d["prev"], d["pin"] = _state(env.cloth)
    d["grabbed_after"] = _idx(env.cloth, env.gripper.grabbed_pts)
    d["reward"] = rew; d["done"] = done
    d["info"] = np.array([info["num_steps"], info["num_sim_steps"], info["actual_coverage"], info["variance_inv"],
                          float(info["have_tear"]), float(info["out_of_bounds"])])
    print("tear: have_tear", info["have_tear"], "sim_steps", info["num_sim_steps"], "rew", rew, "done", done,
          "still grabbed", len(d["grabbed_after"]))
    assert info["have_tear"]
    # a second action on the torn cloth: loop exits after one update
    a2 = (0.2, -0.3, -0.5, 0.1)
    obs, rew, done, info = env.step(a2)
    d["action2"] = np.array(a2)
    d["pos2"], d["prev2"], d["pin2"] = _state(env.cloth)
    d["grabbed_after2"] = _idx(env.cloth, env.gripper.grabbed_pts)
    d["info2"] = np.array([info["num_steps"], info["num_sim_steps"], info["actual_coverage"], info["variance_inv"],
                           float(info["have_tear"]), float(info["out_of_bounds"])])
    d["reward2"] = rew
    print("tear: second action sim_steps", info["num_sim_steps"], "grabbed", len(d["grabbed_after2"]))
    np.savez_compressed(os.path.join(out, "tear.npz"), **d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["all", "kat", "phases", "decode", "tear", "env", "policy", "policy_highest", "policy_wrinkle", "policy_reveal", "state", "bench_pool"])
    ap.add_argument("--tier", type=int, default=1)
    ap.add_argument("--seed", type=int, default=1337)
    ap.add_argument("--actions", type=int, default=3)
    ap.add_argument("--out", default=HERE)
    a = ap.parse_args()
    from oracle.build_ref import build
    if not build(REFERENCE):
        sys.exit("reference physics could not be built")
    if a.what == "all":
        jobs = [["kat"], ["phases"], ["decode"], ["tear"],
                ["env", "--tier", "1", "--seed", "1337", "--actions", "3"],
                ["env", "--tier", "1", "--seed", "1338", "--actions", "3"],
                ["env", "--tier", "2", "--seed", "1337", "--actions", "2"],
                ["env", "--tier", "3", "--seed", "1337", "--actions", "2"], ["policy"],
                ["policy_highest", "--tier", "3", "--seed", "1337"], ["state"]]
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__)] + j + ["--out", a.out]) for j in jobs]
        rc = [p.wait() for p in procs]
        print("exit codes", rc)
        sys.exit(max(rc))
    if a.what == "policy_highest":
        return gen_policy(a.out, tier=a.tier, seed=a.seed, episodes=1, kind="highest", max_t=3)
    if a.what == "policy_reveal":
        return gen_policy_reveal(a.out, tier=a.tier, seed=a.seed)
    if a.what == "policy_wrinkle":      # WrinklesPolicy (analytic.py:551-720): ground-truth state, no image, no RNG
        return gen_policy(a.out, tier=a.tier, seed=a.seed, episodes=1, kind="wrinkle", max_t=3)
    {"kat": gen_kat, "phases": gen_phases, "decode": gen_decode, "tear": gen_tear, "policy": gen_policy, "state": gen_state, "bench_pool": gen_bench_pool}.get(
        a.what, lambda out: gen_env(out, a.tier, a.seed, a.actions))(a.out)


if __name__ == "__main__":
    main()
