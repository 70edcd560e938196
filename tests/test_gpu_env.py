"""GPU tests of the reference-shaped environment layer (gym_cloth_b200.envs): the single-env ClothEnv facade in
f64 must reproduce the reference ClothEnv's reset() and step() bit for bit from the same seed, because it
draws from np.random.RandomState in the reference's order and runs the same arithmetic on the device."""
import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tier,seed", [(1, 1337), (1, 1338), (2, 1337), (3, 1337)])
def test_facade_reset_and_steps_match_reference(tier, seed):
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import ClothEnv
    g = load_golden("env_t%d_s%d.npz" % (tier, seed))
    env = ClothEnv(cfg_path(tier), dtype="f64")
    env.seed(seed)
    obs = env.reset()
    ia = np.array(env._b.init_actions[0])
    assert np.array_equal(ia, g["init_actions"]), (ia, g["init_actions"])
    assert env.cloth.init_side == bool(g["init_side"])
    assert np.array_equal(obs.reshape(-1, 3), g["pos_reset"])
    assert np.array_equal(env.cloth._host()[1], g["prev_reset"])
    assert abs(env._start_coverage - float(g["start_coverage"])) < 1e-12
    p = env.cloth.pts[26]
    assert (p.x, p.y, p.z) == tuple(g["pos_reset"][26])
    for t, a in enumerate(g["actions"]):
        obs, rew, done, info = env.step(tuple(a))
        assert np.array_equal(obs.reshape(-1, 3), g["pos_a%d" % t])
        steps, sim, cov, vinv, tear, oob = g["infos"][t]
        assert (info["num_steps"], info["num_sim_steps"]) == (int(steps), int(sim))
        assert abs(info["actual_coverage"] - cov) < 1e-12 and abs(rew - g["rewards"][t]) < 1e-11
        assert done == bool(g["dones"][t]) and info["have_tear"] == bool(tear) and info["out_of_bounds"] == bool(oob)


def test_facade_gripper_and_update_calls():
    """env.gripper.grab_top / adjust / release and env.cloth.update() used directly, as in App. D."""
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import ClothEnv
    g = load_golden("kat_appendix_d.npz")
    env = ClothEnv(cfg_path(1), dtype="f64")
    env._b.cloth.reset_grid("tier1")
    env.gripper.grab_top(0.5, 0.5)
    assert sorted(env.cloth.pts.index(p) for p in env.gripper.grabbed_pts) == g["grabbed"].tolist()
    for _ in range(50):
        env.gripper.adjust(0, 0, 0.0025); env.cloth.update()
    assert np.array_equal(env.cloth.allpts_arr, g["pos_50"])
    assert env.cloth.pts[312].pinned and not env.cloth.pts[0].pinned


def test_batched_env_matches_single_envs_and_sharding():
    """Environment i of a batch seeded s equals a single env seeded s+i, wherever the batch is split."""
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import BatchedClothEnv
    full = BatchedClothEnv(cfg_path(1), 6, dtype="f64", seed=100)
    full.reset()
    lo = BatchedClothEnv(cfg_path(1), 3, dtype="f64", seed=100, env_offset=0)
    hi = BatchedClothEnv(cfg_path(1), 3, dtype="f64", seed=100, env_offset=3)
    lo.reset(); hi.reset()
    assert torch.equal(full.cloth.pos[:3], lo.cloth.pos) and torch.equal(full.cloth.pos[3:], hi.cloth.pos)
    acts = np.random.RandomState(5).uniform(-1, 1, size=(6, 4))
    o, r, d, info = full.step(acts)
    o1, r1, d1, _ = lo.step(acts[:3]); o2, r2, d2, _ = hi.step(acts[3:])
    assert np.array_equal(o[:3], o1) and np.array_equal(o[3:], o2)
    assert np.array_equal(r, np.concatenate([r1, r2])) and np.array_equal(d, np.concatenate([d1, d2]))
    # partial reset touches only the selected environments
    before = full.cloth.pos.clone()
    full.reset(envs=[1, 4])
    same = (full.cloth.pos == before).flatten(1).all(1).cpu().numpy()
    assert same.tolist() == [True, False, True, True, False, True]


def test_start_state_path_reads_the_reference_pickle_and_steps_like_the_reference(tmp_path):
    """cloth_env.py:120-124, 736-741, 343-350: a ClothEnv started from the file the REFERENCE's save_state wrote resets to
    that state, steps from it exactly like a reference env started from the same file, re-resets to it, and writes a
    file that reads back to the same arrays."""
    import os
    from gym_cloth_b200 import cfg_path, state_io
    from gym_cloth_b200.envs import ClothEnv
    from conftest import GOLDEN
    g = load_golden("state_t1_s1337.npz")
    env = ClothEnv(cfg_path(1), start_state_path=os.path.join(GOLDEN, "state_t1_s1337.pkl"), dtype="f64")
    env.seed(int(g["seed"]) + 1)
    obs = env.reset()
    assert np.array_equal(obs.reshape(-1, 3), g["pos_reset"]) and np.array_equal(env.cloth._host()[1], g["prev_reset"])
    assert env.cloth.init_side == bool(g["init_side"])
    assert abs(env._start_coverage - float(g["start_coverage"])) < 1e-12
    assert env.cloth.pts[30].orig_x == g["orig_saved"][30, 0]
    obs, rew, done, info = env.step(tuple(g["action"]))
    assert np.array_equal(obs.reshape(-1, 3), g["pos_a0"]) and np.array_equal(env.cloth._host()[1], g["prev_a0"])
    steps, sim, cov, vinv, tear, oob = g["info"]
    assert (info["num_steps"], info["num_sim_steps"]) == (int(steps), int(sim))
    assert abs(info["actual_coverage"] - cov) < 1e-12 and abs(rew - float(g["reward"])) < 1e-11 and done == bool(g["done"])
    out = str(tmp_path / "mine.pkl")
    env.save_state(out)
    a = state_io.state_to_arrays(state_io.load_state(out), 25)
    assert np.array_equal(a["pos"], g["pos_a0"]) and np.array_equal(a["prev"], g["prev_a0"])
    assert np.array_equal(a["orig"], g["orig_saved"])
    from gym_cloth_b200.batched import spring_slots
    assert np.array_equal(a["rest"][spring_slots(25)], g["rest_saved"])
    obs = env.reset()                                   # every reset starts from a copy of the file's state
    assert np.array_equal(obs.reshape(-1, 3), g["pos_reset2"])


def test_reward_type_coverage():
    """cfg env.reward_type 'coverage' (cloth_env.py:657-659): the reward is the coverage itself (+ the same bonuses and
    penalties), 'coverage-delta' (:660-662) its change; states and termination are the same."""
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import BatchedClothEnv
    from gym_cloth_b200.envs.cloth_env import load_cfg
    cfg = load_cfg(cfg_path(1))
    acts = np.array([[0.1, 0.2, 0.3, -0.2], [-0.4, 0.3, -0.2, 0.25], [0.99, 0.99, 0.1, 0.1]])
    out = {}
    for rt in ("coverage-delta", "coverage"):
        cfg["env"]["reward_type"] = rt
        env = BatchedClothEnv(cfg, 3, dtype="f64", seed=5)
        env.reset()
        start = env.start_coverage.cpu().numpy().copy()
        _, rew, done, info = env.step(acts)
        out[rt] = (np.array(rew), np.array(info["actual_coverage"]), np.array(info["no_grab"]), start, np.array(done))
    d, c = out["coverage-delta"], out["coverage"]
    assert np.array_equal(d[1], c[1]) and np.array_equal(d[4], c[4])
    pen = np.where(c[2], -0.01, 0.0) + np.where(c[1] > 0.92, 5.0, 0.0)
    assert np.allclose(c[0], pen + c[1], atol=1e-15)
    assert np.allclose(d[0], pen + (d[1] - d[3]), atol=1e-15)
    cfg["env"]["reward_type"] = "height"
    with pytest.raises(AssertionError):
        BatchedClothEnv(cfg, 1)


def test_nan_action_does_not_hang():
    """A NaN action component survives np.clip and the reference's iters_pull loop (cloth_env.py:462-467) never ends;
    here that environment does nothing and reports BADSTATE | NOGRAB, through both decode paths."""
    from gym_cloth_b200 import lib as L
    from gym_cloth_b200.batched import BatchedCloth
    acts = np.array([[0.0, 0.0, 0.2, 0.1], [0.1, np.nan, 0.2, 0.1], [0.3, 0.3, np.nan, np.nan], [np.inf, -np.inf, 0.1, 0.1]])
    for dt in (torch.float32, torch.float64):
        bc = BatchedCloth(L.default_params(), 4, dtype=dt)
        before = bc.pos.clone()
        bc.step_actions(torch.from_numpy(acts).to("cuda", dt).contiguous())
        torch.cuda.synchronize()
        f = bc.flags.cpu().numpy(); s = bc.sim_steps.cpu().numpy()
        assert s[0] > 1400 and s[1] == 0 and s[2] == 0
        assert (f[1] & L.FLAG_BADSTATE) and (f[1] & L.FLAG_NOGRAB) and (f[2] & L.FLAG_BADSTATE) and not (f[0] & L.FLAG_BADSTATE)
        assert not (f[3] & L.FLAG_BADSTATE)                     # infinities are clipped like any other value
        assert torch.equal(bc.pos[1], before[1])
        bc2 = BatchedCloth(L.default_params(), 4, dtype=dt)
        out = {"flags": np.zeros(4, np.int32), "sim_steps": np.zeros(4, np.int32)}
        bc2.step_host(acts, out)
        assert np.array_equal(out["flags"], f) and np.array_equal(out["sim_steps"], s)
