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
