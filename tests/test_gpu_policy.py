"""BASELINE config #1 (`examples/analytic.py oracle --tier=1`, single env) through this package: the facade ClothEnv in
f64 driven by our restated OracleCornerPolicy must reproduce, bit for bit, the episode the reference's own policy and
env produced (tests/golden/policy_oracle_t1_s1337.npz) - actions, states, rewards, termination - over two episodes
(the second one checks that the np_random stream stays aligned across resets).  Plus the batched policy."""
import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_oracle_policy_episodes_match_reference():
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import ClothEnv
    from gym_cloth_b200.policies import OracleCornerPolicy
    g = load_golden("policy_oracle_t1_s1337.npz")
    env = ClothEnv(cfg_path(1), dtype="f64")
    env.seed(int(g["seed"]))
    policy = OracleCornerPolicy()
    policy.set_env_cfg(env, env.cfg)
    k = 0
    for ep, length in enumerate(g["episode_lengths"]):
        obs = env.reset()
        assert np.array_equal(obs.reshape(-1, 3), g["pos_reset_e%d" % ep])
        assert abs(env._start_coverage - float(g["start_coverage_e%d" % ep])) < 1e-12
        for t in range(int(length)):
            a = policy.get_action(obs, t)
            assert np.array_equal(np.array(a, np.float64), g["action_%d" % k]), (ep, t)
            obs, rew, done, info = env.step(a)
            rew_ref, done_ref, cov_ref, sim_ref = g["result_%d" % k]
            assert np.array_equal(obs.reshape(-1, 3), g["pos_%d" % k])
            assert abs(rew - rew_ref) < 1e-11 and done == bool(done_ref)
            assert abs(info["actual_coverage"] - cov_ref) < 1e-12 and info["num_sim_steps"] == int(sim_ref)
            k += 1
        assert done            # the oracle policy flattens tier-1 cloths within a couple of actions


def test_batched_oracle_policy_matches_single_env_policy_and_improves_coverage():
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import BatchedClothEnv
    from gym_cloth_b200.policies import OracleCornerPolicy, oracle_corner_actions
    n = 64
    benv = BatchedClothEnv(cfg_path(1), n, dtype="f32", seed=500)
    benv.reset()
    acts = oracle_corner_actions(benv)
    assert acts.shape == (n, 4) and acts.dtype == torch.float32

    class _View(object):       # env.cloth.pts[i].x view of environment e, as the single-env policy expects
        def __init__(self, pos):
            class P(object):
                def __init__(s, xyz): s.x, s.y, s.z = (float(v) for v in xyz)
            self.pts = [P(p) for p in pos]
            self.init_side = True
    pol = OracleCornerPolicy()
    for e in (0, 7, 63):
        class E(object):
            pass
        env = E(); env.cloth = _View(benv.cloth.pos[e, :, :3].double().cpu().numpy())
        pol.set_env_cfg(env, benv.cfg)
        ref = np.array(pol.get_action(None, 0))
        assert np.allclose(acts[e].double().cpu().numpy(), ref, atol=1e-6)
    before = benv.start_coverage.mean().item()
    for t in range(3):
        obs, rew, done, info = benv.step(oracle_corner_actions(benv))
    torch.cuda.synchronize()
    after = info["actual_coverage"].mean().item()
    print("batched oracle policy: mean coverage %.3f -> %.3f after 3 actions (%d envs)" % (before, after, n))
    assert after > before + 0.05 and after > 0.9


def test_highest_point_policy_matches_reference():
    """HighestPointPolicy (analytic.py:716-808) on a tier-3 start: the reference draws the per-episode image
    randomisation values and the policy's point choice from the GLOBAL np.random, so the whole episode only matches
    if the facade consumes that stream in the reference's order, and `pt.orig_*` are the construction-time positions."""
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import ClothEnv
    from gym_cloth_b200.policies import HighestPointPolicy
    g = load_golden("policy_highest_t3_s1337.npz")
    env = ClothEnv(cfg_path(3), dtype="f64")
    env.seed(int(g["seed"]))
    policy = HighestPointPolicy()
    policy.set_env_cfg(env, env.cfg)
    np.random.seed(int(g["seed"]))
    obs = env.reset()
    assert np.array_equal(obs.reshape(-1, 3), g["pos_reset_e0"])
    for k in range(int(g["episode_lengths"][0])):
        a = policy.get_action(obs, k)
        assert np.array_equal(np.array(a, np.float64), g["action_%d" % k]), k
        obs, rew, done, info = env.step(a)
        rew_ref, done_ref, cov_ref, sim_ref = g["result_%d" % k]
        assert np.array_equal(obs.reshape(-1, 3), g["pos_%d" % k])
        assert abs(rew - rew_ref) < 1e-11 and done == bool(done_ref) and info["num_sim_steps"] == int(sim_ref)


def test_batched_highest_point_policy():
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import BatchedClothEnv
    from gym_cloth_b200.policies import highest_point_actions
    n = 32
    benv = BatchedClothEnv(cfg_path(3), n, dtype="f32", seed=77)
    benv.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    acts = highest_point_actions(benv, generator=gen)
    assert acts.shape == (n, 4)
    pos = benv.cloth.pos
    z = pos[:, :, 2]
    a = acts.double().cpu().numpy()
    for e in range(n):
        # the grabbed location is one of the five highest points, the pull goes 90 % of the way to its flat-grid place
        xy = a[e, :2] / 2.0 + 0.5
        top = torch.sort(z[e], descending=True, stable=True).indices[:5].cpu().numpy()
        pts = pos[e, top, :2].double().cpu().numpy()
        j = np.argmin(np.abs(pts - xy).sum(1))
        assert np.abs(pts[j] - xy).max() < 1e-6
        targ = benv.orig_pos[e, top[j], :2].cpu().numpy()
        assert np.allclose(a[e, 2:], (targ - pts[j]) * 0.9, atol=1e-6)
    cov0 = benv.start_coverage.mean().item()
    for t in range(4):
        obs, rew, done, info = benv.step(highest_point_actions(benv, generator=gen))
    assert info["actual_coverage"].mean().item() > cov0


def test_wrinkle_policy_matches_reference():
    """WrinklesPolicy (analytic.py:551-720): the reference's own policy driving the reference env for three actions
    (tests/golden/policy_wrinkle_t1_s1337.npz); ours must choose the same actions bit for bit and the f64 facade must
    land in the same states."""
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import ClothEnv
    from gym_cloth_b200.policies import WrinklesPolicy
    for tier, name in ((1, "policy_wrinkle_t1_s1337.npz"), (3, "policy_wrinkle_t3_s1337.npz")):
        g = load_golden(name)
        env = ClothEnv(cfg_path(tier), dtype="f64")
        env.seed(int(g["seed"]))
        np.random.seed(int(g["seed"]))                 # the generator seeds the global stream before reset (dom-rand draws)
        policy = WrinklesPolicy()
        policy.set_env_cfg(env, env.cfg)
        obs = env.reset()
        assert np.array_equal(obs.reshape(-1, 3), g["pos_reset_e0"])
        for k in range(int(g["episode_lengths"][0])):
            a = policy.get_action(obs, k)
            assert np.array_equal(np.array(a, np.float64), g["action_%d" % k]), (tier, k, a, g["action_%d" % k])
            obs, rew, done, info = env.step(a)
            rew_ref, done_ref, cov_ref, sim_ref = g["result_%d" % k]
            assert np.array_equal(obs.reshape(-1, 3), g["pos_%d" % k])
            assert abs(rew - rew_ref) < 1e-11 and done == bool(done_ref) and info["num_sim_steps"] == int(sim_ref)


def test_batched_wrinkle_policy_matches_single_env_policy():
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import BatchedClothEnv
    from gym_cloth_b200.policies import WrinklesPolicy, wrinkle_actions
    n = 48
    benv = BatchedClothEnv(cfg_path(1), n, dtype="f64", seed=900)
    benv.reset()
    acts = wrinkle_actions(benv, chunk=20)
    assert acts.shape == (n, 4) and acts.dtype == torch.float64

    class _P(object):
        def __init__(s, xyz): s.x, s.y, s.z = (float(v) for v in xyz)

    class _E(object):
        pass
    pol = WrinklesPolicy()
    for e in (0, 5, 19, 20, 47):
        env = _E(); env.cloth = _E(); env.cloth.pts = [_P(p) for p in benv.cloth.pos[e, :, :3].cpu().numpy()]
        pol.set_env_cfg(env, benv.cfg)
        ref = np.array(pol.get_action(None, 0), np.float64)
        assert np.allclose(acts[e].cpu().numpy(), ref, atol=1e-9), (e, acts[e], ref)
    benv.step(acts)                                   # the actions are valid env.step input
    torch.cuda.synchronize()
    assert int(((benv.cloth.flags & 4) != 0).sum().item()) <= n // 4        # most grips catch the cloth edge point aimed at


def test_batched_reveal_policy_matches_single_env_policy():
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import BatchedClothEnv
    from gym_cloth_b200.policies import OracleCornerRevealPolicy, oracle_corner_reveal_actions
    n = 32
    benv = BatchedClothEnv(cfg_path(3), n, dtype="f64", seed=77)
    benv.reset()
    rng = np.random.RandomState(3)
    occ = rng.rand(n, 4) < 0.5
    occ[0] = True; occ[1] = False                       # all occluded (the 1-D observation default) / all visible
    acts = oracle_corner_reveal_actions(benv, occ)
    assert acts.shape == (n, 4) and acts.dtype == torch.float64

    class _P(object):
        def __init__(s, xyz): s.x, s.y, s.z = (float(v) for v in xyz)

    class _E(object):
        pass
    pol = OracleCornerRevealPolicy()
    for e in range(n):
        env = _E(); env.cloth = _E(); env.cloth.pts = [_P(p) for p in benv.cloth.pos[e, :, :3].cpu().numpy()]
        env.cloth.init_side = True; env._occlusion_vec = [bool(v) for v in occ[e]]
        pol.set_env_cfg(env, benv.cfg)
        ref = np.array(pol.get_action(None, 0), np.float64)
        assert np.allclose(acts[e].cpu().numpy(), ref, atol=1e-12), (e, occ[e], acts[e], ref)
    benv.step(acts)
    torch.cuda.synchronize()
