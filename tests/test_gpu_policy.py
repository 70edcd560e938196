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
