"""The wrinkle policy's decision rule on the CPU (no device needed: the policy only reads env.cloth.pts): the states
the reference env was in and the actions the reference's WrinklesPolicy chose there (analytic.py:551-720)."""
import numpy as np

from conftest import load_golden


class _P(object):
    def __init__(self, xyz):
        self.x, self.y, self.z = (float(v) for v in xyz)


class _E(object):
    pass


def test_wrinkle_policy_actions_on_reference_states():
    from gym_cloth_b200.policies import WrinklesPolicy, _neighbors, _wrinkle_levels
    pol = WrinklesPolicy()
    for name in ("policy_wrinkle_t1_s1337.npz", "policy_wrinkle_t3_s1337.npz"):     # 1 action (episode over) + 3 actions
        g = load_golden(name)
        states = [g["pos_reset_e0"]] + [g["pos_%d" % k] for k in range(int(g["episode_lengths"][0]) - 1)]
        for k, pos in enumerate(states):
            env = _E(); env.cloth = _E(); env.cloth.pts = [_P(p) for p in pos]
            pol.set_env_cfg(env, {})
            a = np.array(pol.get_action(None, k), np.float64)
            assert np.array_equal(a, g["action_%d" % k]), (name, k, a, g["action_%d" % k])
    lv = _wrinkle_levels()
    assert len(lv) == 50 and lv[0] == 1 and lv[-1] > 0
    assert _neighbors(0, 0, 25, 25) == [0, 26, 25, 1] and len(_neighbors(12, 12, 25, 25)) == 9


def test_reveal_policy_actions_on_reference_states():
    """OracleCornerRevealPolicy (analytic.py:217-358): the reference's policy asked for an action on two reference states
    under each of the 16 occlusion vectors (tests/golden/policy_reveal_t3_s1337.npz) - ours must answer bit for bit."""
    from gym_cloth_b200.policies import OracleCornerRevealPolicy
    g = load_golden("policy_reveal_t3_s1337.npz")
    pol = OracleCornerRevealPolicy()
    cfg = {"env": {"delta_actions": True, "clip_act_space": True}, "init": {"type": "tier3"}}
    seen_signs = set()
    for pos, occ, act in zip(g["pos"], g["occlusion"], g["actions"]):
        env = _E(); env.cloth = _E(); env.cloth.pts = [_P(p) for p in pos]; env.cloth.init_side = True
        env._occlusion_vec = [bool(v) for v in occ]
        pol.set_env_cfg(env, cfg)
        a = np.array(pol.get_action(None, 0), np.float64)
        assert np.array_equal(a, act), (occ, a, act)
        seen_signs.add(pol._sign)
    assert seen_signs == {1, -1} and len(g["actions"]) == 32
