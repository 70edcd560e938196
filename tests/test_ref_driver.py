"""The reference-arm driver (oracle/ref_driver.py: the reference's compiled Cloth/Gripper inside a restated step loop)
against a fixture recorded from the reference ClothEnv itself, and the action generator's sharding invariance."""
import numpy as np
import pytest

from conftest import load_golden
from oracle.build_ref import ref_built


@pytest.mark.skipif(not ref_built(), reason="oracle/_ref not built")
def test_ref_driver_reproduces_env_step():
    from oracle.ref_driver import make_ref_cloth, ref_coverage, ref_step
    g = load_golden("env_t1_s1338.npz")
    c, grip = make_ref_cloth(g["pos_reset"], g["prev_reset"], g["rest"])
    n = ref_step(c, grip, g["actions"][0])
    assert n == int(g["infos"][0][1])
    pos = np.array([[p.x, p.y, p.z] for p in c.pts])
    assert np.array_equal(pos, g["pos_a0"])
    assert abs(ref_coverage(c) - g["infos"][0][2]) < 1e-15


def test_cpu_port_driver_and_action_generator():
    import bench
    from oracle.ref_driver import cpu_env_steps
    g = load_golden("env_t1_s1337.npz")
    a = bench.actions_for_step(7, 3, 0, 4096)
    b = np.concatenate([bench.actions_for_step(7, 3, 0, 1500), bench.actions_for_step(7, 3, 1500, 4096)])
    assert np.array_equal(a, b) and a.min() >= -1 and a.max() <= 1 and abs(a.mean()) < 0.02
    r = cpu_env_steps("port", [(g["pos_reset"], g["prev_reset"])] * 2, g["actions"][:1].repeat(2, 0), cores=2)
    assert r["n"] == 2 and r["substeps"] == 2 * int(g["infos"][0][1])
    assert abs(r["coverage"][0] - g["infos"][0][2]) < 1e-12


def test_episode_driver_matches_single_steps_and_restarts():
    """cpu_env_episodes (bench.py's reference arm): per-environment processes replay what chained single env.step calls
    give - aimed grips from the current state, restart from the pool when the episode ends."""
    import bench
    from oracle.ref_driver import cpu_env_episodes, cpu_env_steps
    d = np.load(bench.POOL_FIXTURE)
    pos, prev = d["pos"], d["prev"]
    n, T = 2, 2
    raw = np.stack([bench.draw_actions(1337, t, 0, n)[0] for t in range(T)])
    pick = np.stack([bench.draw_actions(1337, t, 0, n)[1] for t in range(T)])
    choice = np.stack([bench.restart_choice(1337, t, 0, n, len(pos)) for t in range(T)])
    r = cpu_env_episodes("port", pos, prev, [0, 1], raw, pick, choice, warmup=0, cores=2, max_actions=2)
    assert r["n"] == 4 and r["nograb_frac"] == 0.0 and r["done_frac"] >= 0.5       # max_actions=2: step 1 always ends the episode
    states = [(pos[i], prev[i]) for i in range(n)]
    for t in range(T):
        a = raw[t].copy()
        for i in range(n):
            a[i, :2] = (states[i][0][pick[t, i], :2] - 0.5) * 2
        s = cpu_env_steps("port", states, a, cores=2)
        assert s["n_updates"] == [r["n_updates"][i][t] for i in range(n)]
        assert np.allclose(s["coverage"], [r["coverage"][i][t] for i in range(n)], atol=0, rtol=0)
        states = list(s["states"])
        for i in range(n):                                     # ClothEnv._terminal (max_actions never binds at t = 0)
            if s["tear"][i] or s["oob"][i] or s["coverage"][i] > 0.92:
                states[i] = (pos[choice[t, i]], prev[choice[t, i]])
    assert r["substeps"] == sum(sum(e) for e in r["n_updates"]) and 0.5 < r["core_busy_frac"] <= 1.0
