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
