"""The CPU oracle side by side with the reference's own compiled Cython physics
(oracle/_ref, built by oracle/build_ref.py).  Skipped when oracle/_ref has not been built."""
import numpy as np
import pytest

from oracle.build_ref import ref_built
from oracle.oracle import OracleCloth, params_from_cfg

pytestmark = pytest.mark.skipif(not ref_built(), reason="oracle/_ref not built (python oracle/build_ref.py)")


def _ref_cfg():
    # cfg/t1_rgbd.yaml values; the reference Cloth reads only these keys (cloth.pyx:53-79, 175-186)
    return {"cloth": {"num_width_points": 25, "num_height_points": 25, "width": 1, "height": 1,
                      "density": 200.0, "ks": 10000.0, "damping": 2.0, "thickness": 0.02,
                      "plane_friction": 1.0, "tear_thresh": 2.0, "pin_cond": "y=0", "color_pts": "None"},
            "frames_per_sec": 30, "simulation_steps": 30, "seed": 1, "init": {"type": "tier1"}}


def _state(cloth):
    pos = np.array([[p.x, p.y, p.z] for p in cloth.pts])
    prev = np.array([[p.px, p.py, p.pz] for p in cloth.pts])
    pin = np.array([bool(p.pinned) for p in cloth.pts], np.uint8)
    return pos, prev, pin


def _same(c, g, o):
    a, b = _state(c), o.get_state()
    lut = {id(p): i for i, p in enumerate(c.pts)}
    return (all(np.array_equal(x, y) for x, y in zip(a, b))
            and [lut[id(p)] for p in g.grabbed_pts] == o.grabbed.tolist()
            and bool(c.have_tear) == o.tear)


@pytest.mark.parametrize("seed", [0, 1])
def test_random_schedule_bit_exact(seed):
    from oracle.ref_loader import load_physics
    Cloth, Gripper, _ = load_physics()
    cfg = _ref_cfg()
    rng = np.random.RandomState(seed)
    c = Cloth(params=cfg, render=False, random_state=np.random.RandomState(0))
    g = Gripper(c, 0.003, 1, 0.02)
    o = OracleCloth()
    for rnd in range(2):
        x, y = rng.uniform(0.1, 0.9, 2)
        g.grab_top(x, y); o.grab_top(x, y)
        assert _same(c, g, o)
        ang = rng.uniform(-np.pi, np.pi)
        d = (0.002 * np.cos(ang), 0.002 * np.sin(ang), 0.0)
        for k, adj in ((30, (0.0, 0.0, 0.0025)), (10, None), (60, d), (20, None)):
            for _ in range(k):
                if adj is not None:
                    g.adjust(*adj); o.adjust(*adj)
                c.update(); o.update()
            assert _same(c, g, o)
        g.release(); o.release()
        for _ in range(40):
            c.update(); o.update()
        assert _same(c, g, o)


def test_tier2_grid_and_rest_lengths():
    from oracle.ref_loader import load_physics
    Cloth, _, _ = load_physics()
    cfg = _ref_cfg(); cfg["init"]["type"] = "tier2"
    rs = np.random.RandomState(5)
    c = Cloth(params=cfg, render=False, random_state=rs)
    rs2 = np.random.RandomState(5)
    side = rs2.rand() > 0.5
    noise = np.array([rs2.rand() * 0.01 - 0.005 for _ in range(625)])
    o = OracleCloth(init_type="tier2", noise=noise, init_side=side)
    assert bool(c.init_side) == bool(side)
    pos, prev, _ = _state(c)
    opos, oprev, _ = o.get_state()
    assert np.array_equal(pos, opos) and np.array_equal(prev, oprev)
    assert np.array_equal(np.array([s.rest_length for s in c.springs]), o.springs()[3])
    for _ in range(25):
        c.update(); o.update()
    assert np.array_equal(_state(c)[0], o.pos)
