"""The CPU oracle (oracle/cloth_oracle.c) against the golden fixtures that the reference
itself produced (tests/golden/make_golden.py).  Bit-exact for positions / indices / counts;
coverage to 1e-12 (Qhull vs monotone chain), variance to 1e-12 relative (numpy pairwise sum)."""
import json

import numpy as np
import pytest

from conftest import load_golden
from oracle.oracle import OracleCloth, OraclePlan, hull_area, lib, params_from_cfg
import ctypes as C


def _eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b))


def test_kat_appendix_d():
    g = load_golden("kat_appendix_d.npz")
    o = OracleCloth()
    assert o.N == 625 and o.S == 3502
    o.grab_top(0.5, 0.5)
    assert o.grabbed.tolist() == g["grabbed"].tolist() == [287, 311, 312, 313, 337]
    n = 0

    def run(k, adj=None):
        nonlocal n
        for _ in range(k):
            if adj is not None:
                o.adjust(*adj)
            o.update()
            n += 1
            if "pos_%d" % n in g.files:
                pos, prev, pin = o.get_state()
                assert _eq(pos, g["pos_%d" % n]), n
                assert _eq(prev, g["prev_%d" % n]), n
                assert _eq(pin, g["pin_%d" % n]), n

    run(50, (0, 0, 0.0025)); run(80); run(100, (0.002 * 0.6, 0.002 * 0.8, 0)); run(300)
    o.release(); run(1000)
    assert not o.tear and not bool(g["tear"])
    assert abs(o.coverage() - float(g["coverage"])) < 1e-12
    assert abs(float(g["coverage"]) - 0.8181217075228853) < 1e-15


def test_phases_one_update():
    g = load_golden("phases.npz")
    o = OracleCloth()
    o.set_state(g["pos_0"], g["prev_0"], g["pin_0"], grabbed=g["grabbed"])
    o.adjust(0.0016, -0.0012, 0.0)
    pos, prev, _ = o.get_state()
    assert _eq(pos, g["pos_adjust"]) and _eq(prev, g["prev_adjust"])
    o.phase("gravity"); o.phase("hookes")
    assert _eq(o.force(), g["force"])
    o.phase("verlet")
    pos, prev, _ = o.get_state()
    assert _eq(pos, g["pos_verlet"]) and _eq(prev, g["prev_verlet"])
    o.phase("build_map"); o.phase("self_collide")
    assert _eq(o.pos, g["pos_collide"])
    o.phase("plane")
    assert _eq(o.pos, g["pos_plane"])
    o.phase("limit")
    pos, prev, _ = o.get_state()
    assert _eq(pos, g["pos_limit"]) and _eq(prev, g["prev_limit"])
    assert o.tear == bool(g["tear"])
    # the fixture exercises every branch
    assert int(g["n_collide_moved"]) > 0 and int(g["n_plane_moved"]) > 0 and int(g["n_limit_moved"]) > 0


def test_decode_matches_env():
    t = load_golden("decode.npz")["table"]
    P = params_from_cfg(None)
    L = lib()
    plan = OraclePlan()
    for row in t:
        L.oracle_decode_action(C.byref(P), np.ascontiguousarray(row[:4]), C.byref(plan))
        assert (plan.gx, plan.gy) == (row[4], row[5])
        assert (plan.dxr, plan.dyr) == (row[6], row[7])
        assert plan.iters_pull == int(row[8])
        assert int(row[9]) == 50 + 80 + int(row[8]) + 300 + 1000  # iterations, cloth_env.py:475
    assert t[:, 8].max() == 707 and t[:, 8].min() == 0


def test_tear_breaks_loop_and_sticks():
    g = load_golden("tear.npz")
    P = params_from_cfg(None)
    P.reduce_factor = float(g["reduce_factor"])
    o = OracleCloth(P)
    n, ng, ip = o.step_action(g["action"])
    info = g["info"]
    assert n == int(info[1]) and o.tear
    pos, prev, pin = o.get_state()
    assert _eq(pos, g["pos"]) and _eq(prev, g["prev"]) and _eq(pin, g["pin"])
    assert o.grabbed.tolist() == g["grabbed_after"].tolist()   # not released on tear
    assert abs(o.coverage() - info[2]) < 1e-12
    n2, _, _ = o.step_action(g["action2"])
    assert n2 == 1 and n + n2 == int(g["info2"][1])
    pos, prev, pin = o.get_state()
    assert _eq(pos, g["pos2"]) and _eq(prev, g["prev2"]) and _eq(pin, g["pin2"])
    assert o.grabbed.tolist() == g["grabbed_after2"].tolist()


@pytest.mark.parametrize("name", ["env_t1_s1337.npz", "env_t1_s1338.npz", "env_t2_s1337.npz", "env_t3_s1337.npz"])
def test_env_steps(name):
    """reset-state -> K x step(action) of the reference ClothEnv, replayed by the oracle."""
    g = load_golden(name)
    o = OracleCloth()
    o.set_rest(g["rest"])   # tier-2 rest lengths carry the per-cloth x-noise (cloth.pyx:101-108, 417)
    o.set_state(g["pos_reset"], g["prev_reset"], g["pin_reset"], grabbed=np.zeros(0, np.int32),
                tear=bool(g["tear_reset"]))
    assert abs(o.coverage() - float(g["start_coverage"])) < 1e-12
    assert abs(o.variance_inv() - float(g["start_variance_inv"])) <= 1e-12 * abs(float(g["start_variance_inv"]))
    prev_cov = float(g["start_coverage"])
    sim = 0
    for t, a in enumerate(g["actions"]):
        plan = o.decode(a)
        assert (plan.gx, plan.gy) == tuple(g["grab_xy_a%d" % t])
        n, ng, ip = o.step_action(a)
        sim += n
        steps, sim_ref, cov, var_inv, tear, oob = g["infos"][t]
        assert sim == int(sim_ref)
        pull = g["pull_a%d" % t]
        if ng:
            assert ip == int(pull[0]) and (plan.dxr, plan.dyr) == (pull[1], pull[2])
        pos, prev, pin = o.get_state()
        assert _eq(pos, g["pos_a%d" % t]) and _eq(prev, g["prev_a%d" % t]) and _eq(pin, g["pin_a%d" % t])
        assert ng == len(g["grabbed_a%d" % t])
        c = o.coverage()
        assert abs(c - cov) < 1e-12
        assert abs(o.variance_inv() - var_inv) <= 1e-12 * abs(var_inv)
        assert o.tear == bool(tear) and o.out_of_bounds() == bool(oob)
        # reward, cloth_env.py:536-682 (coverage-delta)
        rew = (-0.01 if ng == 0 else 0.0) + (5.0 if c > 0.92 else 0.0) + (c - prev_cov)
        prev_cov = c
        assert abs(rew - g["rewards"][t]) < 1e-11
        done = (steps >= 10) or bool(tear) or bool(oob) or (c > 0.92)
        assert done == bool(g["dones"][t])


def test_hull_area_vs_scipy():
    from scipy.spatial import ConvexHull
    rng = np.random.RandomState(0)
    for n in (3, 4, 10, 625):
        for _ in range(20):
            pts = np.clip(rng.normal(0.5, 0.4, size=(n, 2)), 0, 1)   # many clipped/duplicated points
            assert abs(hull_area(pts) - ConvexHull(pts).volume) < 1e-12
    # flat cloth grid: area exactly 1
    gx, gy = np.meshgrid(np.arange(25) / 24.0, np.arange(25) / 24.0, indexing="ij")
    assert abs(hull_area(np.stack([gx.ravel(), gy.ravel()], 1)) - 1.0) < 1e-15
    # degenerate: collinear -> 0 (the reference's QhullError branch, cloth_env.py:634-637)
    assert hull_area(np.stack([np.linspace(0, 1, 50), np.linspace(0, 1, 50)], 1)) == 0.0
    assert hull_area(np.zeros((625, 2))) == 0.0
