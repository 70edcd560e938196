"""GPU edge cases of the hot path: force_grab, batches where nothing is gripped, coincident points (the reference's
ZeroDivisionError), out-of-bounds flags, the per-environment rest-length table in f32 (tier 2), measure-only calls."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle.oracle import OracleCloth, params_from_cfg as oracle_params  # noqa: E402  (checker only)


def _L():
    from gym_cloth_b200 import lib
    return lib


def _bc(n, dtype, P=None, **kw):
    from gym_cloth_b200.batched import BatchedCloth
    return BatchedCloth(P if P is not None else _L().default_params(), n, dtype=dtype, **kw)


def test_force_grab_matches_reference_semantics():
    """cfg env.force_grab (cloth_env.py:434-444): the grip radius grows by 0.02 until something is gripped."""
    L = _L()
    g = load_golden("env_t1_s1337.npz")
    # fixture action a1 grips nothing on the state left by a0
    assert len(g["grabbed_a1"]) == 0
    o = OracleCloth()
    o.set_state(g["pos_a0"], g["prev_a0"], g["pin_a0"], grabbed=np.zeros(0, np.int32))
    n_ref, ng_ref, _ = o.step_action(g["actions"][1], force_grab=True)
    assert ng_ref > 0 and n_ref > 1400
    P = L.default_params(); P.force_grab = 1
    bc = _bc(1, torch.float64, P=P)
    bc.set_state(g["pos_a0"], g["prev_a0"], g["pin_a0"])
    out = {"sim_steps": np.zeros(1, np.int32), "flags": np.zeros(1, np.int32)}
    bc.step_host(g["actions"][1][None, :], out)
    assert out["sim_steps"][0] == n_ref and bc.n_grabbed[0].item() == ng_ref and not (out["flags"][0] & L.FLAG_NOGRAB)
    pos, prev, pin, _ = bc.get_state()
    op, oq, opin = o.get_state()
    assert np.array_equal(pos, op) and np.array_equal(prev, oq)


def test_nothing_gripped_leaves_state_and_costs_the_penalty():
    L = _L()
    g = load_golden("env_t1_s1337.npz")
    n = 5
    for dt in (torch.float32, torch.float64):
        bc = _bc(n, dt)
        bc.set_state(g["pos_a0"], g["prev_a0"], g["pin_a0"])
        bc.measure(); torch.cuda.synchronize(); bc.prev_coverage.copy_(bc.coverage)
        before = bc.pos.clone()
        acts = np.tile(g["actions"][1], (n, 1))
        out = {"sim_steps": np.zeros(n, np.int32), "flags": np.zeros(n, np.int32), "reward": np.zeros(n), "done": np.zeros(n, np.int32)}
        bc.step_host(acts, out)
        assert (out["sim_steps"] == 0).all() and ((out["flags"] & L.FLAG_NOGRAB) != 0).all()
        assert torch.equal(bc.pos, before)
        assert np.allclose(out["reward"], -0.01, atol=1e-15)          # cloth_env.py:564-566 (+0 coverage delta)
        assert bc.num_steps.cpu().tolist() == [1] * n and bc.num_sim_steps.cpu().tolist() == [0] * n


def test_coincident_points_raise_the_badstate_flag():
    """Two coincident, connected points: the reference raises ZeroDivisionError (cloth.pyx:232); we flag the environment."""
    L = _L()
    bc = _bc(2, torch.float32)
    p = bc.pos.clone()
    p[1, 313, :3] = p[1, 312, :3]
    bc.pos.copy_(p); bc.prev.copy_(p)
    bc.update(1); torch.cuda.synchronize()
    f = bc.flags.cpu().numpy()
    assert not (f[0] & L.FLAG_BADSTATE) and (f[1] & L.FLAG_BADSTATE)
    o = OracleCloth()
    pos = p[1, :, :3].double().cpu().numpy()
    o.set_state(pos, pos, np.zeros(625, np.uint8))
    with pytest.raises(ZeroDivisionError):
        o.update(1)


def test_out_of_bounds_and_measure_only():
    L = _L()
    bc = _bc(3, torch.float64)
    p = bc.pos.clone()
    p[1, 0, 0] = -0.2500001            # x < -0.25
    p[2, 624, 2] = 1.0                 # z >= 1
    bc.pos.copy_(p)
    before = bc.pos.clone()
    bc.measure(); torch.cuda.synchronize()
    assert torch.equal(bc.pos, before)                     # measure never writes the state
    f = bc.flags.cpu().numpy()
    assert [(int(x) & L.FLAG_OOB) != 0 for x in f] == [False, True, True]
    for e in range(3):
        o = OracleCloth()
        pos = bc.pos[e, :, :3].cpu().numpy()
        o.set_state(pos, pos, np.zeros(625, np.uint8))
        assert o.out_of_bounds() == bool(f[e] & L.FLAG_OOB)
        assert abs(o.coverage() - bc.coverage[e].item()) < 1e-12
        assert abs(o.variance_inv() - bc.variance_inv[e].item()) <= 1e-12 * abs(o.variance_inv())


def test_tier2_per_env_rest_table_f32_tracks_f64():
    """Tier-2 cloths carry per-cloth rest lengths (x-noise enters Spring.rest_length, cloth.pyx:101-108, 417):
    the f32 build reads them from the per-environment table and stays within the stated tolerance of f64."""
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import BatchedClothEnv
    a = BatchedClothEnv(cfg_path(2), 3, dtype="f32", seed=7)
    b = BatchedClothEnv(cfg_path(2), 3, dtype="f64", seed=7)
    a.reset(); b.reset()
    assert a.cloth.rest is not None and a.cloth.rest_env_stride == 6 * 625
    assert torch.equal(a.cloth.rest.double(), b.cloth.rest.float().double())
    d0 = (a.cloth.pos.double() - b.cloth.pos)[:, :, :3].abs()
    print("tier-2 reset (5482 substeps) f32 vs f64: max |dpos| %.3e mean %.3e" % (d0.max().item(), d0.mean().item()))
    a.cloth.pos.copy_(b.cloth.pos.float()); a.cloth.prev.copy_(b.cloth.prev.float())
    acts = np.random.RandomState(2).uniform(-1, 1, size=(3, 4))
    pidx = [300, 30, 600]
    xy = b.cloth.pos[torch.arange(3), torch.tensor(pidx), :2].cpu().numpy()
    acts[:, :2] = (xy - 0.5) * 2
    oa, ra, da, ia = a.step(acts)
    ob, rb, db, ib = b.step(acts)
    assert np.array_equal(ia["sim_steps"], ib["sim_steps"]) and (ia["sim_steps"] > 0).all()
    assert np.abs(oa - ob).max() <= 0.35 and np.abs(ia["actual_coverage"] - ib["actual_coverage"]).max() <= 5e-2
