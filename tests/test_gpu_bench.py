"""bench.py's arms at small sizes on the GPU: every --config builds its batch and pool, steps through the device-resident
and the host-buffer arm with restarts, and the f64 arm of configs[1] reproduces what the CPU oracle does with the same
start state and the same aimed action."""
import argparse

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _args(**kw):
    d = dict(gpus=1, steps=1, warmup=1, impl="ours", config=2, envs=0, dtype="f32", mode="reference_order", actions="touch_cloth",
             relax_iters=2, seed=1337, no_cpu_baseline=True, no_extras=True, no_pairs=False)
    d.update(kw)
    return argparse.Namespace(**d)


def _sync():
    torch.cuda.synchronize()


@pytest.mark.parametrize("config,n,mode", [(2, 48, 0), (3, 32, 0), (4, 6, 1), (5, 40, 0)])
def test_arms_run_every_config(config, n, mode):
    import bench
    a = _args(config=config, envs=n)
    arm = bench.Arm(a, 0, n, "f32", mode)
    r = arm.run_device(2, 1, _sync)
    assert r["substeps"] > 0 and r["launches"] >= 2 and r["elapsed_ms"] > 0
    assert r["nograb"] <= n // 8                       # aimed grips catch the cloth (a grip on a point hidden under a fold may not)
    h = arm.run_host(2, 1, _sync)
    assert h["seconds"] > 0 and h["d2h"] == n * (3 * arm.np_ * 4 + 36)
    if mode == 0:
        p = arm.measure_pairs(77)
        assert p > 1000.0                              # a 25x25 cloth makes thousands of pair tests per substep
    c = arm.c
    assert torch.isfinite(c.pos).all() and int((c.flags & 8).sum().item()) == 0       # no BADSTATE


def test_f64_arm_step_equals_oracle():
    import bench
    from oracle.oracle import OracleCloth
    n = 6
    arm = bench.Arm(_args(envs=n), 0, n, "f64", 0)
    c = arm.c
    pos0 = c.pos[:, :, :3].cpu().numpy().copy(); prev0 = c.prev[:, :, :3].cpu().numpy().copy()
    drawn = arm.device_actions(0)
    acts = arm.device_step(0, drawn).cpu().numpy()
    _sync()
    for e in range(n):
        o = OracleCloth()
        o.set_state(pos0[e], prev0[e], np.zeros(625, np.uint8))
        # the device decode squares with x*x where CPython calls pow: iters_pull may differ by one for 0.085 % of the
        # actions; compare only when the plans agree
        nupd, ng, ip = o.step_action(np.clip(acts[e], -1, 1))
        if nupd != int(c.sim_steps[e].item()):
            continue
        assert np.array_equal(c.pos[e, :, :3].cpu().numpy(), o.get_state()[0]), e
        assert abs(float(c.coverage[e].item()) - o.coverage()) < 1e-12
