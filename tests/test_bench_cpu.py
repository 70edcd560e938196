"""Host logic of bench.py that needs no GPU: both arms describe the same workload with the same `config`, the roofline
bookkeeping is SURVEY.md 8(d)'s, and the action / restart draws do not depend on how the batch is split."""
import argparse

import numpy as np


def _args(**kw):
    d = dict(gpus=1, steps=5, warmup=3, impl="ours", config=2, envs=0, dtype="f32", mode="reference_order", actions="touch_cloth",
             relax_iters=2, seed=1337, no_cpu_baseline=False, no_extras=False, no_pairs=False)
    d.update(kw)
    return argparse.Namespace(**d)


def test_both_arms_print_the_same_config():
    import bench
    for world in (1, 2, 8):
        a = _args()
        wl, scaling, n, w = bench.workload(a, world)
        ours = bench.config_dict(a, wl, n, n * world, w, False)
        b = _args(impl="reference")
        wl2, scaling2, n2, w2 = bench.workload(b, world)
        ref = bench.config_dict(b, wl2, n2, n2 * world, w2, False)
        assert ours == ref and scaling == scaling2 == "weak" and n == 4096 and ours["actions"] == "touch_cloth"
    wl, scaling, n, w = bench.workload(_args(config=5), 8)
    assert scaling == "strong" and n == 8192 and w == 25
    assert bench.workload(_args(config=3), 1)[2] == 16384 and bench.workload(_args(config=4), 1)[3] == 64


def test_roofline_bookkeeping_is_the_surveys():
    import bench
    assert bench.n_springs(25, 25) == 3502
    assert bench.smem_bytes_per_substep(25, 25, bench.P_FLAT) == 411608 == bench.SMEM_B_PER_SUBSTEP
    assert bench.flop_per_substep(25, 25, bench.P_FLAT) == 170650 == bench.FLOP_PER_SUBSTEP
    assert bench.n_springs(64, 64) == 2 * 64 * 63 + 2 * 63 * 63 + 2 * 64 * 62


def test_draws_are_sharding_invariant_and_deterministic():
    import bench
    raw, pick = bench.draw_actions(1337, 4, 0, 3000)
    r1, p1 = bench.draw_actions(1337, 4, 0, 1100); r2, p2 = bench.draw_actions(1337, 4, 1100, 3000)
    assert np.array_equal(raw, np.concatenate([r1, r2])) and np.array_equal(pick, np.concatenate([p1, p2]))
    assert pick.min() >= 0 and pick.max() < 625 and raw.min() >= -1 and raw.max() <= 1
    assert not np.array_equal(raw, bench.draw_actions(1337, 5, 0, 3000)[0])
    assert np.array_equal(bench.actions_for_step(1337, 4, 0, 3000), raw)
    c = bench.restart_choice(1337, 3, 0, 50, 4096)
    assert np.array_equal(c, bench.restart_choice(1337, 3, 0, 50, 4096)) and c.max() < 4096
    assert np.array_equal(bench.restart_choice(7, 2, 0, 3000, 500), np.concatenate([bench.restart_choice(7, 2, 0, 1234, 500), bench.restart_choice(7, 2, 1234, 3000, 500)]))
    assert bench.draw_actions(1, 0, 0, 16, n_points=4096)[1].max() < 4096
