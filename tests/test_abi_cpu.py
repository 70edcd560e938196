"""CPU-side checks of the C ABI: the shared library loads without a GPU, exports every symbol that
include/clothb200.h declares, and its host-side pieces (config constants, initial grid, exact action
decode) agree with fixtures produced by the reference.  No compute kernels are launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from gym_cloth_b200 import lib as L


@pytest.fixture(scope="module")
def so():
    from gym_cloth_b200.build import build
    build()
    return L.lib()


def _declared():
    src = open(os.path.join(ROOT, "include", "clothb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clothb200_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(so):
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(so, n), "missing export: " + n
    assert sorted(L.exported_symbols()) == names


def test_struct_layouts_and_defaults(so):
    assert so.clothb200_version() == 110
    P = L.default_params()
    assert (P.num_width_points, P.ks, P.iters_rest, P.grip_radius, P.max_actions) == (25, 10000.0, 1000.0, 0.003, 10)
    assert so.clothb200_error_string(-1).startswith(b"bad argument")
    Q = L.copy_params(P); Q.num_height_points = 24
    assert so.clothb200_params_validate(C.byref(Q)) == -2


def test_params_from_cfg_raises_like_reference(so):
    import yaml
    cfg = yaml.safe_load(open(os.path.join(ROOT, "gym_cloth_b200", "cfg", "t1_1d.yaml")))
    P = L.params_from_cfg(cfg)
    assert P.reduce_factor == 0.002 and P.gripper_height == 1
    bad = yaml.safe_load(open(os.path.join(ROOT, "gym_cloth_b200", "cfg", "t1_1d.yaml")))
    bad["cloth"]["pin_cond"] = "nonsense"
    with pytest.raises(ValueError):
        L.params_from_cfg(bad)
    bad = yaml.safe_load(open(os.path.join(ROOT, "gym_cloth_b200", "cfg", "t1_1d.yaml")))
    bad["init"]["type"] = "tier9"
    with pytest.raises(ValueError):
        L.params_from_cfg(bad)


def test_init_grid_matches_reference_grid(so):
    g = load_golden("kat_appendix_d.npz")   # pt 624 stays at (1,1,0): App. D
    P = L.default_params()
    pos = np.zeros((625, 4)); prev = np.zeros((625, 4)); rest = np.zeros(6 * 625)
    L.check(so.clothb200_init_grid_f64(C.byref(P), 1, None, 1, pos.ctypes.data, prev.ctypes.data, rest.ctypes.data))
    dx = 1.0 / 24
    r, c = np.divmod(np.arange(625), 25)
    assert np.array_equal(pos[:, 0], dx * r) and np.array_equal(pos[:, 1], dx * c) and not pos[:, 2:].any()
    assert np.array_equal(pos, prev)
    from gym_cloth_b200.batched import spring_slots
    slots = spring_slots(25)
    assert len(slots) == 3502 and np.count_nonzero(rest) == 3502
    t = load_golden("env_t1_s1337.npz")
    assert np.array_equal(rest[slots], t["rest"])            # Spring.rest_length of the reference, bit for bit
    # tier 2: vertical sheet with the reference's noise convention
    t2 = load_golden("env_t2_s1337.npz")
    import json
    log = json.loads(str(t2["rng_log"]))
    draws = [e[3] for e in log if e[0] == "rand"]
    side = draws[0] > 0.5
    noise = np.array(draws[1:626]) * 0.01 - 0.005
    L.check(so.clothb200_init_grid_f64(C.byref(P), 2, noise.ctypes.data, int(side), pos.ctypes.data, prev.ctypes.data, rest.ctypes.data))
    assert bool(t2["init_side"]) == side
    assert np.array_equal(rest[slots], t2["rest"])


def test_decode_host_is_bit_exact_with_env(so):
    t = load_golden("decode.npz")["table"]
    P = L.default_params()
    acts = np.ascontiguousarray(t[:, :4])
    plans = (L.Plan * len(acts))()
    L.check(so.clothb200_decode_actions_host(C.byref(P), len(acts), acts.ctypes.data, C.addressof(plans)))
    for i, row in enumerate(t):
        assert (plans[i].gx, plans[i].gy, plans[i].dxr, plans[i].dyr, plans[i].iters_pull) == (row[4], row[5], row[6], row[7], int(row[8]))


def test_argument_errors_do_not_need_a_gpu(so):
    P = L.default_params()
    assert so.clothb200_step_plans_f32(C.byref(P), 0, -1, None, None, 0, None) == -1
    assert so.clothb200_update_n_f32(C.byref(P), 0, 4, 1, None, None) == -1
    io = L.Step()
    assert so.clothb200_update_n_f32(C.byref(P), 0, 0, 1, C.byref(io), None) == 0      # empty batch is fine
    assert so.clothb200_update_n_f32(C.byref(P), 0, 4, 1, C.byref(io), None) == -1     # NULL tensors
    io.pos = 8; io.prev = 16
    assert so.clothb200_update_n_f32(C.byref(P), 0, 4, 1, C.byref(io), None) == -1     # misaligned for TMA
