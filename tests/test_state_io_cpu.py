"""The reference's cloth-state pickle (cloth_env.py:343-350 / 120-124): gym_cloth_b200/state_io.py reads the file the
reference itself wrote (tests/golden/state_t1_s1337.pkl, produced by tests/golden/make_golden.py state) without any
gym_cloth install, and writes files the reference's classes load."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from gym_cloth_b200 import state_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def _fixture():
    return np.load(os.path.join(G, "state_t1_s1337.npz")), os.path.join(G, "state_t1_s1337.pkl")


def test_reads_reference_pickle_bit_exact():
    d, pkl = _fixture()
    a = state_io.state_to_arrays(state_io.load_state(pkl), 25)
    assert np.array_equal(a["pos"], d["pos_saved"]) and np.array_equal(a["prev"], d["prev_saved"])
    assert np.array_equal(a["pinned"], d["pin_saved"].astype(bool))
    assert np.array_equal(a["orig"], d["orig_saved"])
    from gym_cloth_b200.batched import spring_slots
    assert np.array_equal(a["rest"][spring_slots(25)], d["rest_saved"])       # Spring.rest_length in creation order
    assert np.isnan(a["rest"]).sum() == 6 * 625 - 3502


def test_round_trip_through_stand_in_classes(tmp_path):
    d, pkl = _fixture()
    a = state_io.state_to_arrays(state_io.load_state(pkl), 25)
    out = str(tmp_path / "again.pkl")
    state_io.save_state(out, a["pos"], a["prev"], a["pinned"], a["orig"], a["rest"], 25)
    b = state_io.state_to_arrays(state_io.load_state(out), 25)
    for k in ("pos", "prev", "pinned", "orig"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(np.nan_to_num(a["rest"]), np.nan_to_num(b["rest"]))
    # the stream names the reference's classes, not ours
    raw = open(out, "rb").read()
    assert b"gym_cloth.physics.point" in raw and b"gym_cloth.physics.cloth" in raw and b"state_io" not in raw
    assert "gym_cloth.physics.point" not in sys.modules or hasattr(sys.modules["gym_cloth.physics.point"], "Point")


def test_rejects_foreign_topology():
    _, pkl = _fixture()
    st = state_io.load_state(pkl)
    st["springs"] = st["springs"][:-1]
    with pytest.raises(ValueError):
        state_io.state_to_arrays(st, 25)
    st = state_io.load_state(pkl)
    st["springs"][10], st["springs"][11] = st["springs"][11], st["springs"][10]
    with pytest.raises(ValueError):
        state_io.state_to_arrays(st, 25)


def test_reference_classes_load_our_file(tmp_path):
    """A file written here is unpickled by the reference's own compiled Point / Spring classes (oracle/_ref), in a
    fresh interpreter that has never imported this package."""
    from oracle.build_ref import ref_built
    if not ref_built():
        pytest.skip("oracle/_ref not built")
    d, pkl = _fixture()
    a = state_io.state_to_arrays(state_io.load_state(pkl), 25)
    out = str(tmp_path / "ours.pkl")
    state_io.save_state(out, a["pos"], a["prev"], a["pinned"], a["orig"], a["rest"], 25)
    code = (
        "import sys, pickle, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from oracle.ref_loader import load_physics\n"
        "Cloth, Gripper, Point = load_physics()\n"
        "st = pickle.load(open(%r, 'rb'))\n"
        "assert type(st['pts'][0]) is Point, type(st['pts'][0])\n"
        "assert type(st['springs'][0]).__module__ == 'gym_cloth.physics.cloth'\n"
        "np.save(%r, np.array([[p.x, p.y, p.z, p.px, p.py, p.pz, p.orig_x, p.orig_y, p.orig_z, float(p.pinned)] for p in st['pts']]))\n"
        "np.save(%r, np.array([s.rest_length for s in st['springs']]))\n"
        "assert st['springs'][0].ptB is st['pts'][1] and st['springs'][0].ptA is st['pts'][0]\n"
    ) % (ROOT, out, str(tmp_path / "pts.npy"), str(tmp_path / "rest.npy"))
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    pts = np.load(str(tmp_path / "pts.npy")); rest = np.load(str(tmp_path / "rest.npy"))
    assert np.array_equal(pts[:, 0:3], d["pos_saved"]) and np.array_equal(pts[:, 3:6], d["prev_saved"])
    assert np.array_equal(pts[:, 6:9], d["orig_saved"]) and np.array_equal(rest, d["rest_saved"])
