"""Import the compiled reference physics (TEST INFRASTRUCTURE - see oracle/README.md).

`load_physics()` returns the reference's own `Cloth`, `Gripper`, `Point` classes
from oracle/_ref (built by oracle/build_ref.py from /root/reference/gym_cloth/
physics/*.pyx).  Works on the GPU box too, because only the built extension
modules and our import stubs are needed.

`load_env()` additionally imports the reference's pure-Python `ClothEnv`
(gym_cloth/envs/cloth_env.py) straight from /root/reference; that only works in
the build container and is used by tests/golden/make_golden.py alone.
"""
import contextlib
import io
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")


def available():
    from oracle.build_ref import ref_built
    return ref_built()


def _ensure_paths():
    try:
        import gym  # noqa: F401  (a real gym, if one is ever installed, wins)
    except ImportError:
        if STUBS not in sys.path:
            sys.path.insert(0, STUBS)
    for mod in ("trimesh", "matplotlib"):
        try:
            __import__(mod)
        except ImportError:
            if STUBS not in sys.path:
                sys.path.insert(0, STUBS)


def _ensure_pkg(reference=None):
    """Make `gym_cloth` a package whose physics comes from oracle/_ref and whose
    envs (optionally) come from the reference checkout."""
    if "gym_cloth" in sys.modules and getattr(sys.modules["gym_cloth"], "_clothb200_ref", False):
        pkg = sys.modules["gym_cloth"]
    else:
        pkg = types.ModuleType("gym_cloth")
        pkg._clothb200_ref = True
        pkg.__path__ = [os.path.join(REF_OUT, "gym_cloth")]
        sys.modules["gym_cloth"] = pkg
    if reference is not None:
        p = os.path.join(reference, "gym_cloth")
        if p not in pkg.__path__:
            pkg.__path__.append(p)
    return pkg


def load_physics():
    _ensure_paths()
    _ensure_pkg()
    # point.pyx prints "Yes, cython compiled." at import (point.pyx:10-14)
    with contextlib.redirect_stdout(io.StringIO()):
        from gym_cloth.physics.cloth import Cloth
        from gym_cloth.physics.gripper import Gripper
        from gym_cloth.physics.point import Point
    return Cloth, Gripper, Point


def load_env(reference="/root/reference"):
    if not os.path.isdir(os.path.join(reference, "gym_cloth", "envs")):
        raise RuntimeError("reference checkout not present: %s" % reference)
    load_physics()
    _ensure_pkg(reference)
    # gym_cloth/envs/__init__.py lives in the reference checkout; physics stays ours
    envs = types.ModuleType("gym_cloth.envs")
    envs.__path__ = [os.path.join(reference, "gym_cloth", "envs")]
    sys.modules.setdefault("gym_cloth.envs", envs)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from gym_cloth.envs.cloth_env import ClothEnv
    sys.modules["gym_cloth.envs"].ClothEnv = ClothEnv   # what gym_cloth/envs/__init__.py exports
    return ClothEnv
