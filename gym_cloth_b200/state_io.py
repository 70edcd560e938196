"""The reference's cloth-state pickle: `ClothEnv.save_state` writes `{"pts": cloth.pts, "springs": cloth.springs}`
(cloth_env.py:343-350) and `ClothEnv(cfg, start_state_path=...)` loads it back and hands a deep copy to
`Cloth(state=...)` at every reset (cloth_env.py:120-124, 736-741; cloth.pyx:87-89).

The pickled objects are instances of `gym_cloth.physics.point.Point` (point.pyx:17-58) and
`gym_cloth.physics.cloth.Spring` (cloth.pyx:411-417) - plain Python classes compiled by Cython, so the stream holds
the class path plus each instance's attribute dict.  This module reads such a file into the arrays the B200 path keeps
(and checks that the spring list is the grid topology of cloth.pyx:134-146, the only one the kernels implement), and
writes one that the reference loads: with `gym_cloth` importable the reference's own classes are instantiated,
otherwise attribute-compatible stand-ins are pickled under the reference's class paths.
"""
import contextlib
import importlib
import pickle
import sys
import types

import numpy as np

_POINT_PATH = ("gym_cloth.physics.point", "Point")
_SPRING_PATH = ("gym_cloth.physics.cloth", "Spring")
_KINDS = ("STRUCTURAL", "STRUCTURAL", "SHEARING", "SHEARING", "BENDING", "BENDING")   # k = 0..5, cloth.pyx:135-146


class Point(object):
    """Attribute container with the names of point.pyx:34-58."""

    def __str__(self):
        return "({:.3f}, {:.3f}, {:.3f})".format(self.x, self.y, self.z)

    __repr__ = __str__


class Spring(object):
    """Attribute container with the names of cloth.pyx:413-417."""


Point.__module__, Point.__qualname__ = _POINT_PATH
Spring.__module__, Spring.__qualname__ = _SPRING_PATH


class _Unpickler(pickle.Unpickler):
    """Resolves the two reference classes to the stand-ins above (no gym_cloth install needed)."""

    def find_class(self, module, name):
        if (module, name) == _POINT_PATH:
            return Point
        if (module, name) == _SPRING_PATH:
            return Spring
        return super().find_class(module, name)


def _koffsets(W):
    return (W, 1, W + 1, W - 1, 2 * W, 2)        # q - ptA for the k-th spring created by point q


def state_to_arrays(state, W):
    """{"pts", "springs"} -> dict(pos, prev, pinned, orig [N,3] / [N], rest [6N] in slot order q*6+k, NaN where the
    grid has no spring).  Raises ValueError when the lists are not a W x W grid in the reference's creation order."""
    pts, springs = state["pts"], state["springs"]
    N = W * W
    if len(pts) != N:
        raise ValueError("state holds %d points, cfg says %d" % (len(pts), N))
    f = lambda names: np.array([[float(getattr(p, n)) for n in names] for p in pts], np.float64)
    out = {"pos": f(("x", "y", "z")), "prev": f(("px", "py", "pz")), "orig": f(("orig_x", "orig_y", "orig_z")),
           "pinned": np.array([bool(p.pinned) for p in pts], bool)}
    index = {id(p): i for i, p in enumerate(pts)}
    rest = np.full(6 * N, np.nan)
    off = _koffsets(W)
    it = iter(springs)
    n_seen = 0
    for q in range(N):
        r, c = divmod(q, W)
        valid = (r > 0, c > 0, r > 0 and c > 0, r > 0 and c + 1 < W, r > 1, c > 1)
        for k in range(6):
            if not valid[k]:
                continue
            sp = next(it, None)
            if sp is None:
                raise ValueError("state has fewer springs than a %dx%d grid" % (W, W))
            n_seen += 1
            a, b = index.get(id(sp.ptA)), index.get(id(sp.ptB))
            if a != q - off[k] or b != q or sp.type != _KINDS[k]:
                raise ValueError("spring %d is not the grid spring (%d, kind %d) of cloth.pyx:134-146" % (n_seen - 1, q, k))
            rest[q * 6 + k] = float(sp.rest_length)
    if next(it, None) is not None:
        raise ValueError("state has more springs than a %dx%d grid" % (W, W))
    out["rest"] = rest
    return out


def load_state(path):
    """Read a pickle written by the reference's ClothEnv.save_state (or by save_state below)."""
    with open(path, "rb") as fh:
        state = _Unpickler(fh).load()
    if not (isinstance(state, dict) and "pts" in state and "springs" in state):
        raise ValueError("%s is not a {'pts', 'springs'} cloth state" % path)
    return state


def _reference_classes():
    try:
        pm = importlib.import_module(_POINT_PATH[0]); cm = importlib.import_module(_SPRING_PATH[0])
        return getattr(pm, "Point"), getattr(cm, "Spring")
    except Exception:
        return None


@contextlib.contextmanager
def _stub_modules():
    """Make `gym_cloth.physics.{point,cloth}` resolve to the stand-ins while pickling (pickle stores classes by path and
    verifies the path)."""
    added = []
    for name, attr, cls in (("gym_cloth", None, None), ("gym_cloth.physics", None, None),
                            (_POINT_PATH[0], "Point", Point), (_SPRING_PATH[0], "Spring", Spring)):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name); added.append(name)
        if attr:
            setattr(sys.modules[name], attr, cls)
    try:
        yield
    finally:
        for name in added:
            sys.modules.pop(name, None)


def arrays_to_state(pos, prev, pinned, orig, rest, W, bounds=(1, 1, 1), use_reference_classes=True):
    """The reference's object graph for these arrays: W*W Points in row-major creation order and the springs of
    cloth.pyx:134-146, sharing the Point objects."""
    N = W * W
    ref = _reference_classes() if use_reference_classes else None
    pts = []
    for i in range(N):
        r, c = divmod(i, W)
        if ref:
            p = ref[0](float(pos[i, 0]), float(pos[i, 1]), float(pos[i, 2]), bounds[0], bounds[1], bounds[2], r, c)
        else:
            p = Point()
            p.boundsx, p.boundsy, p.boundsz = float(bounds[0]), float(bounds[1]), float(bounds[2])
            p.identity_0, p.identity_1 = float(r), float(c)
        p.x, p.y, p.z = float(pos[i, 0]), float(pos[i, 1]), float(pos[i, 2])
        p.px, p.py, p.pz = float(prev[i, 0]), float(prev[i, 1]), float(prev[i, 2])
        p.fx = p.fy = p.fz = 0.0
        p.pinned = bool(pinned[i])
        p.orig_x, p.orig_y, p.orig_z = float(orig[i, 0]), float(orig[i, 1]), float(orig[i, 2])
        pts.append(p)
    springs = []
    off = _koffsets(W)
    for q in range(N):
        r, c = divmod(q, W)
        valid = (r > 0, c > 0, r > 0 and c > 0, r > 0 and c + 1 < W, r > 1, c > 1)
        for k in range(6):
            if not valid[k]:
                continue
            if ref:
                sp = ref[1](pts[q - off[k]], pts[q], _KINDS[k])
            else:
                sp = Spring(); sp.ptA = pts[q - off[k]]; sp.ptB = pts[q]; sp.type = _KINDS[k]
            sp.rest_length = float(rest[q * 6 + k])
            springs.append(sp)
    return {"pts": pts, "springs": springs}, ref is not None


def save_state(path, pos, prev, pinned, orig, rest, W, bounds=(1, 1, 1)):
    """Write `{"pts", "springs"}` the way cloth_env.py:343-350 does."""
    state, real = arrays_to_state(pos, prev, pinned, orig, rest, W, bounds)
    # the object graph is deep (every spring references two points): the reference relies on the default recursion limit
    with open(path, "wb") as fh:
        if real:
            pickle.dump(state, fh)
        else:
            with _stub_modules():
                pickle.dump(state, fh)
