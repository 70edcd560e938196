"""trimesh stand-in (only needed so the reference env module imports; the
Blender/obj observation path that uses it is out of scope and never called)."""


class Trimesh(object):
    def __init__(self, *a, **k):
        raise RuntimeError("trimesh stub: mesh export is out of scope")
