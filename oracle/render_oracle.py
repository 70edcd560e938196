"""CPU checker for the image observations (TEST INFRASTRUCTURE - see oracle/README.md).

PARITY UNPINNED for the rendered pixels: the reference produces its images by running Blender 2.79
(gym_cloth/blender/get_image_rep_279.py) and Blender is not available here, so there is nothing to pin the shading
against.  This file restates, in float64 numpy, the scene that script builds (camera :114-122/:267-277, bed
:143-156, floor :126-140/:455-462, lamp :467-469, two-sided cloth colours :236-253, mesh faces
cloth_env.py:226-231, Z pass normalisation :398-411) the way gym_cloth_b200/csrc/cloth_render.cuh draws it; tests
compare the CUDA images with it and pin the geometry separately.

The post-processing half IS pinned: `post_depth` / `post_rgb` call the very cv2 functions cloth_env.py:296-315 calls.
"""
import numpy as np

DEFAULT_SCENE = dict(
    height=224, width=224, samples=2, lens_mm=40.0, sensor_mm=36.0,
    cam_pos=(0.5, 0.5, 1.45), cam_deg=(0.0, 0.0, 0.0),
    lamp_pos=(4.07625, 1.00545, 5.90386), lamp_energy=1.5, diffuse_intensity=0.8, horizon=0.051,
    bed_z=-0.05, bed=(0.0, 1.0, 0.0, 1.0), floor_z=-0.25, floor=(-0.5, 1.5, -0.25, 1.25),
    front=(0.070, 0.050, 0.600), back=(0.070, 0.300, 0.900), bed_color=(1.0, 1.0, 1.0),
)


def _f32(x):
    return np.asarray(np.float32(x), np.float64)


def _camera(sc):
    C = _f32(sc["cam_pos"])
    d = np.deg2rad(_f32(sc["cam_deg"]))
    cx, sx, cy, sy, cz, sz = np.cos(d[0]), np.sin(d[0]), np.cos(d[1]), np.sin(d[1]), np.cos(d[2]), np.sin(d[2])
    R = np.array([[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx],
                  [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx],
                  [-sy, cy * sx, cy * cx]])            # Blender XYZ Euler, camera-to-world
    return C, R


def faces(W):
    f = []
    for r in range(W - 1):
        for c in range(W - 1):
            pp = r * W + c
            f.append([pp, pp + W, pp + 1])
            f.append([pp + 1, pp + W, pp + W + 1])
    return np.asarray(f)


def srgb(x):
    x = np.clip(x, 0.0, 1.0)
    return np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.power(x, 1 / 2.4) - 0.055)


def _raster(pts, W, sc, S):
    """Per sample: nearest cloth triangle id (-1 none), its barycentrics and depth."""
    H_, W_ = sc["height"], sc["width"]
    C, R = _camera(sc)
    fpx = float(np.float32(sc["lens_mm"]) / np.float32(sc["sensor_mm"])) * max(H_, W_)
    pc = (pts - C) @ R
    d = -pc[:, 2]
    sx = 0.5 * W_ + fpx * pc[:, 0] / d
    sy = 0.5 * H_ - fpx * pc[:, 1] / d
    w = 1.0 / d
    SH, SW = H_ * S, W_ * S
    zb = np.full((SH, SW), np.inf)
    tid = np.full((SH, SW), -1, np.int64)
    bar = np.zeros((SH, SW, 3))
    F = faces(W)
    for t, (a, b, c) in enumerate(F):
        if d[a] <= 0 or d[b] <= 0 or d[c] <= 0:
            continue
        area = (sx[b] - sx[a]) * (sy[c] - sy[a]) - (sy[b] - sy[a]) * (sx[c] - sx[a])
        if abs(area) < 1e-12:
            continue
        xs = (sx[a], sx[b], sx[c]); ys = (sy[a], sy[b], sy[c])
        j0 = max(0, int(np.floor(min(xs) * S - 0.5))); j1 = min(SW - 1, int(np.ceil(max(xs) * S - 0.5)))
        i0 = max(0, int(np.floor(min(ys) * S - 0.5))); i1 = min(SH - 1, int(np.ceil(max(ys) * S - 0.5)))
        if j1 < j0 or i1 < i0:
            continue
        X, Y = np.meshgrid((np.arange(j0, j1 + 1) + 0.5) / S, (np.arange(i0, i1 + 1) + 0.5) / S)
        l0 = ((sx[b] - X) * (sy[c] - Y) - (sy[b] - Y) * (sx[c] - X)) / area
        l1 = ((sx[c] - X) * (sy[a] - Y) - (sy[c] - Y) * (sx[a] - X)) / area
        l2 = 1.0 - l0 - l1
        dep = 1.0 / (l0 * w[a] + l1 * w[b] + l2 * w[c])
        sub_z = zb[i0:i1 + 1, j0:j1 + 1]
        m = (l0 >= 0) & (l1 >= 0) & (l2 >= 0) & (dep < sub_z) & (dep > 0.05) & (dep < 4.0)
        sub_z[m] = dep[m]
        tid[i0:i1 + 1, j0:j1 + 1][m] = t
        bar[i0:i1 + 1, j0:j1 + 1][m] = np.stack([l0, l1, l2], -1)[m]
    return dict(zb=zb, tid=tid, bar=bar, F=F, w=w, C=C, R=R, fpx=fpx)


def _planes(sc, rs, S, depth_mode):
    H_, W_ = sc["height"], sc["width"]
    SH, SW = H_ * S, W_ * S
    X, Y = np.meshgrid((np.arange(SW) + 0.5) / S, (np.arange(SH) + 0.5) / S)
    u = (X - 0.5 * W_) / rs["fpx"]; v = (0.5 * H_ - Y) / rs["fpx"]
    R, C = rs["R"], rs["C"]
    dirs = np.stack([R[0, 0] * u + R[0, 1] * v - R[0, 2], R[1, 0] * u + R[1, 1] * v - R[1, 2], R[2, 0] * u + R[2, 1] * v - R[2, 2]], -1)
    dpl = np.full((SH, SW), 1e10); which = np.zeros((SH, SW), np.int64)
    down = dirs[..., 2] < 0
    with np.errstate(divide="ignore", invalid="ignore"):
        tb = (float(np.float32(sc["bed_z"])) - C[2]) / dirs[..., 2]
        bx = C[0] + tb * dirs[..., 0]; by = C[1] + tb * dirs[..., 1]
        x0, x1, y0, y1 = [float(np.float32(q)) for q in sc["bed"]]
        mb = down & (tb > 0) & (bx >= x0) & (bx <= x1) & (by >= y0) & (by <= y1)
        dpl[mb] = tb[mb]; which[mb] = 1
        if depth_mode:
            tf = (float(np.float32(sc["floor_z"])) - C[2]) / dirs[..., 2]
            fx = C[0] + tf * dirs[..., 0]; fy = C[1] + tf * dirs[..., 1]
            x0, x1, y0, y1 = [float(np.float32(q)) for q in sc["floor"]]
            mf = down & ~mb & (tf > 0) & (fx >= x0) & (fx <= x1) & (fy >= y0) & (fy <= y1)
            dpl[mf] = tf[mf]; which[mf] = 2
    return dirs, dpl, which


def vertex_normals(pts, W):
    F = faces(W)
    fn = np.cross(pts[F[:, 1]] - pts[F[:, 0]], pts[F[:, 2]] - pts[F[:, 0]])
    vn = np.zeros_like(pts)
    for k in range(3):
        np.add.at(vn, F[:, k], fn)
    return vn / np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-15)


def render_rgb(pts, W, scene=None, swap_sides=False):
    """uint8 BGR [H, W, 3]: the PNG Blender would write for a colour observation, as cv2.imread returns it."""
    sc = dict(DEFAULT_SCENE); sc.update(scene or {})
    S = sc["samples"]
    pts = np.asarray(pts, np.float32).astype(np.float64)
    rs = _raster(pts, W, sc, S)
    dirs, dpl, which = _planes(sc, rs, S, False)
    C = rs["C"]
    kd = float(np.float32(sc["diffuse_intensity"])) * float(np.float32(sc["lamp_energy"]))
    lamp = _f32(sc["lamp_pos"])
    front, back = _f32(sc["front"]), _f32(sc["back"])
    if swap_sides:
        front, back = back, front
    col = np.full(dpl.shape + (3,), float(np.float32(sc["horizon"])))
    # bed
    mb = which == 1
    wp = C + dpl[..., None] * dirs
    L = lamp - np.concatenate([wp[..., :2], np.full(dpl.shape + (1,), float(np.float32(sc["bed_z"])))], -1)
    lam = np.maximum(0, L[..., 2] / np.linalg.norm(L, axis=-1)) * kd
    col[mb] = (_f32(sc["bed_color"])[None, :] * lam[mb][:, None])
    # cloth
    mc = (rs["tid"] >= 0) & (rs["zb"] <= dpl)
    if mc.any():
        F = rs["F"][rs["tid"][mc]]
        l = rs["bar"][mc]
        wv = rs["w"][F]
        dep = rs["zb"][mc]
        b = l * wv * dep[:, None]
        P = (b[..., None] * pts[F]).sum(1)
        vn = vertex_normals(pts, W)
        n = (b[..., None] * vn[F]).sum(1)
        g = np.cross(pts[F[:, 1]] - pts[F[:, 0]], pts[F[:, 2]] - pts[F[:, 0]])
        e = C - P
        is_front = (g * e).sum(1) >= 0
        flip = (n * e).sum(1) < 0
        n[flip] *= -1
        n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-15)
        Lc = lamp - P
        Lc /= np.linalg.norm(Lc, axis=1, keepdims=True)
        lamc = np.maximum(0, (n * Lc).sum(1)) * kd
        col[mc] = np.where(is_front[:, None], front[None, :], back[None, :]) * lamc[:, None]
    col = np.minimum(col, 1.0)
    H_, W_ = sc["height"], sc["width"]
    px = col.reshape(H_, S, W_, S, 3).mean(axis=(1, 3))
    img = np.floor(srgb(px) * 255.0 + 0.5).astype(np.uint8)
    return img[..., ::-1].copy()


def render_depth_raw(pts, W, scene=None, return_z=False):
    """uint8 [H, W]: the normalised Z pass Blender would write for a depth observation."""
    sc = dict(DEFAULT_SCENE); sc.update(scene or {})
    pts = np.asarray(pts, np.float32).astype(np.float64)
    rs = _raster(pts, W, sc, 1)
    dirs, dpl, which = _planes(sc, rs, 1, True)
    z = np.where((rs["tid"] >= 0) & (rs["zb"] <= dpl), rs["zb"], dpl)
    hit = z < 1e9
    v = np.ones_like(z)
    if hit.any():
        zmin, zmax = z[hit].min(), z[hit].max()
        v[hit] = (z[hit] - zmin) / (zmax - zmin) if zmax > zmin else 0.0
    img = np.floor(srgb(v) * 255.0 + 0.5).astype(np.uint8)
    return (img, z) if return_z else img


def post_depth(gray, gval=50.0, noise=None):
    """cloth_env.py:296-305, 312-315 on the three-channel PNG."""
    import cv2
    img = np.repeat(gray[:, :, None], 3, axis=2)
    img = cv2.bilateralFilter(img, 7, 50, 50)
    img = np.uint8(np.maximum(0, np.double(img) - gval))
    if noise is not None:
        img = np.uint8(np.minimum(np.maximum(np.double(img) + noise, 0), 255))
    return img


def post_rgb(bgr, gamma=None, noise=None):
    """cloth_env.py:306-315."""
    import cv2
    img = bgr
    if gamma is not None:
        inv = 1.0 / gamma
        table = np.array([((i / 255.0) ** inv) * 255 for i in np.arange(0, 256)]).astype("uint8")
        img = cv2.LUT(img, table)
    if noise is not None:
        img = np.uint8(np.minimum(np.maximum(np.double(img) + noise, 0), 255))
    return img
