"""Image observations: the CUDA renderer against the numpy checker, geometry pins, and the
post-processing against the cv2 calls the reference makes (cloth_env.py:296-315)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _flat(W=25):
    r, c = np.meshgrid(np.arange(W), np.arange(W), indexing="ij")
    return np.stack([r / (W - 1.0), c / (W - 1.0), np.zeros_like(r, float)], -1).reshape(-1, 3)


def _folded():
    pts = _flat()
    m = (pts[:, 0] + pts[:, 1]) < 0.6
    out = pts.copy()
    out[m, 0] = 0.6 - pts[m, 1]; out[m, 1] = 0.6 - pts[m, 0]; out[m, 2] = 0.03
    return out


def _crumpled(n_actions=2, seed=3):
    import torch
    from gym_cloth_b200 import lib as L
    from gym_cloth_b200.batched import BatchedCloth
    bc = BatchedCloth(L.default_params(), 4, dtype=torch.float64)
    rng = np.random.RandomState(seed)
    for _ in range(n_actions):
        a = rng.uniform(-1, 1, size=(4, 4)); a[:, :2] *= 0.8
        bc.step_host(a, {})
    return bc.pos[:, :, :3].cpu().numpy()


def _render_all(states, dtype="f64"):
    import torch
    from gym_cloth_b200 import lib as L
    from gym_cloth_b200.render import ClothRenderer
    n = len(states)
    P = L.default_params()
    pos = torch.zeros(n, 625, 4, dtype=torch.float64 if dtype == "f64" else torch.float32, device="cuda")
    pos[:, :, :3] = torch.from_numpy(np.stack(states)).to(pos.dtype).cuda()
    r = ClothRenderer(P, n)
    return r, pos


def _close(a, b, tol, frac):
    bad = (np.abs(a.astype(int) - b.astype(int)) > tol)
    if bad.ndim == 3:
        bad = bad.any(-1)
    return bad.mean() <= frac, bad.mean()


def test_colour_image_matches_checker():
    from oracle import render_oracle as ro
    states = [_flat(), _folded()] + list(_crumpled())
    for dtype in ("f64", "f32"):
        r, pos = _render_all(states, dtype)
        img = r.rgb_raw(pos).cpu().numpy()
        assert img.shape == (len(states), 224, 224, 3) and img.dtype == np.uint8
        for i, st in enumerate(states):
            ref = ro.render_rgb(st, 25)
            # shading in f32 vs f64: 1 grey level; pixels where two cloth layers are closer than the depth key
            # resolves, or a sample sits exactly on an edge, may pick the other layer - bounded fraction
            ok, frac = _close(img[i], ref, 2, 0.01)
            assert ok, (dtype, i, frac)
            assert np.abs(img[i].astype(int) - ref.astype(int)).mean() < 0.5


def test_depth_image_matches_checker():
    from oracle import render_oracle as ro
    states = [_flat(), _folded()] + list(_crumpled())
    r, pos = _render_all(states, "f64")
    gray = r.depth_raw(pos).cpu().numpy()
    z = r.camera_depth().cpu().numpy()
    for i, st in enumerate(states):
        ref, zref = ro.render_depth_raw(st, 25, return_z=True)
        hit = (zref < 1e9) & (z[i] < 1e9)
        assert ((zref < 1e9) == (z[i] < 1e9)).mean() > 0.999
        assert (np.abs(z[i][hit] - zref[hit]) > 1e-4).mean() < 0.005
        ok, frac = _close(gray[i], ref, 2, 0.01)
        assert ok, (i, frac)


def test_geometry_pins():
    """Facts that follow from get_image_rep_279.py alone: camera 1.45 above (0.5, 0.5) looking down, 40 mm lens on a
    36 mm sensor, image x = world x, image up = world y; bed 0.05 below the cloth plane; front / back colours."""
    states = [_flat(), _folded()]
    r, pos = _render_all(states, "f64")
    img = r.rgb_raw(pos).cpu().numpy()
    f = 40.0 / 36.0 * 224
    half = f * 0.5 / 1.45                     # the flat cloth's half-width in pixels
    lo, hi = 112 - half, 112 + half
    flat = img[0]
    front = flat[112, 112]
    assert front[0] > 150 and front[1] < 100 and front[2] < 100           # BGR: the dark blue front side
    # cloth edge: inside is cloth, outside is the grey world (the bed is smaller in the image and hidden)
    for (row, col), inside in ((((112, int(lo) + 3)), True), ((112, int(lo) - 3), False), ((int(hi) - 3, 112), True), ((int(hi) + 3, 112), False)):
        px = flat[row, col]
        assert (np.abs(px.astype(int) - front.astype(int)).max() <= 10) == inside, (row, col, px)
    assert np.abs(flat[3, 3].astype(int) - 64).max() <= 2                  # horizon 0.051 through the display transform
    fold = img[1]
    # world (0.1, 0.1) is uncovered by the fold: white bed; image row grows downwards as world y shrinks
    col = int(112 + f * (0.1 - 0.5) / 1.5); row = int(112 - f * (0.1 - 0.5) / 1.5)
    assert fold[row, col].min() >= 250
    # world (0.4, 0.4) lies under the folded-over flap, which shows the lighter back side
    col = int(112 + f * (0.4 - 0.5) / 1.42); row = int(112 - f * (0.4 - 0.5) / 1.42)
    back = fold[row, col]
    assert back[0] > 230 and back[1] > 130 and back[2] < 100, back
    # exchanging the sides (tier2, init_side == -1) exchanges the colours
    r.set_env_values(swap_sides=np.array([1, 1]))
    sw = r.rgb_raw(pos).cpu().numpy()
    assert np.abs(sw[0][112, 112].astype(int) - back.astype(int)).max() <= 12
    # depth: nearest point black, floor white before the subtraction, bed in between
    r.set_env_values(swap_sides=None)
    g = r.depth_raw(pos).cpu().numpy()
    assert g[0][112, 112] == 0 and g[0][112, 3] == 255
    z = r.camera_depth().cpu().numpy()
    assert abs(z[0][112, 112] - 1.45) < 1e-5 and abs(z[0][112, 3] - 1.70) < 1e-5
    assert abs(z[1][int(112 - f * (0.1 - 0.5) / 1.5), int(112 + f * (0.1 - 0.5) / 1.5)] - 1.5) < 1e-5      # the bed


def test_camera_randomisation_moves_the_image():
    states = [_flat(), _flat()]
    r, pos = _render_all(states, "f64")
    r.set_env_values(cam_pos_offset=np.array([[0, 0, 0], [0.1, 0, 0]], np.float32), cam_deg=np.array([[0, 0, 0], [0, 0, 90.0]], np.float32))
    z = None
    img = r.rgb_raw(pos).cpu().numpy()
    f = 40.0 / 36.0 * 224
    cloth0 = (img[0][..., 0] > 150) & (img[0][..., 2] < 100)
    cloth1 = (img[1][..., 0] > 150) & (img[1][..., 2] < 100)
    c0 = np.array(np.nonzero(cloth0)).mean(1); c1 = np.array(np.nonzero(cloth1)).mean(1)
    assert np.abs(c0 - 111.5).max() < 0.6
    # camera moved +0.1 in world x and rolled 90 degrees about its axis: the cloth centre moves 0.1 * f / 1.45 px along
    # the image axis world x now maps to
    shift = 0.1 * f / 1.45
    assert abs(np.abs(c1 - 111.5).max() - shift) < 1.0 and np.abs(c1 - 111.5).min() < 0.6


def test_post_processing_is_the_reference_cv2_pipeline():
    import torch
    from oracle import render_oracle as ro
    from gym_cloth_b200.render import gamma_lut
    states = [_folded()] + list(_crumpled())
    r, pos = _render_all(states, "f64")
    n = len(states)
    rng = np.random.RandomState(0)
    noise = rng.uniform(-9, 9, size=(n, 224, 224, 3)).astype(np.float32)
    gv = rng.uniform(40, 50, size=n).astype(np.float32)
    gam = rng.uniform(0.7, 1.3, size=n)
    tn = torch.from_numpy(noise).cuda(); tg = torch.from_numpy(gv).cuda()
    lut = torch.from_numpy(np.stack([gamma_lut(g) for g in gam])).cuda()
    gray = r.depth_raw(pos).cpu().numpy()
    raw = r.rgb_raw(pos).cpu().numpy()
    d_plain = r.depth(pos).cpu().numpy()
    d_dr = r.depth(pos, sub=tg, noise=tn).cpu().numpy()
    c_dr = r.rgb(pos, lut=lut, noise=tn).cpu().numpy()
    for i in range(n):
        want = ro.post_depth(gray[i])                                       # bilateral 7/50/50, minus 50
        ok, frac = _close(d_plain[i], want, 1, 0.001)
        assert ok, frac
        assert (d_plain[i] != want).mean() < 0.02
        want = ro.post_depth(gray[i], gval=float(gv[i]), noise=noise[i].astype(np.float64))
        ok, frac = _close(d_dr[i], want, 1, 0.002)
        assert ok, frac
        want = ro.post_rgb(raw[i], gamma=gam[i], noise=noise[i].astype(np.float64))
        assert np.array_equal(c_dr[i], want)
    four = r.rgbd(pos).cpu().numpy()
    assert four.shape == (n, 224, 224, 4)
    assert np.array_equal(four[..., :3], raw) and np.array_equal(four[..., 3], d_plain[..., 0])


def test_env_with_image_observations():
    """obs_type 'blender' through the ClothEnv facade and the batched env (cloth_env.py:149-154, 201-209)."""
    import torch
    from gym_cloth_b200 import cfg_path
    from gym_cloth_b200.envs import ClothEnv, BatchedClothEnv
    env = ClothEnv(cfg_path(1, "rgbd"), dtype="f64")
    env.seed(1337)
    np.random.seed(5)
    obs = env.reset()
    assert obs.shape == (224, 224, 4) and obs.dtype == np.uint8
    assert env.observation_space.shape == (224, 224, 3)
    o2, rew, done, info = env.step((0.1, -0.2, 0.3, 0.3))
    assert o2.shape == (224, 224, 4) and (o2 != obs).any()
    # physics is untouched by the observation type: same seed, 1-D env, same coverage
    e1 = ClothEnv(cfg_path(1), dtype="f64"); e1.seed(1337); e1.reset()
    _, rew1, _, info1 = e1.step((0.1, -0.2, 0.3, 0.3))
    assert rew1 == rew and info1["actual_coverage"] == info["actual_coverage"]
    benv = BatchedClothEnv(cfg_path(3, "rgbd"), 8, dtype="f32", seed=1)
    ob = benv.reset()
    assert tuple(ob.shape) == (8, 224, 224, 4) and ob.dtype == torch.uint8
    a = torch.rand(8, 4, device="cuda") * 2 - 1
    ob2, rew, done, info = benv.step(a)
    assert tuple(ob2.shape) == (8, 224, 224, 4)
    ob3, rew, done, info = benv.step(np.random.uniform(-1, 1, size=(8, 4)))
    assert ob3.shape == (8, 224, 224, 4) and ob3.dtype == np.uint8


def test_large_grid_and_many_envs():
    import torch
    from gym_cloth_b200 import lib as L
    from gym_cloth_b200.render import ClothRenderer
    from oracle import render_oracle as ro
    P = L.default_params(); P.num_width_points = P.num_height_points = 64
    st = _flat(64); st[:, 2] = 0.05 * np.sin(6 * st[:, 0]) ** 2
    pos = torch.zeros(3, 4096, 4, dtype=torch.float32, device="cuda"); pos[:, :, :3] = torch.from_numpy(st).float().cuda()
    r = ClothRenderer(P, 3)
    img = r.rgb_raw(pos).cpu().numpy()
    ok, frac = _close(img[1], ro.render_rgb(st, 64), 2, 0.01)
    assert ok, frac
    n = 2048
    pos = torch.zeros(n, 625, 4, dtype=torch.float32, device="cuda"); pos[:, :, :3] = torch.from_numpy(_folded()).float().cuda()
    r = ClothRenderer(L.default_params(), n)
    img = r.rgbd(pos)
    assert torch.equal(img[0], img[n - 1])
