"""The numpy checker of the image observations (oracle/render_oracle.py) against facts that follow from
gym_cloth/blender/get_image_rep_279.py alone; the CUDA renderer is compared with this checker in test_gpu_render.py."""
import numpy as np

from oracle import render_oracle as ro


def _flat(W=25):
    r, c = np.meshgrid(np.arange(W), np.arange(W), indexing="ij")
    return np.stack([r / (W - 1.0), c / (W - 1.0), np.zeros_like(r, float)], -1).reshape(-1, 3)


def test_faces_are_the_reference_mesh():
    # cloth_env.py:226-231
    F = ro.faces(25)
    assert F.shape == (2 * 24 * 24, 3)
    assert F[0].tolist() == [0, 25, 1] and F[1].tolist() == [1, 25, 26]
    assert F[-1].tolist() == [24 * 25 - 1, 624 - 1, 624]


def test_flat_cloth_projection_and_colours():
    img = ro.render_rgb(_flat(), 25)
    assert img.shape == (224, 224, 3) and img.dtype == np.uint8
    f = 40.0 / 36.0 * 224
    half = f * 0.5 / 1.45
    cloth = (img[..., 0] > 150) & (img[..., 2] < 100)            # BGR, dark blue front
    rows, cols = np.nonzero(cloth)
    assert abs(rows.min() - (112 - half)) < 1.5 and abs(rows.max() - (111 + half)) < 1.5
    assert abs(cols.min() - (112 - half)) < 1.5 and abs(cols.max() - (111 + half)) < 1.5
    assert np.abs(img[2, 2].astype(int) - 64).max() <= 1          # world horizon 0.051 -> sRGB
    # lifting a corner reveals the white bed, whose image is smaller (it is 0.05 further away)
    pts = _flat(); pts[:, 2] = 0.0
    pts[(pts[:, 0] < 0.2) & (pts[:, 1] < 0.2)] += np.array([0.3, 0.3, 0.05])
    img = ro.render_rgb(pts, 25)
    col = int(112 + f * (0.1 - 0.5) / 1.5); row = int(112 - f * (0.1 - 0.5) / 1.5)
    assert img[row, col].min() >= 250


def test_depth_normalisation_and_post_processing():
    g, z = ro.render_depth_raw(_flat(), 25, return_z=True)
    assert abs(z[112, 112] - 1.45) < 1e-6 and abs(z[112, 2] - 1.70) < 1e-6
    assert g[112, 112] == 0 and g[112, 2] == 255
    # top/bottom rows look past the floor (its y extent is 1.5, the view's 1.53 at that distance): background = 1.0
    assert z[0, 112] > 1e9 and g[0, 112] == 255
    d = ro.post_depth(g)
    assert d.shape == (224, 224, 3) and d[112, 2, 0] == 205 and d[112, 112, 0] == 0
    c = ro.post_rgb(ro.render_rgb(_flat(), 25), gamma=1.0)
    assert np.array_equal(c, ro.render_rgb(_flat(), 25))
