"""Image observations on the GPU.

`ClothRenderer` stands where `ClothEnv.get_blender_rep` stands in the reference (gym_cloth/envs/cloth_env.py:212-330):
that method writes the cloth as an .obj, starts a Blender process running gym_cloth/blender/get_image_rep_279.py,
sleeps a second, reads the PNG and post-processes it with cv2.  Here the same scene is rasterised for all
environments by one kernel launch (csrc/cloth_render.cuh) and post-processed by another; nothing leaves the device.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _l


def gamma_lut(gamma):
    """The table of ClothEnv._adjust_gamma (cloth_env.py:1217-1228)."""
    inv = 1.0 / gamma
    return np.array([((i / 255.0) ** inv) * 255 for i in np.arange(0, 256)]).astype("uint8")


class ClothRenderer(object):
    def __init__(self, params, n_env, device=None, scene=None):
        self.L = _l.lib()
        self.P = params
        self.n_env = int(n_env)
        self.device = torch.device(device if device is not None else "cuda")
        self.scene = scene if scene is not None else _l.default_scene()
        self.N = params.num_width_points * params.num_height_points
        self._env = _l.SceneEnv()
        self._keep = {}
        H, W = self.scene.height, self.scene.width
        self._zbuf = None
        self._minmax = None
        self._gray = None
        self.H, self.W = H, W

    # ---- per-environment scene values (domain randomisation, cloth_env.py:786-794) ----
    def set_env_values(self, **kw):
        """cam_pos_offset, cam_deg, front, back, bed: [n_env, 3] float; swap_sides: [n_env] int.  None clears."""
        for k, v in kw.items():
            if k not in ("cam_pos_offset", "cam_deg", "front", "back", "bed", "swap_sides"):
                raise KeyError(k)
            if v is None:
                self._keep.pop(k, None); setattr(self._env, k, None); continue
            dt = torch.int32 if k == "swap_sides" else torch.float32
            t = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).to(self.device, dt).contiguous()
            assert t.shape == ((self.n_env,) if k == "swap_sides" else (self.n_env, 3)), (k, tuple(t.shape))
            self._keep[k] = t
            setattr(self._env, k, t.data_ptr())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_pos(self, pos):
        assert pos.is_cuda and pos.is_contiguous() and tuple(pos.shape) == (self.n_env, self.N, 4), tuple(pos.shape)
        return "_f32" if pos.dtype == torch.float32 else "_f64"

    # ---- Blender's part: the PNG it would have written ----
    def rgb_raw(self, pos, out=None):
        sfx = self._check_pos(pos)
        if out is None:
            out = torch.empty(self.n_env, self.H, self.W, 3, dtype=torch.uint8, device=self.device)
        fn = getattr(self.L, "clothb200_render_rgb" + sfx)
        _l.check(fn(C.byref(self.P), C.byref(self.scene), C.byref(self._env), self.n_env, C.c_void_p(pos.data_ptr()),
                    C.c_void_p(out.data_ptr()), self._stream()), "render_rgb")
        return out

    def depth_raw(self, pos, out=None):
        sfx = self._check_pos(pos)
        if self._zbuf is None:
            self._zbuf = torch.empty(self.n_env, self.H, self.W, dtype=torch.float32, device=self.device)
            self._minmax = torch.empty(self.n_env, 2, dtype=torch.int32, device=self.device)
        if out is None:
            out = torch.empty(self.n_env, self.H, self.W, dtype=torch.uint8, device=self.device)
        fn = getattr(self.L, "clothb200_render_depth" + sfx)
        _l.check(fn(C.byref(self.P), C.byref(self.scene), C.byref(self._env), self.n_env, C.c_void_p(pos.data_ptr()),
                    C.c_void_p(self._zbuf.data_ptr()), C.c_void_p(self._minmax.data_ptr()), C.c_void_p(out.data_ptr()),
                    self._stream()), "render_depth")
        return out

    def camera_depth(self):
        """Camera-space depth [n_env, H, W] of the last depth render (1e10 where nothing was hit)."""
        return self._zbuf

    # ---- cloth_env.py's part: what it does to the loaded PNG (:296-315) ----
    def rgb(self, pos, lut=None, noise=None, out=None):
        """lut: [n_env, 256] uint8 (gamma tables) or None; noise: [n_env, H, W, 3] float32 or None."""
        img = self.rgb_raw(pos, out)
        if lut is not None or noise is not None:
            _l.check(self.L.clothb200_post_rgb(self.n_env, self.H, self.W, C.c_void_p(img.data_ptr()),
                                                C.c_void_p(lut.data_ptr()) if lut is not None else None,
                                                C.c_void_p(noise.data_ptr()) if noise is not None else None, self._stream()), "post_rgb")
        return img

    def depth(self, pos, sub=None, noise=None, out=None):
        """Three-channel depth observation: bilateral filter, minus `sub` ([n_env] float32, default 50), noise."""
        if self._gray is None:
            self._gray = torch.empty(self.n_env, self.H, self.W, dtype=torch.uint8, device=self.device)
        gray = self.depth_raw(pos, self._gray)
        if out is None:
            out = torch.empty(self.n_env, self.H, self.W, 3, dtype=torch.uint8, device=self.device)
        _l.check(self.L.clothb200_post_depth(self.n_env, self.H, self.W, C.c_void_p(gray.data_ptr()),
                                              C.c_void_p(sub.data_ptr()) if sub is not None else None,
                                              C.c_void_p(noise.data_ptr()) if noise is not None else None,
                                              C.c_void_p(out.data_ptr()), self._stream()), "post_depth")
        return out

    def rgbd(self, pos, lut=None, sub=None, noise_rgb=None, noise_depth=None):
        """np.dstack((img_rgb, img_d[:, :, 0])) of cloth_env.py:204-206, batched: [n_env, H, W, 4]."""
        c = self.rgb(pos, lut=lut, noise=noise_rgb)
        d = self.depth(pos, sub=sub, noise=noise_depth)
        return torch.cat([c, d[..., :1]], dim=-1)
