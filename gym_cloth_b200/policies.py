"""Analytic policies that drive the hot path, restated from examples/analytic.py.

`OracleCornerPolicy` / `RandomPolicy` keep the reference's interface (`set_env_cfg(env, cfg)`, `get_action(obs, t)`)
for the single-env facade; `oracle_corner_actions(benv)` is the batched form: one action per environment computed on
the device from the four near-corner points (indices 26, 48, 576, 598; analytic.py:109-125)."""
import numpy as np
import torch


class Policy(object):                                  # analytic.py:30-41
    def get_action(self, obs, t):
        raise NotImplementedError()

    def set_env_cfg(self, env, cfg):
        self.env = env
        self.cfg = cfg


def _corner_indices(tier2_flipped):
    # (ur, lr, ll, ul) in the order the reference evaluates them (analytic.py:109-125)
    if tier2_flipped:
        ll, ul, lr, ur = 576, 598, 26, 48
    else:
        ll, ul, lr, ur = 26, 48, 576, 598
    return ur, lr, ll, ul


_TARGETS = ((1, 1), (1, 0), (0, 0), (0, 1))            # targets of ur, lr, ll, ul


class OracleCornerPolicy(Policy):
    """Pull the cloth corner that is farthest from its target 90 % of the way there (analytic.py:70-155)."""

    def get_action(self, obs, t):
        if not self.cfg["env"]["delta_actions"]:
            raise NotImplementedError("the reference discourages the no-delta variant (analytic.py:157-160)")
        pts = self.env.cloth.pts
        assert len(pts) == 625, len(pts)
        flipped = self.cfg["init"]["type"] == "tier2" and (not self.env.cloth.init_side)
        best = None
        data = []
        for idx, (tx, ty) in zip(_corner_indices(flipped), _TARGETS):
            pt = pts[idx]
            x, y = pt.x, pt.y
            cx = (x - 0.5) * 2.0
            cy = (y - 0.5) * 2.0
            dx = (tx - x) * 0.90
            dy = (ty - y) * 0.90
            dist = np.sqrt((x - tx) ** 2 + (y - ty) ** 2)
            data.append((x, y, cx, cy, dx, dy, dist))
        maxdist = max(d[6] for d in data)
        for d in data:                                  # first corner that attains the maximum (analytic.py:140-148)
            if d[6] == maxdist:
                best = d
                break
        x, y, cx, cy, dx, dy, _ = best
        return (cx, cy, dx, dy) if self.cfg["env"]["clip_act_space"] else (x, y, dx, dy)


class HighestPointPolicy(Policy):
    """Pull one of the `top_k` highest points (chosen with the global np.random, as the reference does) 90 % of the
    way to where that point sits on the flat grid (analytic.py:716-808)."""

    def __init__(self):
        self.top_k = 5

    def _get_targ_xy(self, pt):                          # analytic.py:741-790
        if self.cfg["init"]["type"] in ("tier1", "tier3"):
            return pt.orig_x, pt.orig_y
        if self.env.cloth.init_side:
            return pt.orig_z, pt.orig_y
        return 1.0 - pt.orig_z, pt.orig_y

    def get_action(self, obs, t):
        assert self.cfg["env"]["delta_actions"]
        pts = self.env.cloth.pts
        order = sorted(range(len(pts)), key=lambda i: pts[i].z, reverse=True)      # stable, like sorted(pts, key=z)
        pt = pts[order[np.random.randint(self.top_k)]]
        targx, targy = self._get_targ_xy(pt)
        x, y = pt.x, pt.y
        cx = (x - 0.5) * 2.0
        cy = (y - 0.5) * 2.0
        dx = (targx - x) * 0.90
        dy = (targy - y) * 0.90
        return (cx, cy, dx, dy) if self.cfg["env"]["clip_act_space"] else (x, y, dx, dy)


class RandomPolicy(Policy):                             # analytic.py:811-822
    def get_action(self, obs, t):
        return self.env.get_random_action(atype="over_xy_plane")


def oracle_corner_actions(benv):
    """Batched OracleCornerPolicy: [n_env, 4] actions (clip space) as a device tensor of the env's dtype."""
    c = benv.cloth
    pos = c.pos
    n = benv.n_env
    flipped = torch.as_tensor((benv.init_type == "tier2") & (~benv.init_side), device=pos.device)
    idx_n = torch.tensor(_corner_indices(False), device=pos.device)
    idx_f = torch.tensor(_corner_indices(True), device=pos.device)
    idx = torch.where(flipped[:, None], idx_f[None, :], idx_n[None, :])            # [n, 4]
    xy = pos[torch.arange(n, device=pos.device)[:, None], idx, :2].double()        # [n, 4, 2]
    targ = torch.tensor(_TARGETS, dtype=torch.float64, device=pos.device)[None]    # [1, 4, 2]
    dist = ((xy - targ) ** 2).sum(-1).sqrt()
    k = torch.argmax((dist == dist.max(dim=1, keepdim=True).values).to(torch.int8), dim=1)   # first maximal corner
    sel = xy[torch.arange(n, device=pos.device), k]
    tsel = targ[0][k]
    act = torch.cat([(sel - 0.5) * 2.0, (tsel - sel) * 0.90], dim=1)
    return act.to(c.dtype)


def highest_point_actions(benv, top_k=5, generator=None):
    """Batched HighestPointPolicy: [n_env, 4] actions (clip space).  One of each environment's `top_k` highest points
    (ties in index order, like Python's stable sort) is chosen with `generator` (a torch.Generator on the env's device)."""
    pos = benv.cloth.pos
    n = benv.n_env
    z = pos[:, :, 2].double()
    order = torch.sort(z, dim=1, descending=True, stable=True).indices[:, :top_k]              # [n, k]
    pick = torch.randint(0, top_k, (n,), device=pos.device, generator=generator)
    ar = torch.arange(n, device=pos.device)
    idx = order[ar, pick]
    xy = pos[ar, idx, :2].double()
    orig = benv.orig_pos[ar, idx]                                                               # [n, 3]
    if benv.init_type == "tier2":
        side = torch.as_tensor(benv.init_side, device=pos.device)
        tx = torch.where(side, orig[:, 2], 1.0 - orig[:, 2])
    else:
        tx = orig[:, 0]
    targ = torch.stack([tx, orig[:, 1]], dim=1)
    act = torch.cat([(xy - 0.5) * 2.0, (targ - xy) * 0.90], dim=1)
    return act.to(benv.torch_dtype)
