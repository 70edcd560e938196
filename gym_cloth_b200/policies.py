"""Analytic policies that drive the hot path, restated from examples/analytic.py.

`OracleCornerPolicy`, `OracleCornerRevealPolicy`, `HighestPointPolicy`, `WrinklesPolicy`, `RandomPolicy` keep the reference's interface (`set_env_cfg(env, cfg)`, `get_action(obs, t)`)
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


class OracleCornerRevealPolicy(Policy):
    """analytic.py:217-358 ('alg1', delta actions): the oracle corner policy restricted to the corners the camera can see
    (`env._occlusion_vec`, order ur, lr, ll, ul; True = occluded) - while a visible corner is further than 0.09 from its
    target it is pulled there; otherwise the occluded corner furthest from its target is pulled AWAY from the target by
    half a bed width, to reveal it.  The vector is whatever the environment holds: the reference fills it from Blender's
    ray casts when it renders (cloth_env.py:287-292) and leaves it all-True otherwise (:115), as this package does."""

    def __init__(self):
        self._sign = 1

    def get_action(self, obs, t):
        if not self.cfg["env"]["delta_actions"]:
            raise NotImplementedError("the reference discourages the no-delta variant")
        pts = self.env.cloth.pts
        assert len(pts) == 625, len(pts)
        flipped = self.cfg["init"]["type"] == "tier2" and (not self.env.cloth.init_side)
        data = []
        for idx, (tx, ty) in zip(_corner_indices(flipped), _TARGETS):
            x, y = pts[idx].x, pts[idx].y
            data.append((x, y, (x - 0.5) * 2.0, (y - 0.5) * 2.0, (tx - x) * 0.90, (ty - y) * 0.90, np.sqrt((x - tx) ** 2 + (y - ty) ** 2)))
        occ = list(self.env._occlusion_vec)
        distances = [d[6] for d in data]
        vis = [d if not occ[i] else 0 for i, d in enumerate(distances)]
        hid = [d if occ[i] else 0 for i, d in enumerate(distances)]
        maxvis, maxhid = max(vis), max(hid)
        thresh_dist = 0.09
        if (False in occ and maxvis > thresh_dist) or (True not in occ and maxvis < thresh_dist):
            maxdist, distances, self._sign = maxvis, vis, 1
        else:
            maxdist, distances, self._sign = maxhid, hid, -1
        x, y, cx, cy, dx, dy, _ = next(d for d, dist in zip(data, distances) if dist == maxdist)   # first corner at the maximum
        if self._sign == -1:
            scaling_factor = 0.5 / maxdist
            dx *= scaling_factor
            dy *= scaling_factor
        if self.cfg["env"]["clip_act_space"]:
            return (cx, cy, dx * self._sign, dy * self._sign)
        return (x, y, dx * self._sign, dy * self._sign)


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


def oracle_corner_reveal_actions(benv, occlusion):
    """Batched OracleCornerRevealPolicy: `occlusion` [n_env, 4] bool (ur, lr, ll, ul; True = occluded) -> [n_env, 4] actions."""
    c = benv.cloth
    pos = c.pos
    n = benv.n_env
    dev = pos.device
    occ = torch.as_tensor(occlusion, device=dev).bool().reshape(n, 4)
    flipped = torch.as_tensor((benv.init_type == "tier2") & (~benv.init_side), device=dev)
    idx = torch.where(flipped[:, None], torch.tensor(_corner_indices(True), device=dev)[None, :],
                      torch.tensor(_corner_indices(False), device=dev)[None, :])
    ar = torch.arange(n, device=dev)
    xy = pos[ar[:, None], idx, :2].double()
    targ = torch.tensor(_TARGETS, dtype=torch.float64, device=dev)[None]
    dist = ((xy - targ) ** 2).sum(-1).sqrt()
    vis = torch.where(occ, torch.zeros_like(dist), dist); hid = torch.where(occ, dist, torch.zeros_like(dist))
    maxvis = vis.max(dim=1).values
    pull = ((~occ).any(dim=1) & (maxvis > 0.09)) | ((~occ).all(dim=1) & (maxvis < 0.09))
    d = torch.where(pull[:, None], vis, hid)
    maxd = d.max(dim=1, keepdim=True).values
    k = torch.argmax((d == maxd).to(torch.int8), dim=1)
    sel = xy[ar, k]
    delta = (targ[0][k] - sel) * 0.90
    scale = torch.where(pull, torch.ones_like(maxd[:, 0]), -0.5 / maxd[:, 0])
    return torch.cat([(sel - 0.5) * 2.0, delta * scale[:, None]], dim=1).to(c.dtype)


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


# ---------------------------------------------------------------------------------------------- wrinkle policy
def _wrinkle_levels():
    """The z levels is_point_on_top scans (analytic.py:571-586): 1, 1-0.02, ... by repeated subtraction while > 0."""
    out = []
    z = 1
    while z > 0:
        out.append(z)
        z -= 0.02
    return np.array(out, np.float64)


def _neighbors(r, c, h, w):
    """get_neighbors (analytic.py:588-612), in the reference's order (its `c < h - 1` included)."""
    pts = [r * w + c]
    if r > 0 and c > 0:
        pts.append((r - 1) * w + c - 1)
    if r > 0:
        pts.append((r - 1) * w + c)
    if c > 0:
        pts.append(r * w + c - 1)
    if r < h - 1 and c < w - 1:
        pts.append((r + 1) * w + c + 1)
    if r < h - 1:
        pts.append((r + 1) * w + c)
    if c < h - 1:
        pts.append(r * w + c + 1)
    if r > 0 and c < w - 1:
        pts.append((r - 1) * w + c + 1)
    if r < h - 1 and c > 0:
        pts.append((r + 1) * w + c - 1)
    return pts


def _wrinkle_pull(center, wrinkle_pt, xs, ys):
    """analytic.py:676-720: pull perpendicular to the wrinkle, from the cloth point nearest to where that line (or the
    diagonal it is closest to) leaves the unit square.  center / wrinkle_pt: (x, y); xs, ys: all points."""
    if wrinkle_pt[0] == center[0]:
        slope = 1000
    else:
        slope = (wrinkle_pt[1] - center[1]) / (wrinkle_pt[0] - center[0])
    perp_slope = -1 / slope
    r2 = np.sqrt(2)
    if perp_slope > r2 + 1 or perp_slope < -(r2 + 1):
        x1 = (1 - center[1]) / perp_slope + center[0]; y1 = 1
        x2 = (-center[1]) / perp_slope + center[0]; y2 = 0
    elif perp_slope > 1 and perp_slope < r2 + 1:
        x1, y1, x2, y2 = 1, 1, 0, 0
    elif perp_slope < r2 - 1 and perp_slope > -(r2 - 1):
        y1 = perp_slope * (1 - center[0]) + center[1]; x1 = 1
        y2 = perp_slope * (-center[0]) + center[1]; x2 = 0
    else:
        x1, y1, x2, y2 = 0, 1, 1, 0
    i1 = int(np.argmin(np.sqrt((x1 - xs) ** 2 + (y1 - ys) ** 2)))
    i2 = int(np.argmin(np.sqrt((x2 - xs) ** 2 + (y2 - ys) ** 2)))
    d1 = np.sqrt((center[0] - xs[i1]) ** 2 + (center[1] - ys[i1]) ** 2)
    d2 = np.sqrt((center[0] - xs[i2]) ** 2 + (center[1] - ys[i2]) ** 2)
    if d1 < d2 or xs[i2] < 0.01 or xs[i2] > 0.99 or ys[i2] < 0.01 or ys[i2] > 0.99:
        x, y = xs[i1], ys[i1]; dx, dy = x1 - x, y1 - y
    else:
        x, y = xs[i2], ys[i2]; dx, dy = x2 - x, y2 - y
    return (2 * (x - 0.5), 2 * (y - 0.5), dx, dy)


class WrinklesPolicy(Policy):
    """examples/analytic.py:551-720 on the ground-truth state: the point whose 3x3 neighbourhood deviates most in height
    (among the points that are on top of the cloth where they are) is the wrinkle's centre, its most deviating neighbour
    gives the wrinkle's direction, and the cloth is pulled perpendicular to it towards the edge of the plane."""

    def get_action(self, obs, t):
        cloth = self.env.cloth
        pts = cloth.pts
        w = h = int(round(len(pts) ** 0.5))
        xs = np.array([p.x for p in pts]); ys = np.array([p.y for p in pts]); zs = np.array([p.z for p in pts])
        levels = _wrinkle_levels()
        deviation = []
        for i in range(len(pts)):
            near = ((xs - xs[i]) * (xs - xs[i]) + (ys - ys[i]) * (ys - ys[i])) < 0.0002      # includes i itself
            band = np.abs(zs[near][:, None] - levels[None, :]) < 2 * 0.02                       # [members, levels]
            hit = band.any(axis=0)
            on_top = bool(hit.any()) and bool(abs(zs[i] - levels[int(np.argmax(hit))]) < 2 * 0.02)
            if on_top:
                r, c = divmod(i, w)
                points = np.array([zs[x] for x in _neighbors(r, c, h, w)])
                deviation.append(np.sum(np.abs(points - np.mean(points))))
            else:
                deviation.append(0)
        p = deviation.index(max(deviation))
        c = p % w
        r = (p - c) // w
        indices = _neighbors(r, c, h, w)
        indices.remove(r * w + c)
        mx = max(indices, key=lambda index: deviation[index])
        return _wrinkle_pull((xs[p], ys[p]), (xs[mx], ys[mx]), xs, ys)


_NB_CACHE = {}


def _neighbor_table(w, device):
    key = (w, str(device))
    if key not in _NB_CACHE:
        tab = np.full((w * w, 9), -1, np.int64)
        for i in range(w * w):
            nb = _neighbors(i // w, i % w, w, w)
            tab[i, :len(nb)] = nb
        _NB_CACHE[key] = torch.from_numpy(tab).to(device)
    return _NB_CACHE[key]


def wrinkle_actions(benv, chunk=128):
    """Batched WrinklesPolicy: [n_env, 4] actions (clip space), float64, computed on the device from the state tensors
    (pairwise tests in chunks of `chunk` environments).  Same decisions as the single-environment class up to ties."""
    pos = benv.cloth.pos
    dev = pos.device
    n, N = benv.n_env, benv.N
    w = int(round(N ** 0.5))
    levels = torch.from_numpy(_wrinkle_levels()).to(dev)
    tab = _neighbor_table(w, dev)
    valid = tab >= 0
    tabc = tab.clamp(min=0)
    out = torch.empty(n, 4, dtype=torch.float64, device=dev)
    r2 = 2.0 ** 0.5
    for s in range(0, n, chunk):
        P = pos[s:s + chunk, :, :3].double()
        m = P.shape[0]
        ar = torch.arange(m, device=dev)
        x, y, z = P[:, :, 0], P[:, :, 1], P[:, :, 2]
        dxm = x[:, None, :] - x[:, :, None]; dym = y[:, None, :] - y[:, :, None]
        near = (dxm * dxm + dym * dym) < 0.0002                                            # [m, N, N]
        band = (z[:, :, None] - levels[None, None, :]).abs() < 2 * 0.02                     # [m, N, L]
        hit = torch.bmm(near.to(torch.float32), band.to(torch.float32)) > 0                 # [m, N, L]: some near point in the band
        first = torch.argmax(hit.to(torch.int8), dim=2)
        on_top = hit.any(dim=2) & torch.gather(band, 2, first[:, :, None])[:, :, 0]
        zn = z[:, tabc]                                                                     # [m, N, 9]
        cnt = valid.sum(1).to(torch.float64)[None, :]
        mean = (zn * valid[None]).sum(2) / cnt
        dev_map = ((zn - mean[:, :, None]).abs() * valid[None]).sum(2)
        dev_map = torch.where(on_top, dev_map, torch.zeros_like(dev_map))
        p = torch.argmax((dev_map == dev_map.max(dim=1, keepdim=True).values).to(torch.int8), dim=1)   # first maximum
        nb = tab[p][:, 1:]; nbv = nb >= 0                                                   # neighbours of the centre, self removed
        nd = torch.where(nbv, dev_map[ar[:, None], nb.clamp(min=0)], torch.full_like(nb, -1.0, dtype=torch.float64))
        k = torch.argmax((nd == nd.max(dim=1, keepdim=True).values).to(torch.int8), dim=1)
        q = nb[ar, k]
        cx, cy = x[ar, p], y[ar, p]
        wx, wy = x[ar, q], y[ar, q]
        slope = torch.where(wx == cx, torch.full_like(cx, 1000.0), (wy - cy) / (wx - cx))
        ps = -1.0 / slope
        ns = (ps > r2 + 1) | (ps < -(r2 + 1))
        ne = (ps > 1) & (ps < r2 + 1) & ~ns
        ew = (ps < r2 - 1) & (ps > -(r2 - 1)) & ~ns & ~ne
        one, zero = torch.ones_like(cx), torch.zeros_like(cx)
        x1 = torch.where(ns, (1 - cy) / ps + cx, torch.where(ne, one, torch.where(ew, one, zero)))
        y1 = torch.where(ns, one, torch.where(ne, one, torch.where(ew, ps * (1 - cx) + cy, one)))
        x2 = torch.where(ns, (-cy) / ps + cx, torch.where(ne, zero, torch.where(ew, zero, one)))
        y2 = torch.where(ns, zero, torch.where(ne, zero, torch.where(ew, ps * (-cx) + cy, zero)))
        i1 = torch.argmin(((x1[:, None] - x) ** 2 + (y1[:, None] - y) ** 2).sqrt(), dim=1)
        i2 = torch.argmin(((x2[:, None] - x) ** 2 + (y2[:, None] - y) ** 2).sqrt(), dim=1)
        ax1, ay1, ax2, ay2 = x[ar, i1], y[ar, i1], x[ar, i2], y[ar, i2]
        d1 = ((cx - ax1) ** 2 + (cy - ay1) ** 2).sqrt(); d2 = ((cx - ax2) ** 2 + (cy - ay2) ** 2).sqrt()
        use1 = (d1 < d2) | (ax2 < 0.01) | (ax2 > 0.99) | (ay2 < 0.01) | (ay2 > 0.99)
        gx = torch.where(use1, ax1, ax2); gy = torch.where(use1, ay1, ay2)
        tx = torch.where(use1, x1, x2); ty = torch.where(use1, y1, y2)
        out[s:s + chunk] = torch.stack([2 * (gx - 0.5), 2 * (gy - 0.5), tx - gx, ty - gy], dim=1)
    return out
