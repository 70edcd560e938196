"""Exploration: which cheap features of (state, action) predict a cloth's cycles per substep in the coming step?
Usage: python scripts/cost_model.py [n_env] [steps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gym_cloth_b200 import cfg_path
from gym_cloth_b200.envs import BatchedClothEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
env = BatchedClothEnv(cfg_path(1), n, dtype="f32", seed=5)
env.reset()
c = env.cloth
rows = []
for t in range(steps):
    a = torch.from_numpy(bench.actions_for_step(5, t, 0, n)).to(env.device, torch.float32)
    pos = c.pos[:, :, :3].clone()
    d = torch.cdist(pos, pos)                                     # [n, 625, 625]
    near = ((d < 0.04).sum((1, 2)) - 625).float() / 2            # pairs closer than the collision threshold
    near2 = ((d < 0.02).sum((1, 2)) - 625).float() / 2
    del d
    zmax = pos[:, :, 2].amax(1); zmean = pos[:, :, 2].mean(1)
    cov0 = c.coverage.clone().float(); cost0 = c.cost.clone()
    env.step(a)
    torch.cuda.synchronize()
    act = c.sim_steps > 0
    f = torch.stack([cost0, cov0, near, near2, zmax, zmean, c.sim_steps.float(), c.n_grabbed.float(), c.coverage.float(), c.cost], 1)
    rows.append(f[act].cpu().numpy())
    if t == 0:
        continue
X = np.concatenate(rows[1:])        # skip the first step (no previous cost)
names = ["cost_prev", "cov_before", "pairs<0.04", "pairs<0.02", "zmax", "zmean", "substeps", "n_grabbed", "cov_after", "cost_now"]
y = X[:, -1]
print("active env-steps:", len(y), " cost_now mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f" % (y.mean(), *np.percentile(y, [50, 90, 99, 100])))
for i, nm in enumerate(names[:-1]):
    print("  corr(cost_now, %-12s) = %+.3f" % (nm, np.corrcoef(X[:, i], y)[0, 1]))
# linear fits on a few feature sets, evaluated by how well they rank the heaviest 10 %
def fit(cols):
    A = np.column_stack([X[:, cols], np.ones(len(y))])
    half = len(y) // 2
    w, *_ = np.linalg.lstsq(A[:half], y[:half], rcond=None)
    p = A[half:] @ w; yy = y[half:]
    top = yy >= np.percentile(yy, 90)
    hit = (p >= np.percentile(p, 90))[top].mean()
    return np.corrcoef(p, yy)[0, 1], hit, w
for cols in ([0], [1], [2], [0, 1], [0, 2], [0, 1, 2], [0, 1, 2, 3, 4, 5]):
    r, hit, w = fit(cols)
    print("  fit on %-40s corr %.3f, recall of heaviest 10%% %.2f, w=%s" % ([names[c] for c in cols], r, hit, np.array2string(w, precision=1)))
