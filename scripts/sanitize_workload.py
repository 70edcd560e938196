"""Workload for compute-sanitizer (scripts/sanitize.sh): a time-sliced launch that hammers the global-memory work queue
(more cloths than resident CTAs, slices of a few substeps), the bucket replay, the limit pass, the tear path and the
measurement tail, small enough to finish under racecheck.  Usage: python scripts/sanitize_workload.py [n_env] [slots] [slice]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_cloth_b200 import lib as L
from gym_cloth_b200.batched import BatchedCloth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 3
q = int(sys.argv[3]) if len(sys.argv) > 3 else 8
quick = len(sys.argv) > 4 and sys.argv[4] == "quick"          # racecheck: ~100x slower per shared-memory access
lib = L.lib()
rng = np.random.RandomState(3)
for dtype, mode in ((torch.float32, L.MODE_REFERENCE_ORDER), (torch.float64, L.MODE_REFERENCE_ORDER), (torch.float32, L.MODE_COLOURED)):
    P = L.default_params()
    P.iters_rest = 120; P.iters_grip_rest = 40                     # ~300-600 substeps per action instead of ~1800
    if quick:
        P.iters_rest = 30; P.iters_grip_rest = 10; P.iters_up = 20; P.iters_up_rest = 10
    bc = BatchedCloth(P, n, dtype=dtype, mode=mode)
    lib.clothb200_debug_set_slicing(slots, q)
    tot = 0
    for step in range(1 if quick else 2):                                           # the second action starts from crumpled cloths
        a = rng.uniform(-1, 1, size=(n, 4)); a[:, :2] *= 0.8; a[:, 2:] *= (0.3 if quick else 1.0)
        bc.step_actions(torch.from_numpy(a).to("cuda", dtype))
        torch.cuda.synchronize()
        tot += int(bc.sim_steps.sum().item())
    lib.clothb200_debug_set_slicing(0, 0)
    host = {"coverage": np.zeros(n), "flags": np.zeros(n, np.int32)}
    bc.step_host(rng.uniform(-0.7, 0.7, size=(n, 4)), host)       # unsliced host entry point
    print("dtype %s mode %d: %d substeps sliced over %d slots (slice %d), coverage %.4f, flags %s" % (
        str(dtype), mode, tot, slots, q, host["coverage"].mean(), sorted(set(host["flags"].tolist()))), flush=True)
print("sanitize workload done")
