"""One-off / on-demand fuzz of the bit-exact build: n crumpled tier-1 states (f64 reset on the device), `steps` aimed random
actions each (bench.py's draws), every resulting state compared bit for bit with the CPU oracle run from the same state
on all host cores.  Usage: python scripts/fuzz_f64.py [n_env] [steps] [seed] [tier]   (tier 2: per-cloth rest lengths)"""
import multiprocessing as mp
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gym_cloth_b200 import cfg_path
from gym_cloth_b200.envs import BatchedClothEnv


def _oracle(job):
    from oracle.oracle import OracleCloth
    pos, prev, act, rest = job
    o = OracleCloth()
    o.set_state(pos, prev, np.zeros(len(pos), np.uint8))
    if rest is not None:
        o.set_rest(rest)
    n, ng, ip = o.step_action(act)
    st = o.get_state()
    return n, st[0], st[1], o.coverage(), bool(o.tear)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 4242
    tier = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    from gym_cloth_b200.batched import spring_slots
    env = BatchedClothEnv(cfg_path(tier), n, dtype="f64", seed=seed)
    env.reset()
    c = env.cloth
    rest = None
    if c.rest is not None and c.rest_env_stride:                 # tier 2: Spring.rest_length per cloth, reference order
        rest = c.rest.cpu().numpy()[:, spring_slots(c.W)]
    bad = 0; total = 0; sub = 0; maxb = 0
    for t in range(steps):
        pos0 = c.pos[:, :, :3].cpu().numpy().copy(); prev0 = c.prev[:, :, :3].cpu().numpy().copy()
        raw, pick = bench.draw_actions(seed, t, 0, n)
        a = raw.copy()
        a[:, :2] = (pos0[np.arange(n), pick, :2] - 0.5) * 2
        a = np.clip(a, -1, 1)
        host = {"coverage": np.zeros(n), "sim_steps": np.zeros(n, np.int32), "flags": np.zeros(n, np.int32)}
        c.step_host(a, host)                                  # host decode: the reference's exact arithmetic
        with mp.get_context("fork").Pool(os.cpu_count()) as pool:
            res = pool.map(_oracle, [(pos0[e], prev0[e], a[e], None if rest is None else rest[e]) for e in range(n)], chunksize=4)
        gp = c.pos[:, :, :3].cpu().numpy(); gq = c.prev[:, :, :3].cpu().numpy()
        for e, (nu, op, oq, cov, tear) in enumerate(res):
            total += 1; sub += nu
            ok = nu == int(host["sim_steps"][e]) and np.array_equal(gp[e], op) and np.array_equal(gq[e], oq) and abs(cov - host["coverage"][e]) < 1e-12 \
                and tear == bool(host["flags"][e] & 1)
            if not ok:
                bad += 1
                print("MISMATCH env %d step %d: substeps %d vs %d, max |dpos| %.3e" % (e, t, nu, int(host["sim_steps"][e]), np.abs(gp[e] - op).max()))
    print("fuzz f64 tier %d: %d env-steps (%d substeps) compared bit for bit with the oracle, %d mismatches" % (tier, total, sub, bad))
    sys.exit(1 if bad else 0)
