"""Exploration script (not the contract bench): time one action over n envs for f32/f64 and thread counts,
and measure f32-vs-f64 divergence.  Usage: python scripts/quick_bench.py [n_env] [what]"""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_cloth_b200 import lib as L
from gym_cloth_b200.batched import BatchedCloth

def actions(rng, n):
    return rng.uniform(-1, 1, size=(n, 4))

def timed_action(n, dtype, seed=0, reps=2):
    rng = np.random.RandomState(seed)
    bc = BatchedCloth(L.default_params(), n, dtype=dtype)
    a0 = torch.from_numpy(actions(rng, n)).to("cuda", dtype)
    bc.step_actions(a0); torch.cuda.synchronize()          # crumple (untimed)
    res = []
    for r in range(reps):
        a = torch.from_numpy(actions(rng, n)).to("cuda", dtype)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); bc.step_actions(a); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); sub = int(bc.sim_steps.sum().item())
        res.append((ms, sub, sub / ms * 1e3, n / ms * 1e3, int(((bc.flags & 4) != 0).sum().item())))
    return res, bc

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    what = sys.argv[2] if len(sys.argv) > 2 else "time"
    print("NT", os.environ.get("CLOTHB200_NT", "128"), "n", n)
    if what == "time":
        for dt in (torch.float32, torch.float64):
            res, bc = timed_action(n, dt)
            for ms, sub, sps, eps, ng in res:
                print("%s: %.1f ms, %d substeps, %.3e substeps/s, %.1f env-steps/s, nograb %d" % (str(dt), ms, sub, sps, eps, ng))
    elif what == "diverge":
        rng = np.random.RandomState(1)
        P = L.default_params()
        a32 = BatchedCloth(P, n, dtype=torch.float32); a64 = BatchedCloth(P, n, dtype=torch.float64)
        for step in range(3):
            acts = actions(rng, n); acts[:, :2] *= 0.9
            # shared start state: copy f64 state (rounded) into the f32 batch before every action
            a32.pos.copy_(a64.pos.float()); a32.prev.copy_(a64.prev.float()); a32.flags.copy_(a64.flags)
            a32.step_host(acts, {}); a64.step_host(acts, {}); torch.cuda.synchronize()
            d = (a32.pos.double() - a64.pos)[:, :, :3].abs()
            mx = d.amax(dim=(1, 2)).cpu().numpy(); mn = d.mean(dim=(1, 2)).cpu().numpy()
            dc = (a32.coverage - a64.coverage).abs().cpu().numpy()
            same = (a32.sim_steps == a64.sim_steps).float().mean().item()
            sameg = (a32.n_grabbed == a64.n_grabbed).float().mean().item()
            q = lambda x: " ".join("%.2e" % v for v in np.percentile(x, [50, 90, 99, 100]))
            print("action %d: max|d| p50/90/99/100: %s | mean|d|: %s | dcov: %s | same substeps %.3f same ngrab %.3f | cov mean f32 %.4f f64 %.4f" % (
                step, q(mx), q(mn), q(dc), same, sameg, a32.coverage.mean().item(), a64.coverage.mean().item()))
    elif what == "phases":
        import ctypes as C
        dt = torch.float32 if (len(sys.argv) < 4 or sys.argv[3] == "f32") else torch.float64
        rng = np.random.RandomState(0)
        bc = BatchedCloth(L.default_params(), n, dtype=dt)
        a0 = torch.from_numpy(actions(rng, n)).to("cuda", dt)
        bc.step_actions(a0); torch.cuda.synchronize()
        prof = torch.zeros(n, 16, dtype=torch.int64, device="cuda")
        L.lib().clothb200_debug_set_profile(C.c_void_p(prof.data_ptr()))
        a = torch.from_numpy(actions(rng, n)).to("cuda", dt)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); bc.step_actions(a); e1.record(); torch.cuda.synchronize()
        L.lib().clothb200_debug_set_profile(None)
        p = prof.cpu().numpy().astype(np.float64)
        act = p[:, 10] > 0
        names = ["hooke_verlet", "commit_hash", "alloc", "scatter", "order", "coll_snap", "coll_first_plane", "coll_replay", "limit_snap", "limit_replay"]
        tot = p[act, :10].sum()
        nsub = p[act, 10].sum()
        print("%s: %.1f ms; active envs %d; substeps %d; cycles/substep/CTA %.0f" % (str(dt), e0.elapsed_time(e1), act.sum(), nsub, tot / nsub))
        for i, nm in enumerate(names):
            print("  %-18s %6.1f %%  %9.0f cyc/substep" % (nm, 100 * p[act, i].sum() / tot, p[act, i].sum() / nsub))
        print("  replay buckets/substep %.2f  limit pops/substep %.2f  shortened/substep %.2f" % (p[act, 11].sum() / nsub, p[act, 12].sum() / nsub, p[act, 13].sum() / nsub))
        per = p[act, :10].sum(1) / p[act, 10]
        print("  per-env cycles/substep percentiles 10/50/90/100:", np.percentile(per, [10, 50, 90, 100]).round(0))
        print("  per-env pops/substep percentiles 50/90/100:", np.percentile(p[act, 12] / p[act, 10], [50, 90, 100]).round(1))
