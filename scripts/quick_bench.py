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
    P = L.default_params()
    if os.environ.get("W64"):
        P.num_width_points = P.num_height_points = 64; P.thickness = 0.006
    P.reserved0 = int(os.environ.get("RELAX", "1"))
    bc = BatchedCloth(P, n, dtype=dtype, mode=int(os.environ.get("MODE", "0")))
    bc.schedule = os.environ.get("NOSCHED") is None
    a0 = torch.from_numpy(actions(rng, n)).to("cuda", dtype)
    bc.step_actions(a0); torch.cuda.synchronize()          # crumple (untimed)
    res = []
    for r in range(reps):
        a = torch.from_numpy(actions(rng, n)).to("cuda", dtype)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); bc.step_actions(a); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); sub = int(bc.sim_steps.sum().item())
        per_env_ms = (bc.cost.double() * bc.sim_steps.double() / 1.965e6).cpu().numpy()
        print("   per-env busy time: max %.1f ms, p99 %.1f, mean(active) %.1f; sum/slots(1184) %.1f ms; kernel %.1f ms" % (
            per_env_ms.max(), np.percentile(per_env_ms, 99), per_env_ms[per_env_ms > 0].mean(), per_env_ms.sum() / 1184, ms)) if not os.environ.get('QUIET') else None
        res.append((ms, sub, sub / ms * 1e3, n / ms * 1e3, int(((bc.flags & 4) != 0).sum().item())))
    return res, bc

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    what = sys.argv[2] if len(sys.argv) > 2 else "time"
    print("NT", os.environ.get("CLOTHB200_NT", "128"), "n", n)
    if what == "time":
        for dt in ((torch.float32,) if os.environ.get("F32ONLY") else (torch.float32, torch.float64)):
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
    elif what in ("phases", "benchphases"):
        import ctypes as C
        dt = torch.float32 if (len(sys.argv) < 4 or sys.argv[3] == "f32") else torch.float64
        rng = np.random.RandomState(0)
        if what == "phases":
            bc = BatchedCloth(L.default_params(), n, dtype=dt, mode=int(os.environ.get("MODE", "0")))
            a0 = torch.from_numpy(actions(rng, n)).to("cuda", dt)
            bc.step_actions(a0); torch.cuda.synchronize()
            a = torch.from_numpy(actions(rng, n)).to("cuda", dt)
            step = lambda: bc.step_actions(a)
        else:
            # the contract bench's workload (bench.py): tier-1 reset pool, warm-up steps with restarts, then one profiled step
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            import bench as B
            from gym_cloth_b200 import cfg_path
            from gym_cloth_b200.envs import BatchedClothEnv
            env = BatchedClothEnv(cfg_path(1), n, dtype="f32" if dt == torch.float32 else "f64", seed=1337)
            env.reset(); pool = env.snapshot(); bc = env.cloth
            for t in range(3):
                env.step(torch.from_numpy(B.actions_for_step(1337, t, 0, n)).to("cuda", dt))
                done = torch.nonzero(bc.done)[:, 0]
                if done.numel():
                    g = np.random.Generator(np.random.Philox(key=1338, counter=[t, 0, 0, 0]))
                    env.reset_from_pool(pool, done, torch.from_numpy(g.integers(0, n, size=int(done.numel()))).to("cuda"))
            a = torch.from_numpy(B.actions_for_step(1337, 3, 0, n)).to("cuda", dt)
            step = lambda: env.step(a)
            torch.cuda.synchronize()
        if os.environ.get("NCU"):
            # ncu --profile-from-start off ... : capture exactly this launch, without the cycle counters
            torch.cuda.profiler.start(); step(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
            print("substeps", int(bc.sim_steps.sum().item()))
            sys.exit(0)
        prof = torch.zeros(n, 16, dtype=torch.int64, device="cuda")
        L.lib().clothb200_debug_set_profile(C.c_void_p(prof.data_ptr()))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        L.lib().clothb200_debug_set_profile(None)
        p = prof.cpu().numpy().astype(np.float64)
        act = p[:, 10] > 0
        names = ["hooke_verlet", "commit_hash", "alloc", "scatter", "(coll: grab, 4 warps)", "(unused)", "(coll: work, 4 warps)", "collide_buckets", "limit_snap", "limit_resolve"]
        nsub = p[act, 10].sum()
        print("  collide_buckets: near members/substep %.1f, replay rounds/substep %.1f (warp groups x subjects)" % (p[act, 4].sum() / nsub, p[act, 6].sum() / nsub))
        p[:, 4] = 0; p[:, 6] = 0                         # event counters, not cycles (5 = piles: small, left in)
        tot = p[act, :10].sum()
        print("%s: %.1f ms; active envs %d; substeps %d; cycles/substep/CTA %.0f" % (str(dt), e0.elapsed_time(e1), act.sum(), nsub, tot / nsub))
        for i, nm in enumerate(names):
            print("  %-18s %6.1f %%  %9.0f cyc/substep" % (nm, 100 * p[act, i].sum() / tot, p[act, i].sum() / nsub))
        print("  members of <=32-point buckets/substep %.1f; ordered pair tests P/substep %.0f (flat cloth: 4704); >32-point buckets/substep %.2f" % (p[act, 14].sum() / nsub, p[act, 15].sum() / nsub, p[act, 5].sum() / nsub))
        print("  replay buckets/substep %.2f  limit pops/substep %.2f  shortened/substep %.2f" % (p[act, 11].sum() / nsub, p[act, 12].sum() / nsub, p[act, 13].sum() / nsub))
        per = p[act, :10].sum(1) / p[act, 10]
        print("  per-env cycles/substep percentiles 10/50/90/100:", np.percentile(per, [10, 50, 90, 100]).round(0))
        print("  per-env pops/substep percentiles 50/90/100:", np.percentile(p[act, 12] / p[act, 10], [50, 90, 100]).round(1))
        heavy = act & (np.where(act, p[:, :10].sum(1) / np.maximum(p[:, 10], 1), 0) >= np.percentile(per, 95))
        toth = p[heavy, :10].sum(); nsh = p[heavy, 10].sum()
        print("  heaviest 5%% of envs (%d): cycles/substep %.0f, pops/substep %.1f, replay buckets/substep %.1f" % (
            heavy.sum(), toth / nsh, p[heavy, 12].sum() / nsh, p[heavy, 11].sum() / nsh))
        for i, nm in enumerate(names):
            print("    %-18s %6.1f %%  %9.0f cyc/substep" % (nm, 100 * p[heavy, i].sum() / toth, p[heavy, i].sum() / nsh))
        idx = np.argsort(-np.where(act, p[:, :10].sum(1) / np.maximum(p[:, 10], 1), 0))
        print("  env  substeps  cyc/substep  pops/sub  moved/sub  limit_replay cyc/pop  buckets/sub  coll_replay cyc/bucket | small members/sub, big members/sub, big buckets/sub | phase cycles/sub: coll_replay limit_snap limit_replay")
        for e in list(idx[:8]) + list(idx[200:204]) + list(idx[600:604]):
            ns = p[e, 10]
            print("  %4d %8d %11.0f %9.1f %9.1f %12.0f %12.1f %12.0f" % (e, ns, p[e, :10].sum() / ns, p[e, 12] / ns, p[e, 13] / ns,
                  p[e, 9] / max(p[e, 12], 1), p[e, 11] / ns, p[e, 7] / max(p[e, 11], 1)),
                  "| %6.1f %6.1f %5.2f | %7.0f %7.0f %7.0f" % (p[e, 14] / ns, p[e, 15] / ns, p[e, 5] / ns, p[e, 7] / ns, p[e, 8] / ns, p[e, 9] / ns))
    elif what == "sched":
        # how much of the launch time is mis-prediction? run the same action from the same state three ways:
        # predicted costs (what a user gets), exact costs (from the first run), no scheduling
        dt = torch.float32
        rng = np.random.RandomState(0)
        bc = BatchedCloth(L.default_params(), n, dtype=dt)
        for _ in range(int(os.environ.get("PRE", "2"))):
            bc.step_actions(torch.from_numpy(actions(rng, n)).to("cuda", dt))
        torch.cuda.synchronize()
        keep = [t.clone() for t in (bc.pos, bc.prev, bc.flags, bc.cost)]
        a = torch.from_numpy(actions(rng, n)).to("cuda", dt)
        def run(label, cost=None, sched=True):
            bc.pos.copy_(keep[0]); bc.prev.copy_(keep[1]); bc.flags.copy_(keep[2]); bc.cost.copy_(keep[3] if cost is None else cost)
            bc.schedule = sched
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); bc.step_actions(a); e1.record(); torch.cuda.synchronize()
            busy = (bc.cost.double() * bc.sim_steps.double() / 1.9e6).cpu().numpy()
            print("%-16s kernel %.1f ms | per-env busy max %.1f p99 %.1f mean(active) %.1f | sum/slots %.1f | active %d" % (
                label, e0.elapsed_time(e1), busy.max(), np.percentile(busy, 99), busy[busy > 0].mean(), busy.sum() / 1184, (busy > 0).sum()))
            return bc.cost.clone()
        c1 = run("predicted")
        c2 = run("exact cost", cost=c1)
        run("exact cost again", cost=c2)
        run("no scheduling", sched=False)
        steps = bc.sim_steps.float()
        run("key = substeps", cost=torch.full_like(c1, 1.0e5))
        run("key = substeps^2", cost=steps.clamp(min=1.0) * 60.0)
        run("key = substeps^3", cost=steps.clamp(min=1.0) ** 2 * 0.03)
        pred = (keep[3].double() * bc.sim_steps.double()).cpu().numpy(); act = (c1.double() * bc.sim_steps.double()).cpu().numpy()
        m = act > 0
        print("prediction error (active envs): corr %.3f, |pred/act-1| p50 %.2f p90 %.2f; envs with unknown cost %d" % (
            np.corrcoef(pred[m], act[m])[0, 1], np.percentile(np.abs(pred[m] / act[m] - 1), 50), np.percentile(np.abs(pred[m] / act[m] - 1), 90),
            int((keep[3][torch.from_numpy(m).cuda()] <= 0).sum().item())))
    elif what == "timeline":
        import ctypes as C
        dt = torch.float32
        rng = np.random.RandomState(0)
        bc = BatchedCloth(L.default_params(), n, dtype=dt)
        bc.schedule = os.environ.get("NOSCHED") is None
        a0 = torch.from_numpy(actions(rng, n)).to("cuda", dt)
        bc.step_actions(a0); torch.cuda.synchronize()
        prof = torch.zeros(n, 16, dtype=torch.int64, device="cuda")
        L.lib().clothb200_debug_set_profile(C.c_void_p(prof.data_ptr()))
        a = torch.from_numpy(actions(rng, n)).to("cuda", dt)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); bc.step_actions(a); e1.record(); torch.cuda.synchronize()
        L.lib().clothb200_debug_set_profile(None)
        p = prof.cpu().numpy()
        t0 = p[:, 14].min(); st = (p[:, 14] - t0) / 1e6; en = (p[:, 15] - t0) / 1e6
        print("kernel %.1f ms; span %.1f ms" % (e0.elapsed_time(e1), en.max()))
        for t in np.linspace(0, en.max(), 21):
            act = ((st <= t) & (en > t))
            heavy = act & (p[:, 10] > 0)
            print("  t=%6.1f ms  resident CTAs %4d  (working %4d)" % (t, act.sum(), heavy.sum()))
        order = np.argsort(st)
        w = p[:, 10] > 0
        print("start time of working CTAs: p50 %.1f p90 %.1f max %.1f ms; their durations p50 %.1f p90 %.1f max %.1f" % (
            np.percentile(st[w], 50), np.percentile(st[w], 90), st[w].max(), np.percentile((en - st)[w], 50), np.percentile((en - st)[w], 90), (en - st)[w].max()))
        last = np.argsort(-en)[:5]
        for e in last:
            print("  late finisher env %d: start %.1f end %.1f substeps %d sm %d" % (e, st[e], en[e], p[e, 10], p[e, 11] // 1000000))
        sm = p[:, 11] // 1000000
        per_sm = np.array([(en - st)[sm == k].sum() for k in range(int(sm.max()) + 1)])
        print("busy CTA-ms per SM: min %.0f mean %.0f max %.0f (6 slots x span = %.0f)" % (per_sm.min(), per_sm.mean(), per_sm.max(), 6 * en.max()))
