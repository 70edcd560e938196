#!/usr/bin/env python
"""bench.py - BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout (rank 0).

    python bench.py --gpus N --steps K --warmup W            our arm (B200 kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU physics on the host cores

metric   cloth env-steps/s: one env-step = one ClothEnv.step(action) = grab + all substeps
         (1430 + iters_pull Cloth.update() calls) + coverage/reward/terminal.
workload BASELINE.json configs[1]: 4096 batched tier-1 25x25 cloths PER GPU (weak scaling), random pull actions
         a ~ U[-1,1]^4 keyed by (seed, step, global env id), reference-order mode.  Start states come from the
         tier-1 reset (2-3 random short pulls, cloth_env.py:843-891) run on the device before timing; an env
         that reports done is re-started from that pool of reset states (device copy, inside the timed region).
value    whole-job env-steps/s with actions already resident in HBM (CUDA events on the launching stream,
         barrier + synchronize on both sides, max over ranks).
e2e      the same steps through the host-buffer C-ABI call clothb200_step_host_* (actions in pinned host memory,
         decode + H2D of the plans, kernel, D2H of obs/reward/done/coverage/flags) - the call ClothEnv.step makes.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cloth env-steps/sec (25x25 tier-1, reference-order mode)"
UNIT = "env-steps/s"
N_POINTS, N_SPRINGS, P_FLAT = 625, 3502, 4704
FLOP_PER_SUBSTEP = 32 * N_SPRINGS + 26 * N_POINTS + 9 * P_FLAT        # 170 650
SMEM_B_PER_SUBSTEP = 80 * N_SPRINGS + 120 * N_POINTS + 12 * P_FLAT    # 411 608
HBM_B_PER_ENV_STEP = 2 * 2 * N_POINTS * 16 + 3 * N_POINTS * 4         # pos+prev load+store (float4) + obs


def actions_for_step(seed, t, lo, hi):
    """U[-1,1]^4 for global env ids lo..hi-1 at step t - independent of how envs are sharded."""
    out = np.empty((hi - lo, 4))
    blk = 1024
    for b in range(lo // blk, (hi + blk - 1) // blk):
        g = np.random.Generator(np.random.Philox(key=seed, counter=[t, b, 0, 0]))
        a = g.uniform(-1.0, 1.0, size=(blk, 4))
        s, e = max(lo, b * blk), min(hi, (b + 1) * blk)
        out[s - lo:e - lo] = a[s - b * blk:e - b * blk]
    return out


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) >= 8 and r[4 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from gym_cloth_b200 import cfg_path, lib as L
    from gym_cloth_b200.envs import BatchedClothEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = L.lib()
    n = args.envs
    lo, hi = rank * n, (rank + 1) * n
    dtype = "f32" if args.dtype == "f32" else "f64"
    tdt = torch.float32 if dtype == "f32" else torch.float64
    env = BatchedClothEnv(cfg_path(1), n, dtype=dtype, seed=args.seed, env_offset=lo)
    t_reset0 = time.perf_counter()
    env.reset()
    torch.cuda.synchronize()
    reset_s = time.perf_counter() - t_reset0
    pool = env.snapshot()
    c = env.cloth
    K, W = args.steps, args.warmup

    def restart_done(t):
        done = torch.nonzero(c.done)[:, 0]
        if done.numel():
            g = np.random.Generator(np.random.Philox(key=args.seed + 1, counter=[t, rank, 0, 0]))
            choice = torch.from_numpy(g.integers(0, n, size=int(done.numel()))).to(c.device)
            env.reset_from_pool(pool, done, choice)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: `value` ----------------
    dev_actions = [torch.from_numpy(actions_for_step(args.seed, t, lo, hi)).to(c.device, tdt) for t in range(W + K)]
    for t in range(W):
        env.step(dev_actions[t]); restart_done(t)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    launches0 = lib.clothb200_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 2)]
    substeps = 0
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=c.device)   # > 126 MB L2
    ev[0].record()
    kernel_ms_events = []
    for t in range(K):
        flush.zero_()                            # L2 flush between timed iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(dev_actions[W + t]); b.record()
        kernel_ms_events.append((a, b))
        substeps_t = c.sim_steps.sum()          # device-side, read after the timed region
        substeps = substeps_t if t == 0 else substeps + substeps_t
        restart_done(W + t)
    ev[1].record()
    barrier()
    busy_cycles = (c.cost.double() * c.sim_steps.double()).clone()      # last timed launch: SM cycles each cloth was worked on
    launches = lib.clothb200_launch_count() - launches0
    elapsed_ms = ev[0].elapsed_time(ev[1])
    kernel_ms = sum(a.elapsed_time(b) for a, b in kernel_ms_events)
    substeps = int(substeps.item())
    clocks = sampler.stop()

    # ---------------- end-to-end arm: `e2e` ----------------
    host = env.host_buffers()
    host_actions = [actions_for_step(args.seed, 1000 + t, lo, hi) for t in range(W + K)]
    for t in range(min(W, 1)):
        env.step(host_actions[t], host_out=host); restart_done(2000 + t)
    barrier()
    t0 = time.perf_counter()
    for t in range(K):
        obs, rew, done, info = env.step(host_actions[min(W, 1) + t], host_out=host)
        _ = float(rew[0])                        # the caller reads its result
        restart_done(3000 + t)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    plan_b = C.sizeof(L.Plan)
    esz = 4 if dtype == "f32" else 8
    h2d = n * plan_b
    d2h = n * (3 * N_POINTS * esz + 8 + 4 + 8 + 8 + 4 + 4)

    # ---------------- max over ranks ----------------
    tt = torch.tensor([elapsed_ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device=c.device)
    cnt = torch.tensor([float(substeps)], dtype=torch.float64, device=c.device)
    cov_stats = torch.stack([c.coverage.sum(), torch.tensor(float(n), device=c.device, dtype=torch.float64)])
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(cov_stats, op=dist.ReduceOp.SUM)      # the optional episode-statistics gather (NCCL)
    elapsed_ms, e2e_ms, kernel_ms = (float(v) for v in tt.tolist())
    total_substeps = float(cnt.item())

    if rank == 0:
        n_total = n * world
        value = n_total * K / (elapsed_ms * 1e-3)
        sub_per_s = total_substeps / (elapsed_ms * 1e-3)
        peaks, peak_src = measured_peaks()
        smem_peak = C.c_double(0); fp32_peak = C.c_double(0)
        lib.clothb200_bench_smem_bandwidth(2000, C.byref(smem_peak), None)
        lib.clothb200_bench_fp32_flops(2000, C.byref(fp32_peak), None)
        per_gpu_sub_per_s = (total_substeps / world) / (kernel_ms * 1e-3)
        smem_ach = per_gpu_sub_per_s * SMEM_B_PER_SUBSTEP / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "step_kernel_dram_bytes_per_env_step.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp))["bytes_per_env_step"] * n
            except Exception:
                traffic = None
        ctas = C.c_int(0); smemb = C.c_int(0); thr = C.c_int(0)
        lib.clothb200_occupancy(C.byref(env.P), int(dtype == "f64"), C.byref(ctas), C.byref(smemb), C.byref(thr))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: 4096 batched tier-1 25x25 cloths per GPU, random pull actions U[-1,1]^4, "
                                   "reference-order mode", "envs_per_gpu": n, "envs_total": n_total, "grid": "25x25",
                       "mode": "reference_order", "threads_per_cloth": thr.value, "smem_bytes_per_cloth": smemb.value,
                       "resident_cloths_per_sm": ctas.value, "reset": "tier-1 reset on device before timing (%.1f s); done envs "
                       "restart from that pool" % reset_s,
                       "l2_flush": "256 MiB buffer written between timed iterations",
                       "l2": "per-step working set = %d MiB of state written+read once per launch (cloths live in shared memory "
                             "while they are worked on; HBM/L2 see one load and one store per env-step plus 40 KB per swap of the "
                             "time-sliced launch)" % (n * 20000 // 2 ** 20)},
            "substeps_per_s": sub_per_s, "substeps_per_env_step": total_substeps / (n_total * K),
            "e2e": {"value": n_total * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms / K},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "smem", "achieved": smem_ach, "peak": smem_peak.value, "unit": "GB/s",
                         "frac": smem_ach / smem_peak.value if smem_peak.value else None, "traffic": traffic,
                         "kernel": "cloth_step_kernel", "kernel_ms_per_launch": kernel_ms / K,
                         "algorithmic_bytes_per_substep": SMEM_B_PER_SUBSTEP, "substeps_per_launch": total_substeps / world / K,
                         "peak_source": "LDS.128 streaming microbenchmark run in this process (clothb200_bench_smem_bandwidth); "
                                        "MEASURED_PEAKS.json has no shared-memory figure",
                         "note": "BASELINE.json's metric names the shared-memory/FP32 roofline for this path; HBM is not the bound"},
            "roofline_fp32": {"achieved": per_gpu_sub_per_s * FLOP_PER_SUBSTEP / 1e12, "peak": fp32_peak.value, "unit": "TFLOP/s",
                              "frac": per_gpu_sub_per_s * FLOP_PER_SUBSTEP / 1e12 / fp32_peak.value if fp32_peak.value else None,
                              "peak_source": "FFMA microbenchmark run in this process"},
            "roofline_hbm": {"achieved": (n * K / (kernel_ms * 1e-3)) * HBM_B_PER_ENV_STEP / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": (n * K / (kernel_ms * 1e-3)) * HBM_B_PER_ENV_STEP / 1e9 / peaks["hbm_gbs"], "peak_source": peak_src},
            "mean_coverage": float(cov_stats[0].item() / cov_stats[1].item()),
        }
        # why a launch takes as long as it does: it cannot end before its slowest cloth nor before total work / resident slots
        mhz = float(clocks.get("sm_mhz") or 1965.0)
        slots = ctas.value * torch.cuda.get_device_properties(c.device).multi_processor_count
        line["launch_balance"] = {"slowest_cloth_ms": float(busy_cycles.max().item()) / (mhz * 1e3),
                                  "work_per_slot_ms": float(busy_cycles.sum().item()) / (mhz * 1e3) / slots,
                                  "active_cloths": int((busy_cycles > 0).sum().item()), "resident_slots": int(slots),
                                  "note": "rank 0, last timed launch; cycles spent inside the substep loop only"}
        if world == 1 and not args.no_extras:
            line["other_builds"] = other_builds(args, pool, dev_actions[W:W + 2])
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(env, pool, args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def other_builds(args, pool, actions):
    """Same workload, two steps each, for the other two builds of the kernel (reported beside the headline):
    graph-coloured f32 and reference-order f64 (the bit-exact parity build)."""
    import torch
    from gym_cloth_b200 import lib as L
    from gym_cloth_b200.batched import BatchedCloth
    out = {}
    for name, dt, mode in (("coloured_f32", torch.float32, L.MODE_COLOURED), ("reference_order_f64", torch.float64, L.MODE_REFERENCE_ORDER)):
        bc = BatchedCloth(L.default_params(), args.envs, dtype=dt, mode=mode)
        bc.pos.copy_(pool["pos"].to(dt)); bc.prev.copy_(pool["prev"].to(dt))
        bc.step_actions(actions[0].to(dt)); torch.cuda.synchronize()        # warm-up (also makes states diverge from the pool)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bc.step_actions(actions[1].to(dt)); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"env_steps_per_s": args.envs / (ms * 1e-3), "substeps_per_s": float(bc.sim_steps.sum().item()) / (ms * 1e-3),
                     "ms_per_step": ms, "steps": 1, "warmup": 1}
    return out


def _pool_states(pool, k):
    pos = pool["pos"][:k, :, :3].double().cpu().numpy(); prev = pool["prev"][:k, :, :3].double().cpu().numpy()
    return [(pos[i], prev[i]) for i in range(k)]


def cpu_baseline(env, pool, args):
    """The reference's CPU path beside the GPU number: one env.step per host core, same start states
    (the first `cores` pool states) and the same action generator."""
    from oracle.build_ref import ref_built
    from oracle.ref_driver import cpu_env_steps
    cores = os.cpu_count() or 1
    kind = "reference" if ref_built() else "port"
    k = cores if kind == "reference" else 8 * cores
    states = _pool_states(pool, min(k, env.n_env))
    acts = actions_for_step(args.seed, 0, 0, len(states))
    # make sure every sampled action does work: aim at a cloth point of its own state
    for i, (pos, _) in enumerate(states):
        acts[i, 0] = (pos[(37 * i) % N_POINTS, 0] - 0.5) * 2; acts[i, 1] = (pos[(37 * i) % N_POINTS, 1] - 0.5) * 2
    r = cpu_env_steps(kind, states, acts, cores)
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": kind,
            "sample": "%d env.step calls (one per host process, %d substeps in total, %.1f s) from the first %d reset-pool states; "
                      "grip point aimed at a cloth point so every call does work" % (r["n"], r["substeps"], r["seconds"], r["n"]),
            "substeps_per_s": r["substeps_per_s"], "substeps_per_s_per_core": r["substeps_per_s"] / r["cores"]}


def run_reference(args):
    """--impl reference: the reference's own compiled physics (oracle/_ref) on all host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as O
    from oracle.build_ref import ref_built
    from oracle.ref_driver import cpu_env_steps
    O.build()
    cores = os.cpu_count() or 1
    kind = "reference" if ref_built() else "port"
    per_step = cores if kind == "reference" else 8 * cores
    # tier-1 start states: flat grid + two short reset pulls, produced with the fast CPU port (untimed)
    rng = np.random.RandomState(args.seed)
    states = []
    for i in range(per_step):
        o = O.OracleCloth()
        for _ in range(2):
            p = rng.randint(N_POINTS); pos = o.pos
            d = rng.uniform(0.08, 0.2, 2) * rng.choice([-1, 1], 2)
            o.step_action(np.array([(pos[p, 0] - 0.5) * 2, (pos[p, 1] - 0.5) * 2, d[0], d[1]]))
        s = o.get_state()
        states.append((s[0], s[1]))
    K, W = args.steps, min(args.warmup, 1)
    times, subs = [], 0
    for t in range(W + K):
        acts = actions_for_step(args.seed, t, 0, per_step)
        for i, (pos, _) in enumerate(states):
            acts[i, 0] = (pos[(37 * i + t) % N_POINTS, 0] - 0.5) * 2; acts[i, 1] = (pos[(37 * i + t) % N_POINTS, 1] - 0.5) * 2
        r = cpu_env_steps(kind, states, acts, cores)
        if t >= W:
            times.append(r["seconds"]); subs += r["substeps"]
    total = sum(times)
    value = per_step * K / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": total / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] workload, bounded sample: each step = %d env.step calls (one per host process) on "
                                   "tier-1 states with random pull actions" % per_step, "grid": "25x25", "mode": "reference (Cython Gauss-Seidel)"},
            "substeps_per_s": subs / total,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d steps x %d env.step calls, %d substeps, %.1f s" % (K, per_step, subs, total)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--seed", type=int, default=1337)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
