#!/usr/bin/env python
"""bench.py - BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout (rank 0).

    python bench.py --gpus N --steps K --warmup W            our arm (B200 kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU physics on the host cores
    python bench.py --config {3,4,5} ...                      the other BASELINE.json configs (same line format)

metric   cloth env-steps/s: one env-step = one ClothEnv.step(action) = grab + all substeps
         (1430 + iters_pull Cloth.update() calls) + coverage/reward/terminal.
workload BASELINE.json configs[1] (the default, --config 2): 4096 batched tier-1 25x25 cloths PER GPU (weak scaling),
         random pull actions, reference-order mode.  BOTH arms run the same thing:
           * environment i (global id) starts from the tier-1 reset state of env seed `seed + i` (cloth_env.py:843-891);
             our arm runs that reset on the device before timing, the reference arm reads the states the reference's
             own reset() produced for seeds 1337.. (tests/golden/bench_pool_t1.npz, recorded by tests/golden/make_golden.py);
           * action t of environment i is draw_actions(seed, t, i): a Philox draw of (x, y, dx, dy) ~ U[-1,1]^4 plus a mesh
             point index; with --actions touch_cloth (default) the grip (x, y) is aimed at that mesh point of the
             environment's CURRENT state - the reference's get_random_action('touch_cloth') (cloth_env.py:1005-1015) -
             so every env-step grips the cloth and runs its 1430 + iters_pull substeps; --actions over_xy_plane keeps the
             raw draw (action_space.sample(), cloth_env.py:1003-1004; about a third of those grip nothing);
           * an environment whose step ends the episode (tear, out of bounds, coverage > 0.92, max_actions) restarts
             from a reset state of the pool, inside the timed region.
value    whole-job env-steps/s with the drawn actions already resident in HBM (CUDA events on the launching stream,
         barrier + synchronize on both sides, max over ranks).
e2e      the same steps through the host-buffer C-ABI call clothb200_step_host_* (actions in host memory, decode + H2D of
         the plans, kernel, D2H of obs/reward/done/coverage/flags into pinned memory) - the call ClothEnv.step makes; the
         host aims each grip from the observation it got back, as a policy would.
parity_build / coloured_build: the same workload and the same K/W on the bit-exact f64 build and the graph-coloured build.
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
POOL_FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_pool_t1.npz")


def n_springs(w, h):
    return w * (h - 1) + h * (w - 1) + 2 * (w - 1) * (h - 1) + w * (h - 2) + h * (w - 2)


def smem_bytes_per_substep(w, h, pairs):
    """80 B per spring visit, 120 B per point, 12 B per ordered self-collision pair test."""
    return 80 * n_springs(w, h) + 120 * w * h + 12 * pairs


def flop_per_substep(w, h, pairs):
    return 32 * n_springs(w, h) + 26 * w * h + 9 * pairs


def draw_actions(seed, t, lo, hi, n_points=N_POINTS):
    """Step t's draw for global env ids lo..hi-1, independent of how envs are sharded:
    raw [n,4] ~ U[-1,1]^4 and pick [n] ~ U{0..n_points-1}."""
    raw = np.empty((hi - lo, 4)); pick = np.empty(hi - lo, np.int64)
    blk = 1024
    for b in range(lo // blk, (hi + blk - 1) // blk):
        g = np.random.Generator(np.random.Philox(key=seed, counter=[t, b, 0, 0]))
        a = g.uniform(-1.0, 1.0, size=(blk, 4))
        p = g.integers(0, n_points, size=blk)
        s, e = max(lo, b * blk), min(hi, (b + 1) * blk)
        raw[s - lo:e - lo] = a[s - b * blk:e - b * blk]; pick[s - lo:e - lo] = p[s - b * blk:e - b * blk]
    return raw, pick


def actions_for_step(seed, t, lo, hi):
    """The raw U[-1,1]^4 draw alone (--actions over_xy_plane)."""
    return draw_actions(seed, t, lo, hi)[0]


def restart_choice(seed, t, lo, hi, n_pool):
    """Pool state that global env ids lo..hi-1 restart from if their episode ends at step t (independent of the sharding)."""
    out = np.empty(hi - lo, np.int64)
    blk = 1024
    for b in range(lo // blk, (hi + blk - 1) // blk):
        g = np.random.Generator(np.random.Philox(key=seed + 1, counter=[t, b, 0, 0]))
        c = g.integers(0, n_pool, size=blk)
        s, e = max(lo, b * blk), min(hi, (b + 1) * blk)
        out[s - lo:e - lo] = c[s - b * blk:e - b * blk]
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


# ---------------------------------------------------------------------------------------------------------------- workloads
def workload(args, world):
    """(name, description, scaling, envs per GPU, grid W) of --config."""
    c = args.config
    if c == 2:
        n = args.envs or 4096
        return ("BASELINE configs[1]: %d batched tier-1 25x25 cloths per GPU, random pull actions (%s), reference-order mode"
                % (n, args.actions), "weak", n, 25)
    if c == 3:
        n = args.envs or 16384
        return ("BASELINE configs[2]: %d envs per GPU in ONE launch, half from tier-2 and half from tier-3 reset states, self-collision "
                "on, random pull actions (%s), reference-order mode" % (n, args.actions), "weak", n, 25)
    if c == 4:
        n = args.envs or 2368
        return ("BASELINE configs[3]: %d 64x64 cloths per GPU (thickness 0.006), graph-coloured Gauss-Seidel mode, %d limit passes per "
                "substep, random pull actions (%s)" % (n, args.relax_iters, args.actions), "weak", n, 64)
    if c == 5:
        tot = args.envs or 65536
        assert tot % world == 0
        return ("BASELINE configs[4]: %d tier-1 25x25 envs in total sharded over %d GPU(s), coverage reward, random pull actions (%s), "
                "reference-order mode" % (tot, world, args.actions), "strong", tot // world, 25)
    raise SystemExit("--config must be 2, 3, 4 or 5")


class Arm(object):
    """One build of the kernel on one rank's shard of a workload: cloth batch + pool of restart states + action draws."""

    def __init__(self, args, rank, n, dtype, mode, pool=None):
        import torch
        from gym_cloth_b200 import cfg_path, lib as L
        from gym_cloth_b200.batched import BatchedCloth
        from gym_cloth_b200.envs import BatchedClothEnv
        self.torch = torch
        self.args, self.rank, self.n = args, rank, n
        self.lo, self.hi = rank * n, (rank + 1) * n
        self.tdt = torch.float32 if dtype == "f32" else torch.float64
        self.dtype = dtype
        self.reset_s = 0.0
        cfgn = args.config
        t0 = time.perf_counter()
        if cfgn in (2, 5):
            env = BatchedClothEnv(cfg_path(1), n, dtype=dtype, seed=args.seed, env_offset=self.lo, mode=mode)
            self.c = env.cloth
            if pool is None:
                env.reset(); pool = env.snapshot()
        elif cfgn == 3:
            # ONE launch holds both tiers: the batch carries per-environment rest lengths (tier 2 perturbs the grid it is
            # built from, cloth.pyx:101-107, so its springs have their own), nominal ones for the tier-3 half
            env = BatchedClothEnv(cfg_path(2), n, dtype=dtype, seed=args.seed, env_offset=self.lo, mode=mode)
            self.c = c = env.cloth
            if pool is None:
                h = n // 2
                env.reset(envs=np.arange(h))
                e3 = BatchedClothEnv(cfg_path(3), n - h, dtype=dtype, seed=args.seed, env_offset=self.lo + h, mode=mode)
                e3.reset()
                c.pos[h:] = e3.cloth.pos; c.prev[h:] = e3.cloth.prev; c.coverage[h:] = e3.cloth.coverage
                c.rest[h:] = BatchedCloth(c.P, 1, dtype=self.tdt, exact_rest=True).rest.reshape(1, -1)
                del e3
                pool = {"pos": c.pos.clone(), "prev": c.prev.clone(), "cov": c.coverage.clone(), "rest": c.rest.clone()}
        else:
            P = L.default_params()
            P.num_width_points = P.num_height_points = 64; P.thickness = 0.006; P.reserved0 = args.relax_iters
            self.c = BatchedCloth(P, n, dtype=self.tdt, mode=mode)
            if pool is None:
                self.c.measure()
                pool = {"pos": self.c.pos[:1].clone(), "prev": self.c.prev[:1].clone(), "cov": self.c.coverage[:1].clone()}
        torch.cuda.synchronize()
        self.reset_s = time.perf_counter() - t0
        self.pool = {k: (v.to(self.tdt) if k != "cov" else v) for k, v in pool.items() if v is not None and (k != "rest" or cfgn == 3)}
        c = self.c
        if "rest" in self.pool:
            c.rest.copy_(self.pool["rest"])
        c.pos.copy_(self.pool["pos"][torch.arange(n, device=c.device) % self.pool["pos"].shape[0]])
        c.prev.copy_(self.pool["prev"][torch.arange(n, device=c.device) % self.pool["prev"].shape[0]])
        c.prev_coverage.copy_(self.pool["cov"][torch.arange(n, device=c.device) % self.pool["cov"].shape[0]])
        c.flags.zero_(); c.num_steps.zero_(); c.num_sim_steps.zero_()
        self.n_pool = int(self.pool["pos"].shape[0])
        self.np_ = c.N
        self.ar = torch.arange(n, device=c.device)
        self.aim = args.actions == "touch_cloth"

    # -- pool restarts
    def choice_for(self, t):
        return self.torch.from_numpy(restart_choice(self.args.seed, t, self.lo, self.hi, self.n_pool)).to(self.c.device)

    def restart(self, t, done_idx, choice=None):
        """Environments done_idx (device index tensor) restart from their pool states of step t."""
        c, torch = self.c, self.torch
        k = (self.choice_for(t) if choice is None else choice)[done_idx]
        c.pos[done_idx] = self.pool["pos"][k]; c.prev[done_idx] = self.pool["prev"][k]
        c.prev_coverage[done_idx] = self.pool["cov"][k]
        if "rest" in self.pool:
            c.rest[done_idx] = self.pool["rest"][k]
        c.flags[done_idx] = 0; c.num_steps[done_idx] = 0; c.num_sim_steps[done_idx] = 0
        return k

    # -- `value`: actions resident on the device
    def device_actions(self, t):
        raw, pick = draw_actions(self.args.seed, t, self.lo, self.hi, self.np_)
        return self.torch.from_numpy(raw).to(self.c.device, self.tdt), self.torch.from_numpy(pick).to(self.c.device)

    def device_step(self, t, drawn):
        c = self.c
        raw, pick = drawn
        if self.aim:
            a = raw.clone()
            a[:, :2] = (c.pos[self.ar, pick, :2] - 0.5) * 2
        else:
            a = raw
        c.step_actions(a.contiguous())
        return a

    def run_device(self, K, W, barrier, flush=None, sampler=None, t_base=0):
        torch, c = self.torch, self.c
        drawn = [self.device_actions(t_base + t) for t in range(W + K)]
        choice = [self.choice_for(t_base + t) for t in range(W + K)]         # like the actions: resident before the clock starts
        for t in range(W):
            self.device_step(t, drawn[t])
            d = torch.nonzero(c.done)[:, 0]
            if d.numel():
                self.restart(t_base + t, d, choice[t])
        barrier()
        if sampler is not None:
            sampler.start()
        L0 = c.L.clothb200_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pairs = []
        sub = torch.zeros((), dtype=torch.int64, device=c.device); nog = torch.zeros_like(sub); dn = torch.zeros_like(sub)
        e0.record()
        for t in range(K):
            if flush is not None:
                flush.zero_()                        # L2 flush between timed iterations
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); self.device_step(W + t, drawn[W + t]); b.record()
            pairs.append((a, b))
            sub += c.sim_steps.sum(); nog += ((c.flags & 4) != 0).sum(); dn += (c.done != 0).sum()
            d = torch.nonzero(c.done)[:, 0]            # (a masked, sync-free restart was measured: 1 % slower end to end)
            if d.numel():
                self.restart(t_base + W + t, d, choice[W + t])
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler is not None else None
        return {"elapsed_ms": e0.elapsed_time(e1), "kernel_ms": sum(a.elapsed_time(b) for a, b in pairs), "substeps": int(sub.item()),
                "nograb": int(nog.item()), "done": int(dn.item()), "launches": int(c.L.clothb200_launch_count() - L0), "clocks": clocks,
                "busy_cycles": (c.cost.double() * c.sim_steps.double()).clone()}

    # -- `e2e`: host buffers in, host buffers out
    def host_buffers(self):
        torch, n, N = self.torch, self.n, self.np_
        pin = lambda *s, dt: torch.zeros(*s, dtype=dt).pin_memory()
        t = {"obs": pin(n, 3 * N, dt=self.tdt), "reward": pin(n, dt=torch.float64), "done": pin(n, dt=torch.int32),
             "coverage": pin(n, dt=torch.float64), "variance_inv": pin(n, dt=torch.float64), "flags": pin(n, dt=torch.int32),
             "sim_steps": pin(n, dt=torch.int32)}
        self._pinned = t
        return {k: v.numpy() for k, v in t.items()}

    def run_host(self, K, W, barrier, t_base=1000):
        torch, c = self.torch, self.c
        host = self.host_buffers()
        pool_obs = self.pool["pos"][:, :, :3].reshape(self.n_pool, -1).cpu().numpy()
        host["obs"][:] = c.pos[:, :, :3].reshape(self.n, -1).cpu().numpy()
        drawn = [draw_actions(self.args.seed, t_base + t, self.lo, self.hi, self.np_) for t in range(W + K)]
        ar = np.arange(self.n)
        t0 = None
        for t in range(W + K):
            if t == W:
                barrier()
                t0 = time.perf_counter()
            raw, pick = drawn[t]
            if self.aim:
                a = raw.copy()
                xy = host["obs"].reshape(self.n, self.np_, 3)[ar, pick, :2].astype(np.float64)
                a[:, :2] = (xy - 0.5) * 2
            else:
                a = raw
            c.step_host(a, host)
            _ = float(host["reward"][0])                      # the caller reads its result
            d = np.nonzero(host["done"])[0]
            if len(d):
                k = self.restart(t_base + t, torch.from_numpy(d).to(c.device))
                host["obs"][d] = pool_obs[k.cpu().numpy()]
        torch.cuda.synchronize()
        s = time.perf_counter() - t0
        barrier()
        import ctypes as C
        from gym_cloth_b200 import lib as L
        esz = 4 if self.dtype == "f32" else 8
        return {"seconds": s, "h2d": self.n * C.sizeof(L.Plan), "d2h": self.n * (3 * self.np_ * esz + 8 + 4 + 8 + 8 + 4 + 4)}

    def measure_pairs(self, t):
        """Ordered self-collision pair tests per substep (SURVEY.md §8(d)'s P) of one more, untimed, profiled step."""
        import ctypes as C
        torch, c = self.torch, self.c
        prof = torch.zeros(self.n, 16, dtype=torch.int64, device=c.device)
        c.L.clothb200_debug_set_profile(C.c_void_p(prof.data_ptr()))
        self.device_step(t, self.device_actions(t)); torch.cuda.synchronize()
        c.L.clothb200_debug_set_profile(None)
        sub = float(c.sim_steps.sum().item())
        d = torch.nonzero(c.done)[:, 0]
        if d.numel():
            self.restart(t, d)
        return float(prof[:, 15].sum().item()) / max(sub, 1.0)


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from gym_cloth_b200 import dist as D, lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()
    wl, scaling, n, gridw = workload(args, world)
    dtype = args.dtype
    coloured = args.config == 4 or args.mode == "coloured"
    mode = L.MODE_COLOURED if coloured else L.MODE_REFERENCE_ORDER
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    arm = Arm(args, rank, n, dtype, mode)
    c = arm.c
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    r = arm.run_device(K, W, barrier, flush=flush, sampler=ClockSampler(local))
    clocks = r["clocks"]
    h = arm.run_host(K, min(W, 1), barrier)
    pairs = arm.measure_pairs(5000) if not (coloured or args.no_pairs) else None

    # ---------------- max over ranks of the times, sums of the counts (gym_cloth_b200/dist.py) ----------------
    elapsed_ms, e2e_ms, kernel_ms = D.reduce_max([r["elapsed_ms"], h["seconds"] * 1e3, r["kernel_ms"]], dev)
    total_substeps, total_nograb, total_done = D.reduce_sum([r["substeps"], r["nograb"], r["done"]], dev)
    stats = D.episode_stats(c.coverage, c.done, dev)      # the optional episode-statistics gather (NCCL)

    if rank == 0:
        n_total = n * world
        value = n_total * K / (elapsed_ms * 1e-3)
        sub_per_s = total_substeps / (elapsed_ms * 1e-3)
        peaks, peak_src = measured_peaks()
        smem_peak = C.c_double(0); fp32_peak = C.c_double(0)
        lib.clothb200_bench_smem_bandwidth(2000, C.byref(smem_peak), None)
        lib.clothb200_bench_fp32_flops(2000, C.byref(fp32_peak), None)
        per_gpu_sub_per_s = (total_substeps / world) / (kernel_ms * 1e-3)
        p_alg = P_FLAT if gridw == 25 else 0
        b_sub = smem_bytes_per_substep(gridw, gridw, p_alg)
        smem_ach = per_gpu_sub_per_s * b_sub / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "step_kernel_dram_bytes_per_env_step.json")
        if os.path.exists(tp) and gridw == 25:
            try:
                traffic = json.load(open(tp))["bytes_per_env_step"] * n
            except Exception:
                traffic = None
        ctas = C.c_int(0); smemb = C.c_int(0); thr = C.c_int(0)
        lib.clothb200_occupancy(C.byref(c.P), int(dtype == "f64"), C.byref(ctas), C.byref(smemb), C.byref(thr))
        esz = 4 if dtype == "f32" else 8
        hbm_b_env_step = 2 * 2 * c.N * 4 * esz + 3 * c.N * esz       # pos+prev load+store + obs
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": config_dict(args, wl, n, n_total, gridw, coloured),
            "substeps_per_s": sub_per_s, "substeps_per_env_step": total_substeps / (n_total * K),
            "nograb_frac": total_nograb / (n_total * K), "done_frac": total_done / (n_total * K),
            "e2e": {"value": n_total * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h["h2d"] * world,
                    "d2h_bytes_per_step": h["d2h"] * world, "ms_per_step": e2e_ms / K},
            "gpu_launches": r["launches"],
            "clocks": clocks,
            "run": {"threads_per_cloth": thr.value, "smem_bytes_per_cloth": smemb.value, "resident_cloths_per_sm": ctas.value,
                    "reset_seconds": arm.reset_s,
                    "l2": "per-step working set = %d MiB of state read and written once per launch (cloths live in shared memory "
                          "while they are worked on; HBM/L2 see one load and one store per env-step plus one per swap of the "
                          "time-sliced launch)" % (n * c.N * 8 * esz // 2 ** 20)},
            "roofline": {"bound": "smem", "achieved": smem_ach, "peak": smem_peak.value, "unit": "GB/s",
                         "frac": smem_ach / smem_peak.value if smem_peak.value else None, "traffic": traffic,
                         "traffic_note": "DRAM bytes per launch from the ncu capture in profiles/ (x envs): 40 KB per env-step are the one "
                                         "load and store of the state, the rest are the 40 KB swaps of the time-sliced launch (DESIGN.md 4); "
                                         "HBM runs at < 0.1 % of its peak either way",
                         "kernel": "cloth_step_kernel", "kernel_ms_per_launch": kernel_ms / K,
                         "algorithmic_bytes_per_substep": b_sub, "substeps_per_launch": total_substeps / world / K,
                         "peak_source": "LDS.128 streaming microbenchmark run in this process (clothb200_bench_smem_bandwidth; its ncu "
                                        "wavefront rate is in profiles/); MEASURED_PEAKS.json has no shared-memory figure",
                         "note": "BASELINE.json's metric names the shared-memory/FP32 roofline for this path; HBM is not the bound. "
                                 "algorithmic bytes use SURVEY.md 8(d)'s flat-cloth pair count P=%d; see pairs_measured" % p_alg},
            "roofline_fp32": {"achieved": per_gpu_sub_per_s * flop_per_substep(gridw, gridw, p_alg) / 1e12, "peak": fp32_peak.value,
                              "unit": "TFLOP/s",
                              "frac": per_gpu_sub_per_s * flop_per_substep(gridw, gridw, p_alg) / 1e12 / fp32_peak.value if fp32_peak.value else None,
                              "peak_source": "FFMA microbenchmark run in this process"},
            "roofline_hbm": {"achieved": (n * K / (kernel_ms * 1e-3)) * hbm_b_env_step / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": (n * K / (kernel_ms * 1e-3)) * hbm_b_env_step / 1e9 / peaks["hbm_gbs"], "peak_source": peak_src},
            "episode_stats": stats,
        }
        if pairs is not None:
            bm = smem_bytes_per_substep(gridw, gridw, pairs)
            line["roofline"]["pairs_measured"] = {"ordered_pair_tests_per_substep": pairs, "algorithmic_bytes_per_substep": bm,
                                                  "achieved": per_gpu_sub_per_s * bm / 1e9,
                                                  "frac": per_gpu_sub_per_s * bm / 1e9 / smem_peak.value if smem_peak.value else None,
                                                  "note": "P counted by the kernel on one more (untimed) step of the same workload"}
        # why a launch takes as long as it does: it cannot end before its slowest cloth nor before total work / resident slots
        mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        slots = ctas.value * torch.cuda.get_device_properties(dev).multi_processor_count
        bc = r["busy_cycles"]
        line["launch_balance"] = {"slowest_cloth_ms": float(bc.max().item()) / (mhz * 1e3),
                                  "work_per_slot_ms": float(bc.sum().item()) / (mhz * 1e3) / slots,
                                  "active_cloths": int((bc > 0).sum().item()), "resident_slots": int(slots),
                                  "note": "rank 0, last timed launch; cycles spent inside the substep loop only"}
        del arm, c
        if world == 1 and args.config == 2 and not args.no_extras and dtype == "f32" and not coloured:
            for key, dt2, md in (("parity_build", "f64", L.MODE_REFERENCE_ORDER), ("coloured_build", "f32", L.MODE_COLOURED)):
                a2 = Arm(args, rank, n, dt2, md)
                r2 = a2.run_device(K, W, barrier, flush=flush)
                line[key] = {"build": "reference_order_f64 (bit-exact against the reference fixtures)" if key == "parity_build"
                             else "graph_coloured_f32", "value": n * K / (r2["elapsed_ms"] * 1e-3), "unit": UNIT, "steps": K, "warmup": W,
                             "ms_per_step": r2["elapsed_ms"] / K, "substeps_per_s": r2["substeps"] / (r2["elapsed_ms"] * 1e-3),
                             "substeps_per_env_step": r2["substeps"] / (n * K), "nograb_frac": r2["nograb"] / (n * K),
                             "mean_coverage": float(a2.c.coverage.mean().item())}
                del a2
        if world == 1 and args.config == 2 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
            line["substeps_ratio_vs_cpu_baseline"] = sub_per_s / line["cpu_baseline"]["substeps_per_s"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def config_dict(args, wl, n, n_total, gridw, coloured):
    """The part of the line both arms print identically (the driver compares it)."""
    return {"workload": wl, "envs_per_gpu": n, "envs_total": n_total, "grid": "%dx%d" % (gridw, gridw),
            "mode": "coloured" if coloured else "reference_order", "actions": args.actions,
            "start_states": "tier-1 reset states of env seeds seed+i; an env that ends its episode restarts from that pool" if args.config in (2, 5)
            else ("tier-2 / tier-3 reset states; restarts from that pool" if args.config == 3 else "flat cloth; restarts flat"),
            "l2_flush": "256 MiB buffer written between timed iterations"}


# ---------------------------------------------------------------------------------------------------------------- CPU legs
def _ref_kind():
    from oracle.build_ref import ref_built
    return "reference" if ref_built() else "port"


def _ref_pool():
    d = np.load(POOL_FIXTURE)
    return d["pos"], d["prev"]


def _ref_episodes(args, W, K, cores):
    """`cores` environments (global ids 0..cores-1) of the configs[1] workload on the host: the same start states, action
    draws, aiming rule and restart rule as the GPU arm's environments with those ids."""
    from oracle.ref_driver import cpu_env_episodes
    pos, prev = _ref_pool()
    T = W + K
    raw = np.stack([draw_actions(args.seed, t, 0, cores)[0] for t in range(T)])
    pick = np.stack([draw_actions(args.seed, t, 0, cores)[1] for t in range(T)]) if args.actions == "touch_cloth" else None
    choice = np.stack([restart_choice(args.seed, t, 0, cores, len(pos)) for t in range(T)])     # lo = 0, hi = cores
    starts = [i % len(pos) for i in range(cores)]
    return cpu_env_episodes(_ref_kind(), pos, prev, starts, raw, pick, choice, W, cores), (pos, prev, raw, pick)


def cpu_baseline(args):
    """The reference's CPU path beside the GPU number: ONE env.step on every host core (bounded sample, ~10 s), same
    start states and actions as GPU environments 0..cores-1 at step 0 - and, because they are the same, a parity check of
    the f64 GPU build against what the reference just computed."""
    import torch
    from gym_cloth_b200 import lib as L
    from gym_cloth_b200.batched import BatchedCloth
    cores = os.cpu_count() or 1
    r, (pos, prev, raw, pick) = _ref_episodes(args, 0, 1, cores)
    out = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": _ref_kind(),
           "sample": "%d env.step calls, one per host process (%d substeps, %.1f s): step 0 of environments 0..%d of this workload from "
                     "the reference-recorded reset states (tests/golden/bench_pool_t1.npz)" % (r["n"], r["substeps"], r["seconds"], cores - 1),
           "substeps_per_s": r["substeps_per_s"], "substeps_per_s_per_core": r["substeps_per_s"] / r["cores"],
           "core_busy_frac": r["core_busy_frac"], "nograb_frac": r["nograb_frac"]}
    try:
        bc = BatchedCloth(L.default_params(), cores, dtype=torch.float64)
        idx = [i % len(pos) for i in range(cores)]
        for i in range(cores):
            bc.set_state(pos[idx[i]], prev[idx[i]], env=i)
        a = raw[0].copy()
        if pick is not None:
            for i in range(cores):
                a[i, :2] = (pos[idx[i]][pick[0][i], :2] - 0.5) * 2
        host = {"coverage": np.zeros(cores), "sim_steps": np.zeros(cores, np.int32)}
        bc.step_host(a, host)
        dpos = max(float(np.abs(bc.get_state(i)[0] - r["final_pos"][i]).max()) for i in range(cores))
        out["parity_check_f64_gpu_vs_this_run"] = {"max_abs_dpos": dpos, "substeps_equal": bool(
            [int(x) for x in host["sim_steps"]] == [e[0] for e in r["n_updates"]]),
            "max_abs_dcoverage": float(np.abs(host["coverage"] - np.array([e[0] for e in r["coverage"]])).max())}
    except Exception as e:                                   # a diagnostic, never a reason to lose the line
        out["parity_check_f64_gpu_vs_this_run"] = {"error": repr(e)}
    return out


def run_reference(args):
    """--impl reference: the reference's own compiled physics (oracle/_ref; the C port when it is not built) on all host
    cores, one environment per core, the configs[1] workload of the GPU arm (see the module docstring)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    if args.config != 2:
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm runs BASELINE configs[1] only (--config 2)"}))
        return
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    world = int(os.environ.get("WORLD_SIZE", 1))
    r, _ = _ref_episodes(args, W, K, cores)
    wl, scaling, n, gridw = workload(args, world)
    value = r["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": r["seconds"] / K * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, wl, n, n * world, gridw, False),
            "substeps_per_s": r["substeps_per_s"], "substeps_per_env_step": r["substeps"] / r["n"],
            "nograb_frac": r["nograb_frac"], "done_frac": r["done_frac"],
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": _ref_kind(),
                             "sample": "environments 0..%d of the workload, one per host process: %d untimed + %d timed env.step calls each, "
                                       "%d substeps, %.1f s, cores busy %.0f %% of the timed region" % (
                                           cores - 1, W, K, r["substeps"], r["seconds"], 100 * r["core_busy_frac"]),
                             "substeps_per_s": r["substeps_per_s"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json configs, 1-based: 2 (default, the metric's), 3, 4, 5")
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (config 5: in total); 0 = the config's own")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--mode", default="reference_order", choices=["reference_order", "coloured"])
    ap.add_argument("--actions", default="touch_cloth", choices=["touch_cloth", "over_xy_plane"])
    ap.add_argument("--relax-iters", type=int, default=2, help="config 4: limit passes per substep")
    ap.add_argument("--seed", type=int, default=1337)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-pairs", action="store_true", help="skip the extra (untimed, instrumented) step that counts pair tests")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
