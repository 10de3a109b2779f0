#!/usr/bin/env python
"""bench.py -- converged NMPC solves/s of the batched closed loop (BASELINE.json metric).

Default workload (BASELINE.json configs[1], --config 2): Python/NMPC_TT.py's NLP (T=1, N=15, 3 obstacles) batched over
4096 randomised UAV initial states and target speeds PER GPU, driven by the reference's shift-and-apply-first-input
closed loop.  --config 3 / 4 / 5 are the other BASELINE configurations at their per-GPU batch sizes (T / Plus
trajectory schedules at 65536; 10_obstacles.py with per-instance obstacle layouts at 32768; Race Track 2 at N = 30 with
per-instance obstacles at 131072).  A "step" is one closed-loop batch step: one NLP solve per instance + plant /
target / warm-start shift.  Warm-up steps include the atypical cold first solve (all-zero warm start, NMPC_TT.py:329).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]      this repo's CUDA path
  python bench.py --impl reference ...                                   CPU restatement of the reference path (oracle/)
Under torchrun every rank drives its own GPU on its own instances (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "converged_nmpc_solves_per_sec"
UNIT = "solves/s"

# BASELINE.json configs[1..4] ("config 2..5" in SURVEY.md section 8d), per GPU.  cpu = (instances, consecutive closed-loop
# steps) of the bounded CPU sample timed beside the GPU number (about 10-30 s of oracle work on 16 threads).
CONFIGS = {
    2: dict(scenario="nmpc_tt", N=15, batch=4096, jitter=0, mix=None, cpu=(4096, 12), ref=(1024, None),
            what="Python/NMPC_TT.py batched over randomised UAV states and target speeds (BASELINE.json configs[1])"),
    3: dict(scenario="t_trajectory", N=15, batch=65536, jitter=0, mix="plus_trajectory", cpu=(16384, 3), ref=(2048, None),
            what="T_Trajectory / Plus Trajectory target paths, half the batch on each schedule, random schedule phase "
                 "(BASELINE.json configs[2])"),
    4: dict(scenario="10_obstacles", N=15, batch=32768, jitter=3, mix=None, cpu=(8192, 4), ref=(1024, None),
            what="10_obstacles.py (15 g rows per stage), per-instance obstacle layouts, 262144 instances over 8 GPUs "
                 "(BASELINE.json configs[3])"),
    5: dict(scenario="race_track_2", N=30, batch=131072, jitter=10, mix=None, cpu=(2048, 3), ref=(256, None),
            what="Race Track 2 at twice the reference horizon (N = 30), per-instance obstacle layouts, 1M instances over "
                 "8 GPUs (BASELINE.json configs[4])"),
}


def make_workload(b200nmpc, cfg, B, seed):
    """Synthetic instances of one BASELINE config: p [B,11], constant target (v, omega) [B,2] or
    None when the scripts' schedules drive the target, per-instance obstacle tables [B,n_obs,3] or None, and the
    schedule arguments (functions, schedule index per instance, phase per instance)."""
    sc = b200nmpc.SCENARIOS[cfg["scenario"]]
    if cfg["N"] != sc.N:
        sc = sc.with_horizon(cfg["N"])
    p, vw = b200nmpc.random_instances(sc, B, seed=seed)
    rng = np.random.default_rng(seed + 7)
    obs = None
    if cfg["jitter"]:
        j = cfg["jitter"]
        obs = np.tile(sc.obstacle_table(), (B, 1, 1))
        obs[:, :j, :2] += rng.uniform(-100, 100, (B, j, 2))
        d = np.linalg.norm(obs[:, :, :2] - p[:, None, :2], axis=2)
        obs[:, :, 0] += np.where(d < obs[:, :, 2] + 20.0, 400.0, 0.0)       # no obstacle within r + 20 of the start
    sched = None
    if cfg["mix"]:
        other = b200nmpc.SCENARIOS[cfg["mix"]]
        which = (np.arange(B) % 2).astype(np.int64)                        # half the batch on each schedule
        phase = np.where(which == 0, rng.integers(0, sc.steps, B), rng.integers(0, other.steps, B)).astype(np.int64)
        sched = dict(schedules=[sc.schedule, other.schedule], schedule_of=which, phase=phase)
        vw = None
    return sc, p, vw, obs, sched


def host_vw(sched, sc_steps, it):
    """(v, omega) [B,2] of closed-loop step `it` for the host-driven arms (the device arm looks it up in-kernel)."""
    fns, which, phase = sched["schedules"], sched["schedule_of"], sched["phase"]
    out = np.empty((len(which), 2))
    for r, fn in enumerate(fns):
        m = which == r
        at = phase[m] + it
        u, inv = np.unique(at, return_inverse=True)
        vals = np.array([fn(int(a)) for a in u])
        out[m] = vals[inv]
    return out


def infeasible_by_construction(sc, p, obs, lbg, ubg, relax=1e-8, tol=1e-8):
    """Sufficient (conservative) certificate that an NLP instance has no feasible point, from p alone:
      * a stage-0 row of g (a function of p only: z, theta, X5, X6, X7 of the current state, obstacle distances)
        violates its relaxed bound by more than IPOPT's tolerance, or
      * the altitude rows z_k <= 150 / z_k >= 75 cannot be met with full control authority (pitch rate at its limit,
        speed at whichever bound helps): z_k,min = z_0 + T sum_j min_v v sin(theta_j,min) still exceeds the bound.
    Returns a boolean mask [B]."""
    R = 5 + sc.n_obs
    lo = lbg[:R] - relax * np.maximum(1.0, np.abs(np.where(np.isfinite(lbg[:R]), lbg[:R], 0.0))) - tol
    hi = ubg[:R] + relax * np.maximum(1.0, np.abs(np.where(np.isfinite(ubg[:R]), ubg[:R], 0.0))) + tol
    g0 = np.empty((p.shape[0], R))
    g0[:, :5] = p[:, [2, 3, 5, 6, 7]]
    ob = obs if obs is not None else np.tile(sc.obstacle_table(), (p.shape[0], 1, 1))
    g0[:, 5:] = ob[:, :, 2] - np.linalg.norm(p[:, None, :2] - ob[:, :, :2], axis=2)
    bad = ((g0 < lo) | (g0 > hi)).any(axis=1)
    T, wmax, thb = sc.T, np.pi / 30 * (1 + relax), 0.2618 * (1 + relax)
    vlo, vhi = 14.0 * (1 - relax), 30.0 * (1 + relax)
    zmin = p[:, 2].copy(); zmax = p[:, 2].copy(); th_dn = p[:, 3].copy(); th_up = p[:, 3].copy()
    for _ in range(sc.N):
        zmin += T * np.minimum(vlo * np.sin(th_dn), vhi * np.sin(th_dn)); th_dn = np.maximum(th_dn - T * wmax, -thb)
        zmax += T * np.maximum(vlo * np.sin(th_up), vhi * np.sin(th_up)); th_up = np.minimum(th_up + T * wmax, thb)
        bad |= (zmin > hi[0]) | (zmax < lo[0])
    return bad


def host_shift(T, p, x, vw):
    """Vectorised shift_timestep (NMPC_TT.py:13-30) on host arrays: p [B,11] in place, returns the warm start."""
    th, ps, v = p[:, 3].copy(), p[:, 4].copy(), x[:, 0]
    p[:, 0] += T * v * np.cos(ps) * np.cos(th)
    p[:, 1] += T * v * np.sin(ps) * np.cos(th)
    p[:, 2] += T * v * np.sin(th)
    p[:, 3:8] += T * x[:, 1:6]
    tt = p[:, 10].copy()
    p[:, 8] += T * vw[:, 0] * np.cos(tt)
    p[:, 9] += T * vw[:, 0] * np.sin(tt)
    p[:, 10] += T * vw[:, 1]
    return np.concatenate([x[:, 6:], x[:, -6:]], axis=1)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):      # (debug knob)
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """Wall-clock marker: call at the start and at the end of the timed region."""
        self.marks = getattr(self, "marks", []) + [time.time()]

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        marks = getattr(self, "marks", [])
        rows, window = [r for _, r in self.rows], "whole run (the device is busy from the first warm-up step on)"
        if len(marks) >= 2:     # samples taken during the timed region; nvidia-smi's period is 100 ms, so keep a margin
            inside = [r for t, r in self.rows if marks[0] - 0.05 <= t <= marks[-1] + 0.05]
            if inside:
                rows, window = inside, "timed region"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def algorithmic_flops(N, n_obs, iters, n_fact, n_ls):
    """SURVEY.md section 8d planning formula (add/mul = 1, FMA = 2): Riccati factorisation N*3379, Riccati solve N*968,
    derivative evaluation + adjoint + residuals N*(560+30*n_obs), one line-search trial N*(60+8*n_obs)."""
    return N * (3379.0 * n_fact + 968.0 * iters + (560.0 + 30.0 * n_obs) * iters + (60.0 + 8.0 * n_obs) * n_ls)


def bytes_per_solve(N, n_obs, per_instance_obs=False):
    """Algorithmic HBM bytes per solve: read p + warm start, write x + f, status + iters."""
    return 8 * (11 + 6 * N) + 8 * (6 * N + 1) + 8 + (24 * n_obs if per_instance_obs else 0)


# --------------------------------------------------------------------------------------------------
def oracle_closed_loop(oracle, sc, p, vw, obs, sched, u, steps, cores, it0=0):
    """`steps` closed-loop steps of the CPU oracle on host arrays (p, u updated in place / returned).  Returns
    (converged, iterations, per-step seconds, status histogram, u)."""
    sp = oracle.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    lbx, ubx, lbg, ubg = sc.bounds()
    ob = sc.obstacle_table() if obs is None else obs
    conv = its = 0; secs = []; hist = np.zeros(8, dtype=np.int64)
    for k in range(steps):
        t0 = time.perf_counter()
        r = oracle.solve(sp, ob, p, u, lbx, ubx, lbg, ubg, obs_per_instance=obs is not None, nthreads=cores, want_g=False, want_lam=False)
        v = vw if sched is None else host_vw(sched, sc.steps, it0 + k)
        u = host_shift(sc.T, p, r["x"], v)
        secs.append(time.perf_counter() - t0)
        conv += int((r["status"] == 0).sum()); its += int(r["iters"].sum()); hist += np.bincount(r["status"], minlength=8)[:8]
    return conv, its, secs, hist, u


def sub_sched(sched, idx):
    return None if sched is None else dict(schedules=sched["schedules"], schedule_of=sched["schedule_of"][idx], phase=sched["phase"][idx])


def run_reference(args, rank, world):
    """CPU arm: the oracle (CPU restatement of the reference's CasADi/IPOPT path -- the reference itself cannot run
    here: casadi is not installable) on all host cores, same closed loop, bounded sample."""
    if rank != 0:
        return
    import b200nmpc
    import oracle
    oracle.build()
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    B = args.ref_batch or cfg["ref"][0]
    sc, p, vw, obs, sched = make_workload(b200nmpc, cfg, B, seed=2000)
    u = np.zeros((B, sc.n_w))
    _, _, _, _, u = oracle_closed_loop(oracle, sc, p, vw, obs, sched, u, args.warmup, cores, 0)
    conv, its, secs, hist, u = oracle_closed_loop(oracle, sc, p, vw, obs, sched, u, args.steps, cores, args.warmup)
    total = sum(secs)
    val = conv / total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config {args.config}: {cfg['what']}; {sc.script} NLP (T={sc.T}, N={sc.N}, n_obs={sc.n_obs}) closed loop",
                       "batch": B, "note": "bounded CPU sample of the GPU arm's workload"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{B} instances x {args.steps} closed-loop steps, oracle/nmpc_oracle.cpp on {cores} threads "
                                       "(CasADi/IPOPT itself is unavailable in this image)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "converged_fraction": conv / (B * args.steps), "mean_iters": its / (B * args.steps), "status_hist": hist.tolist()}
    emit(line)


# --------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import b200nmpc
    from mpc_implementation_b200 import sharding
    from mpc_implementation_b200.closed_loop import PipelinedClosedLoop

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    if args.no_lpt:
        os.environ["NMPC_B200_AUTO_ORDER"] = "0"
    # sub-batches: measured best 8 at B = 4096, 2 at B = 16384 (tools/pipeline_probe.py), and still 2 for the big batches (tools/
    # pack_sweep_big.sh: configs 3 / 4 +6.5 % / +10 % over one batch): about 32768 / B, at least 2, at most 8
    S = max(1, min(args.pipelines, B)) if args.pipelines > 0 else max(1, min(B, max(2, min(8, 32768 // max(B, 1)))))     # PipelinedClosedLoop default
    sc, p, vw, obs, sched = make_workload(b200nmpc, cfg, B, seed=2000 + rank)
    per_obs = obs is not None
    # fill = 2: a sub-batch occupies half as many SMs as it has warps' worth of instances, leaving SMs to the other sub-batches
    mk = lambda n: b200nmpc.nlpsol("solver", "ipm", sc, {"ipopt": {"max_iter": 100}}, device=local_rank, max_batch=n,
                                   fill=2 if S > 1 else 1)
    # the batch advances as S independently pipelined sub-batches (own handle + stream each): one sub-batch's stragglers
    # overlap with the next one's bulk; per-instance results are those of the single-batch loop (tests/test_gpu_parity.py)
    cl = PipelinedClosedLoop(mk, sc, p, target_vw=vw, pipelines=S, obstacles=obs,
                             **({} if sched is None else dict(schedules=sched["schedules"], schedule_of=sched["schedule_of"], phase=sched["phase"])))
    flush = torch.empty(144 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)    # 151 MB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank); sampler.start()     # started early: nvidia-smi needs ~0.3 s to deliver its first line
    cold = None
    for k in range(args.warmup):
        cl.step()
        if k == 0:
            st = cl.stats()
            cold = {"converged_fraction": float(st["success"].double().mean()), "mean_iters": float(st["iter_count"].double().mean())}
    barrier()
    K = args.steps
    ev = [[[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(S)] for _ in range(K)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    keep = []
    barrier()
    sampler.mark()
    wall0 = time.perf_counter()
    cur = torch.cuda.current_stream(dev)
    e_start.record(cur)
    for st_ in cl.streams:
        st_.wait_event(e_start)
    for k in range(K):
        for i, (lp, st_) in enumerate(zip(cl.loops, cl.streams)):
            with torch.cuda.stream(st_):
                if not os.environ.get("BENCH_NO_FLUSH"):    # (debug knob; reported numbers always flush)
                    flush.zero_()                       # L2 flush before every sub-batch step (inputs << L2 except config 5)
                ev[k][i][0].record()
                # ONE launch: solve + shift_timestep + FOV error + the target schedule lookup (nmpc_solve_and_step)
                lp.solver.solve_and_step(lp.p, lp.u_warm, lp.lbx, lp.ubx, lp.lbg, lp.ubg, lp.vw if lp._const_vw else None,
                                         lp.fov, lp.err_sum, obstacles=lp.obstacles)
                ev[k][i][1].record()
                sst = lp.solver._stats               # status / iteration arrays of this step: counted after the timed region
                keep.append((sst["return_status"], sst["iter_count"]))
                lp.mpc_iter += 1
    cl.join(cur)
    e_end.record(cur)
    barrier()
    wall = time.perf_counter() - wall0
    sampler.mark()
    clocks = sampler.stop()
    solve_ms = [ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(K) for i in range(S)]      # per sub-batch launch
    # whole-batch step latency: step k of the batch is complete when its last sub-batch is; latency = spacing of completions
    done_at = [max(e_start.elapsed_time(ev[k][i][1]) for i in range(S)) for k in range(K)]
    batch_step_ms = np.diff(np.array([0.0] + done_at))
    t_rank = e_start.elapsed_time(e_end) * 1e-3
    t_max = sharding.max_over_ranks(t_rank, dev)
    conv_loc = sum(int((a == 0).sum().item()) for a, _ in keep); it_loc = sum(int(b_.sum().item()) for _, b_ in keep)
    hist_loc = torch.zeros(8, dtype=torch.float64, device=dev)
    for a, _ in keep:
        hist_loc += torch.bincount(a.to(torch.int64), minlength=8)[:8].double()
    tot = sharding.sum_counters([conv_loc, it_loc] + hist_loc.tolist(), dev).cpu().numpy()
    conv_all, iters_all, status_hist = float(tot[0]), float(tot[1]), [int(v) for v in tot[2:10]]
    value = conv_all / t_max
    # per-step series (rank 0's own instances): converged fraction and launch time, first / last quarter of the timed steps
    conv_series = [float(np.mean([float((keep[k * S + i][0] == 0).double().mean()) for i in range(S)])) for k in range(K)]

    # ---- result records gathered over NCCL once, after the timed region (the only collective of the path)
    gathered = None
    if world > 1:
        last = cl.stats()
        rec = torch.stack([last["return_status"].double(), last["iter_count"].double(), cl.err_sum], dim=1)
        allrec = sharding.gather_rows(rec, B * world, world)
        gathered = {"records": int(allrec.shape[0]), "converged_last_step": int((allrec[:, 0] == 0).sum().item())}

    # ---- e2e: the SAME closed loop (same instances, same warm-up, same timed steps, same sub-batches) driven through the
    #      host-buffer entry point: every step copies p and the warm start (and the obstacle tables) H2D from pinned
    #      memory, solves, copies x, f, status, iters D2H, and does the shift on the host like the reference script does
    #      (NMPC_TT.py:382).  The sub-batches are issued with nmpc_solve_host_async and completed in turn, so that the
    #      host-side shift of one overlaps the device work of the others.
    Ke = min(K, args.e2e_steps) if args.e2e_steps > 0 else K
    lbx, ubx, lbg, ubg = sc.bounds()
    pin = lambda shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    # sub-batches of the host arm: the device arm's when it has several; a big single batch is split in four so that the copies
    # and the (single-threaded numpy) shift of one quarter overlap the solves of the others -- what a host-side caller would do
    S_e = S if B < 16384 else max(S, 4)
    index_e = cl.index if S_e == S else np.array_split(np.arange(B), S_e)
    hs = []
    for idx in index_e:
        ph = pin((len(idx), 11)); ph[:] = p[idx]
        uh = pin((len(idx), sc.n_w)); uh[:] = 0.0
        oh = None
        if per_obs:
            oh = pin((len(idx), sc.n_obs, 3)); oh[:] = obs[idx]
        hs.append(dict(p=ph, u=uh, obs=oh, vw=None if vw is None else vw[idx].copy(), sched=sub_sched(sched, idx), it=0, solver=mk(len(idx))))
    issue = lambda h: h["solver"](x0=h["u"], p=h["p"], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, obstacles=h["obs"], want_g=False, want_lam=False, blocking=False)
    def host_steps(n, count):
        """n closed-loop steps of every sub-batch, software-pipelined and completion-ordered: whichever sub-batch has
        finished its solve gets its host-side shift and its next solve issued first.  Returns (converged solves, wall
        time at which the k-th step of ALL sub-batches was complete)."""
        conv, marks = 0, []
        pend = [issue(h) for h in hs]
        left = [n] * len(hs)
        t_mark = time.perf_counter()
        active = set(range(len(hs)))
        while active:
            for i in list(active):
                h = hs[i]
                if not h["solver"].done():
                    continue
                h["solver"].wait()
                if count:
                    conv += int(h["solver"].stats()["success"].sum())
                v = h["vw"] if h["sched"] is None else host_vw(h["sched"], sc.steps, h["it"])
                h["u"][:] = host_shift(sc.T, h["p"], pend[i]["x"], v)
                h["it"] += 1
                left[i] -= 1
                if left[i] > 0:
                    pend[i] = issue(h)
                else:
                    active.discard(i)
                done_steps = n - max(left)
                while len(marks) < done_steps:
                    now = time.perf_counter(); marks.append(now - t_mark); t_mark = now
        return conv, marks

    barrier()
    host_steps(args.warmup, False)          # untimed; the pipeline is drained when it returns
    barrier()
    conv_e, times = host_steps(Ke, True)
    t_e = sum(times); e2e_ms = [round(t * 1e3, 3) for t in times]
    barrier()
    t_e_max = sharding.max_over_ranks(t_e, dev)
    conv_e_all = float(sharding.sum_counters([conv_e], dev)[0])
    e2e_val = conv_e_all / t_e_max if t_e_max > 0 else None
    h2d = B * (11 + sc.n_w + (3 * sc.n_obs if per_obs else 0)) * 8 + S_e * (2 * sc.n_w + 2 * sc.n_g + (0 if per_obs else 3 * sc.n_obs)) * 8
    d2h = B * (sc.n_w + 1) * 8 + B * 8

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- the same closed loop advanced by nmpc_run_closed_loop (the scripts' whole `while mpc_iter < sim_time / T` loop on the
    #      device: a warp goes on with step k + 1 of its instance without waiting for the batch; bit-identical per instance,
    #      tests/test_gpu_parity.py::test_free_running_loop_is_the_same_loop).  Reported beside the headline, not as it: a
    #      "step" of the metric is one pass over the batch, which this mode deliberately does not have.
    free = None
    if not args.no_free_running:
        from mpc_implementation_b200.closed_loop import ClosedLoop
        scf, pf, vwf, obsf, schedf = make_workload(b200nmpc, cfg, B, seed=2000 + rank)
        clf = ClosedLoop(b200nmpc.nlpsol("free", "ipm", scf, {"ipopt": {"max_iter": 100}}, device=local_rank, max_batch=B), scf, pf,
                         target_vw=vwf, obstacles=obsf,
                         **({} if schedf is None else dict(schedules=schedf["schedules"], schedule_of=schedf["schedule_of"], phase=schedf["phase"])))
        clf.run_free(args.warmup, log=False)
        chunk = max(1, K // 4); nl = max(1, K // chunk)
        conv_f = torch.zeros((), dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(nl):
            conv_f += clf.run_free(chunk, log=False)["converged"].sum()
        f1.record(); torch.cuda.synchronize()
        ms_f = f0.elapsed_time(f1)
        free = {"value": float(conv_f.item()) / (ms_f * 1e-3), "unit": UNIT, "steps": nl * chunk, "steps_per_launch": chunk, "gpu_launches": nl,
                "ms_per_step": ms_f / (nl * chunk), "converged_fraction": float(conv_f.item()) / (B * nl * chunk), "n_gpus": 1,
                "note": "nmpc_run_closed_loop on rank 0's GPU: one handle, no sub-batch pipelining, no batch-wide barrier between steps, no L2 flush between steps (there is no step boundary to flush at)"}
        del clf

    # ---- one extra untimed step on rank 0: work counters for the flop numerator, and the census of THIS population
    #      (status histogram; how many of the non-converged NLPs are provably infeasible from p alone)
    p_now = cl.p.cpu().numpy(); obs_now = obs
    cert = infeasible_by_construction(sc, p_now, obs_now, lbg, ubg)
    it_step = 0.0; wc = {}
    for lp, st_ in zip(cl.loops, cl.streams):
        with torch.cuda.stream(st_):
            lp.solver(x0=lp.u_warm, p=lp.p, lbx=lp.lbx, ubx=lp.ubx, lbg=lp.lbg, ubg=lp.ubg, obstacles=lp.obstacles, want_g=False, want_lam=False)
    torch.cuda.synchronize()
    st_now = []
    for lp in cl.loops:
        c = lp.solver.work_counters()
        for k_, v_ in c.items():
            wc[k_] = wc.get(k_, 0) + v_
        it_step += float(lp.solver.stats()["iter_count"].sum())
        st_now.append(lp.solver.stats()["return_status"].cpu().numpy())
    st_now = np.concatenate(st_now)
    census = {"status_hist_one_step": np.bincount(st_now, minlength=8)[:8].tolist(), "not_converged": int((st_now != 0).sum()),
              "infeasible_by_construction": int(cert.sum()), "not_converged_and_certified_infeasible": int(((st_now != 0) & cert).sum()),
              "converged_although_certified": int(((st_now == 0) & cert).sum()),
              "note": "certificate = a stage-0 row of g (function of p only) beyond its relaxed bound + tol, or the altitude rows "
                      "unreachable with full control authority (bench.py: infeasible_by_construction); sufficient, not necessary"}
    flops_step = algorithmic_flops(sc.N, sc.n_obs, it_step, wc["factorizations"], wc["ls_trials"])
    k_ms = t_rank * 1e3 / K           # time of one whole-batch step (the S launches of a step overlap each other and their neighbours)
    peak64 = ctypes_fp64_peak(b200nmpc, local_rank)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); which = "of measured" if "hbm_gbs" in peaks else "of fallback"
    traffic_src = None
    traffic = None       # measured DRAM bytes of one launch of this configuration, from the committed ncu capture
    try:
        t = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(f"{sc.N}_{sc.n_obs}_{len(cl.index[0])}")
        if t:
            traffic = t["dram_read_bytes"] + t["dram_write_bytes"]; traffic_src = t["source"]
    except Exception:
        pass
    # Roofline of the dominant (only) kernel.  The S launches of a whole-batch step run concurrently, so the per-launch
    # event time is not the time the launch's work takes; the figure that follows the cross-check
    # (launches x duration = step time) is the aggregate: algorithmic bytes of a whole-batch step / its duration.
    bps = bytes_per_solve(sc.N, sc.n_obs, per_obs)
    launch_ms = float(np.mean(solve_ms))
    agg_gbs = bps * B / (k_ms * 1e-3) / 1e9
    ach_tf = flops_step / (k_ms * 1e-3) / 1e12

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle on a bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        ns, nsteps = cfg["cpu"]
        ns = min(B, args.cpu_sample or ns)
        pc = p_now[:ns].copy(); uc = cl.u_warm.cpu().numpy()[:ns].copy()
        idx = np.arange(ns)
        t0 = time.perf_counter()
        conv_c, it_c, _, hist_c, _ = oracle_closed_loop(oracle, sc, pc, None if vw is None else vw[:ns].copy(), None if obs is None else obs[:ns],
                                                        sub_sched(sched, idx), uc, nsteps, cores, it0=args.warmup + K)
        dt = time.perf_counter() - t0
        cpu = {"value": float(conv_c / dt), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{ns} instances of the GPU batch, {nsteps} consecutive closed-loop steps from the state after the timed steps "
                         f"(warm starts), oracle/nmpc_oracle.cpp on {cores} threads; CasADi/IPOPT itself cannot run in this image",
               "seconds": dt, "converged_fraction": conv_c / (ns * nsteps), "mean_iters": it_c / (ns * nsteps), "status_hist": hist_c.tolist()}

    q = max(1, K // 4)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": t_max * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config {args.config}: {cfg['what']}; {sc.script} NLP (T={sc.T}, N={sc.N}, n_obs={sc.n_obs}, n_w={sc.n_w}, "
                               f"n_g={sc.n_g}) closed loop, {B} instances per GPU",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"instances sharded over {world} GPU(s), no collective on the solve path",
                   "pipelines": S, "l2": "flushed before every sub-batch step (151 MB write > 126 MB L2)",
                   "per_instance_obstacles": per_obs, "target": "constant (v, omega) per instance" if sched is None else "scripts' schedules, looked up on the device (nmpc_set_schedule), random phase per instance",
                   "scheduling": "natural order" if args.no_lpt else "longest-first by the previous step's iteration counts (written by the previous launch's last warp)", "ipopt_options": "max_iter=100 tol=1e-8 (NMPC_TT.py:257-265)"},
        "p50_step_ms": float(np.median(batch_step_ms)), "p90_step_ms": float(np.percentile(batch_step_ms, 90)),
        "p50_note": "whole-batch step latency: spacing of the completion times of successive closed-loop steps of ALL sub-batches (CUDA events)",
        "p50_launch_ms": float(np.median(solve_ms)),
        "converged_fraction": conv_all / (B * world * K), "mean_iters": iters_all / (B * world * K),
        "status_hist": status_hist, "status_names": ["Solve_Succeeded", "Maximum_Iterations_Exceeded", "Restoration_Failed", "Search_Direction_Becomes_Too_Small",
                                                     "Invalid_Number_Detected", "Error_In_Step_Computation", "Infeasible_Problem_Detected", "-"],
        "census": census,
        "step_series": {"converged_fraction_first_quarter": float(np.mean(conv_series[:q])), "converged_fraction_last_quarter": float(np.mean(conv_series[-q:])),
                        "batch_step_ms_first_quarter": float(np.mean(batch_step_ms[:q])), "batch_step_ms_last_quarter": float(np.mean(batch_step_ms[-q:]))},
        "cold_first_step": cold, "wall_s": wall,
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke, "sub_batches": S_e, "step_ms": e2e_ms[:64],
                "api": "b200nmpc.nlpsol(...)(x0=,p=,lbx=,ubx=,lbg=,ubg=[,obstacles=], blocking=False) with pinned numpy buffers -> nmpc_solve_host_async / nmpc_query / nmpc_synchronize, one solver per sub-batch, serviced in completion order"},
        "gpu_launches": K * S,     # nmpc_ipm_kernel (solve + shift + schedule lookup + next call's fetch order), one per sub-batch step
        "gathered": gathered,
        "free_running": free,
        "roofline": {"bound": "hbm", "achieved": agg_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": agg_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_step": bps * B, "bytes_per_solve": bps,
                     "kernel": "nmpc_ipm_kernel", "launches_per_step": S, "step_ms": k_ms, "per_launch_event_ms": launch_ms, "peak_source": which,
                     "note": "aggregate over the S concurrent launches of one whole-batch step (per-launch event times overlap and do not add up); "
                             "not HBM-bound by design (SURVEY 8d): the limiter is the FP64 dependency chain, see fp64"},
        "fp64": {"achieved": ach_tf, "peak": peak64, "unit": "TFLOP/s", "frac": (ach_tf / peak64) if peak64 else None,
                 "flops_per_step": flops_step, "peak_source": "nmpc_measure_fp64_peak (DFMA loop, this GPU, this run)",
                 "iters": it_step, "work_counters_one_step": wc},
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def ctypes_fp64_peak(b200nmpc, device):
    import ctypes as C
    v = C.c_double(0.0)
    rc = b200nmpc._ffi.lib().nmpc_measure_fp64_peak(device, C.byref(v))
    return float(v.value) if rc == 0 else None


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to stdout (NCCL's "NCCL version ..." banner at init): send fd 1 to stderr for the duration of the
    # run and keep the real stdout for the JSON line, so that stdout carries exactly one line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed closed-loop batch steps (0 = per config: a timed region of >= 2 s)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (2 = configs[1], the metric's configuration)")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (0 = the config's)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the host-buffer arm (0 = as many as --steps)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--ref-batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-free-running", action="store_true", help="skip the extra nmpc_run_closed_loop measurement on rank 0")
    ap.add_argument("--pipelines", type=int, default=0, help="independently pipelined sub-batches per GPU (0 = choose from the batch size, 1 = one batch on one stream)")
    ap.add_argument("--no-lpt", action="store_true", help="disable the library's longest-first scheduling (NMPC_B200_AUTO_ORDER=0)")
    ap.add_argument("--count-work", action="store_true", help="read device work counters every timed step (adds a sync)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.steps <= 0:      # default: about 2 s of timed GPU work / a few tens of seconds of CPU work
        args.steps = ({2: 100, 3: 20, 4: 20, 5: 8} if args.impl == "b200" else {2: 20, 3: 6, 4: 4, 5: 2})[args.config]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
