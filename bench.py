#!/usr/bin/env python
"""bench.py -- converged NMPC solves/s of the batched closed loop (BASELINE.json metric).

Workload (BASELINE.json configs[1]): Python/NMPC_TT.py's NLP (T=1, N=15, 3 obstacles) batched over 4096
randomised UAV initial states and target speeds PER GPU, driven by the reference's shift-and-apply-first-input
closed loop.  A "step" is one closed-loop batch step: one NLP solve per instance + plant/target/warm-start shift.
Warm-up steps include the atypical cold first solve (all-zero warm start, NMPC_TT.py:329).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference ...                            CPU restatement of the reference path (oracle/)
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
SCENARIO = "nmpc_tt"


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
def run_reference(args, rank, world):
    """CPU arm: the oracle (CPU restatement of the reference's CasADi/IPOPT path -- the reference itself cannot run
    here: casadi is not installable) on all host cores, same closed loop, bounded sample."""
    if rank != 0:
        return
    import b200nmpc
    import oracle
    oracle.build()
    sc = b200nmpc.SCENARIOS[SCENARIO]
    cores = os.cpu_count() or 1
    B = args.ref_batch or max(64, 16 * cores)
    p, vw = b200nmpc.random_instances(sc, B, seed=2000)
    sp = oracle.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    obs = sc.obstacle_table(); lbx, ubx, lbg, ubg = sc.bounds()
    u = np.zeros((B, sc.n_w))
    conv, iters_sum, step_ms = 0, 0, []
    for k in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = oracle.solve(sp, obs, p, u, lbx, ubx, lbg, ubg, nthreads=cores, want_g=False, want_lam=False)
        u = host_shift(sc.T, p, r["x"], vw)
        dt = time.perf_counter() - t0
        if k >= args.warmup:
            step_ms.append(dt * 1e3); conv += int((r["status"] == 0).sum()); iters_sum += int(r["iters"].sum())
    total = sum(step_ms) * 1e-3
    val = conv / total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(step_ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{sc.script} NLP (T={sc.T}, N={sc.N}, n_obs={sc.n_obs}) closed loop, randomised states/targets",
                       "batch": B, "note": "bounded CPU sample of the GPU arm's workload"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{B} instances x {args.steps} closed-loop steps, oracle/nmpc_oracle.cpp on {cores} threads "
                                       "(CasADi/IPOPT itself is unavailable in this image)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "converged_fraction": conv / (B * args.steps), "mean_iters": iters_sum / (B * args.steps)}
    emit(line)


# --------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import b200nmpc
    from mpc_implementation_b200 import sharding
    from mpc_implementation_b200.closed_loop import ClosedLoop, PipelinedClosedLoop

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sc = b200nmpc.SCENARIOS[SCENARIO]
    B = args.batch
    if args.no_lpt:
        os.environ["NMPC_B200_AUTO_ORDER"] = "0"
    # sub-batches: measured best 8 at B = 4096, 2 at B = 16384 (tools/pipeline_probe.py): about 32768 / B, at most 8
    S = max(1, min(args.pipelines, B)) if args.pipelines > 0 else max(1, min(8, 32768 // max(B, 1)))     # PipelinedClosedLoop default
    p, vw = b200nmpc.random_instances(sc, B, seed=2000 + rank)
    # fill = 2: a sub-batch occupies half as many SMs as it has warps' worth of instances, leaving SMs to the other sub-batches
    mk = lambda n: b200nmpc.nlpsol("solver", "ipm", sc, {"ipopt": {"max_iter": 100}}, device=local_rank, max_batch=n,
                                   fill=2 if S > 1 else 1)
    # the batch advances as S independently pipelined sub-batches (own handle + stream each): one sub-batch's stragglers
    # overlap with the next one's bulk; per-instance results are those of the single-batch loop (tests/test_gpu_parity.py)
    cl = PipelinedClosedLoop(mk, sc, p, target_vw=vw, pipelines=S)
    solver = cl.loops[0].solver
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
    ev = [[[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(S)] for _ in range(K)]
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
                    flush.zero_()                       # L2 flush before every sub-batch step (inputs are ~6 MB << L2)
                lp._schedule_vw()
                ev[k][i][0].record()
                if args.unfused_step:
                    sol = lp.solver(x0=lp.u_warm, p=lp.p, lbx=lp.lbx, ubx=lp.ubx, lbg=lp.lbg, ubg=lp.ubg, want_g=False, want_lam=False)
                    ev[k][i][1].record()
                    lp.solver.step(sol["x"], lp.p, lp.u_warm, lp.vw, lp.fov, lp.err_sum)
                else:       # solve + shift_timestep + FOV error in one launch
                    lp.solver.solve_and_step(lp.p, lp.u_warm, lp.lbx, lp.ubx, lp.lbg, lp.ubg, lp.vw, lp.fov, lp.err_sum)
                    ev[k][i][1].record()
                ev[k][i][2].record()
                sst = lp.solver._stats               # status / iteration arrays of this step: counted after the timed region
                keep.append((sst["return_status"], sst["iter_count"]))
                lp.mpc_iter += 1
    cl.join(cur)
    e_end.record(cur)
    barrier()
    wall = time.perf_counter() - wall0
    sampler.mark()
    clocks = sampler.stop()
    step_ms = [ev[k][i][0].elapsed_time(ev[k][i][2]) for k in range(K) for i in range(S)]     # per sub-batch step
    solve_ms = [ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(K) for i in range(S)]
    fact = ls = 0
    t_rank = e_start.elapsed_time(e_end) * 1e-3
    t_max = sharding.max_over_ranks(t_rank, dev)
    conv_loc = sum(int((a == 0).sum().item()) for a, _ in keep); it_loc = sum(int(b_.sum().item()) for _, b_ in keep)
    tot = sharding.sum_counters([conv_loc, it_loc, fact, ls], dev).cpu().numpy()
    conv_all, iters_all = float(tot[0]), float(tot[1])
    value = conv_all / t_max

    # ---- e2e: the SAME closed loop (same instances, same warm-up, same timed steps, same sub-batches) driven through the
    #      host-buffer entry point: every step copies p and the warm start H2D from pinned memory, solves, copies x, f,
    #      status, iters D2H, and does the shift on the host like the reference script does (NMPC_TT.py:382).  The
    #      sub-batches are issued with nmpc_solve_host_async and completed in turn, so that the host-side shift of one
    #      overlaps the device work of the others.
    Ke = min(K, args.e2e_steps) if args.e2e_steps > 0 else K
    lbx, ubx, lbg, ubg = sc.bounds()
    pin = lambda shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    hs = []
    for idx in cl.index:
        ph = pin((len(idx), 11)); ph[:] = p[idx]
        uh = pin((len(idx), sc.n_w)); uh[:] = 0.0
        hs.append(dict(p=ph, u=uh, vw=vw[idx].copy(), solver=mk(len(idx))))
    issue = lambda h: h["solver"](x0=h["u"], p=h["p"], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, want_g=False, want_lam=False, blocking=False)
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
                h["u"][:] = host_shift(sc.T, h["p"], pend[i]["x"], h["vw"])
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
    h2d = B * (11 + sc.n_w) * 8 + S * (2 * sc.n_w + 2 * sc.n_g + 3 * sc.n_obs) * 8
    d2h = B * (sc.n_w + 1) * 8 + B * 8

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- work counters for the flop numerator (one extra untimed step on rank 0, all sub-batches)
    it_step = 0.0; wc = {"factorizations": 0, "ls_trials": 0}
    for lp, st_ in zip(cl.loops, cl.streams):
        with torch.cuda.stream(st_):
            lp.solver(x0=lp.u_warm, p=lp.p, lbx=lp.lbx, ubx=lp.ubx, lbg=lp.lbg, ubg=lp.ubg, want_g=False, want_lam=False)
    torch.cuda.synchronize()
    for lp in cl.loops:
        c = lp.solver.work_counters(); wc["factorizations"] += c["factorizations"]; wc["ls_trials"] += c["ls_trials"]
        it_step += float(lp.solver.stats()["iter_count"].sum())
    flops_step = algorithmic_flops(sc.N, sc.n_obs, it_step, wc["factorizations"], wc["ls_trials"])
    # time of one whole-batch step; with S > 1 the S launches of a step overlap each other and the neighbouring steps,
    # so the per-launch event times (p50_solve_kernel_ms) are not additive
    k_ms = t_rank * 1e3 / K
    peak64 = ctypes_fp64_peak(b200nmpc, local_rank)
    lbx, ubx, lbg, ubg = sc.bounds()
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
    # roofline.achieved, to the letter of the contract: algorithmic bytes of ONE launch / its average duration (CUDA events
    # on the launching stream; with S > 1 that duration includes the time the launch shares the GPU with its neighbours);
    # "aggregate" = the same bytes per whole-batch step / the time of a whole-batch step
    launch_ms = float(np.mean(solve_ms))
    ach_gbs = bytes_per_solve(sc.N, sc.n_obs) * len(cl.index[0]) / (launch_ms * 1e-3) / 1e9
    agg_gbs = bytes_per_solve(sc.N, sc.n_obs) * B / (k_ms * 1e-3) / 1e9
    ach_tf = flops_step / (k_ms * 1e-3) / 1e12

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle on a bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        ns = min(B, args.cpu_sample or 4096)
        pc = cl.p.cpu().numpy()[:ns].copy(); uc = cl.u_warm.cpu().numpy()[:ns].copy(); vc = vw[:ns].copy()
        sp = oracle.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
        nsteps, conv_c, it_c = 4, 0, 0
        t0 = time.perf_counter()
        for _ in range(nsteps):       # a few consecutive closed-loop steps of the same instances, ~10 s of CPU work
            r = oracle.solve(sp, sc.obstacle_table(), pc, uc, lbx, ubx, lbg, ubg, nthreads=cores, want_g=False, want_lam=False)
            conv_c += int((r["status"] == 0).sum()); it_c += int(r["iters"].sum())
            uc = host_shift(sc.T, pc, r["x"], vc)
        dt = time.perf_counter() - t0
        cpu = {"value": float(conv_c / dt), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{ns} instances of the GPU batch, {nsteps} consecutive closed-loop steps from the state after the timed steps "
                         f"(warm starts), oracle/nmpc_oracle.cpp on {cores} threads; CasADi/IPOPT itself cannot run in this image",
               "seconds": dt, "converged_fraction": conv_c / (ns * nsteps), "mean_iters": it_c / (ns * nsteps)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": t_max * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{sc.script} NLP (T={sc.T}, N={sc.N}, n_obs={sc.n_obs}, n_w={sc.n_w}, n_g={sc.n_g}) closed loop: "
                               f"{B} randomised UAV states / target speeds per GPU (BASELINE.json configs[1])",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"instances sharded over {world} GPU(s), no collective on the solve path",
                   "pipelines": S, "l2": "flushed before every sub-batch step (151 MB write > 126 MB L2)", "scheduling": "natural order" if args.no_lpt else "longest-first by the previous step's iteration counts (written by the previous launch's last warp)", "ipopt_options": "max_iter=100 tol=1e-8 (NMPC_TT.py:257-265)"},
        "p50_step_ms": float(np.median(step_ms)), "p50_solve_kernel_ms": float(np.median(solve_ms)),
        "p50_note": "per sub-batch: stream time from the start of its solve to the end of its shift / of its solve kernel",
        "converged_fraction": conv_all / (B * world * K), "mean_iters": iters_all / (B * world * K),
        "cold_first_step": cold, "wall_s": wall,
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke, "step_ms": e2e_ms,
                "api": "b200nmpc.nlpsol(...)(x0=,p=,lbx=,ubx=,lbg=,ubg=, blocking=False) with pinned numpy buffers -> nmpc_solve_host_async / nmpc_query / nmpc_synchronize, one solver per sub-batch, serviced in completion order"},
        "gpu_launches": (2 if args.unfused_step else 1) * K * S,     # nmpc_ipm_kernel (solve + shift + next call's fetch order) [, nmpc_step_kernel] per sub-batch step
        "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": bytes_per_solve(sc.N, sc.n_obs) * len(cl.index[0]),
                     "kernel": "nmpc_ipm_kernel", "kernel_ms": launch_ms, "launches_per_step": S, "peak_source": which,
                     "aggregate": {"achieved": agg_gbs, "unit": "GB/s", "step_ms": k_ms, "note": "all launches of a whole-batch step together"},
                     "bytes_per_solve": bytes_per_solve(sc.N, sc.n_obs),
                     "note": "not HBM-bound by design (SURVEY 8d): the limiter is the FP64 dependency chain; see fp64"},
        "fp64": {"achieved": ach_tf, "peak": peak64, "unit": "TFLOP/s", "frac": (ach_tf / peak64) if peak64 else None,
                 "flops_per_step": flops_step, "peak_source": "nmpc_measure_fp64_peak (DFMA loop, this GPU, this run)",
                 "iters": it_step, "factorizations": wc["factorizations"], "ls_trials": wc["ls_trials"]},
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the host-buffer arm (0 = as many as --steps)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--ref-batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused-step", action="store_true", help="nmpc_solve + nmpc_step as two launches instead of nmpc_solve_and_step")
    ap.add_argument("--pipelines", type=int, default=0, help="independently pipelined sub-batches per GPU (0 = choose from the batch size, 1 = one batch on one stream)")
    ap.add_argument("--no-lpt", action="store_true", help="disable the library's longest-first scheduling (NMPC_B200_AUTO_ORDER=0)")
    ap.add_argument("--count-work", action="store_true", help="read device work counters every timed step (adds a sync)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
