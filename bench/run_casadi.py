#!/usr/bin/env python
"""Recorder for the REAL reference path (CasADi + IPOPT), guarded: BASELINE.md section 3.5 / SURVEY.md section 7.3-1.

CasADi is not installable in the build image (no network, no wheel), so in this repository the oracle's parity is
"unpinned".  Anyone who has CasADi can pin it with this script:

    python bench/run_casadi.py --reference-root /path/to/MPC-Implementation [--scripts NMPC_TT.py ...] [--max-steps 60]

It runs each reference script UNMODIFIED (runpy), with two shims that change no arithmetic:
  * `casadi.nlpsol` is wrapped so that the Function it returns records, for every call of the closed loop
    (Python/NMPC_TT.py:358-365), the inputs  p, x0  and the outputs  x, f, g, lam_x, lam_g  plus IPOPT's
    return_status / iter_count from solver.stats();
  * matplotlib / mayavi / tvtk (plotting only, NMPC_TT.py:6-9, Race Track 2.py:10-12) are replaced by inert stubs.
The records go to tests/golden/casadi_<scenario>.npz in the layout of the oracle's own fixtures (solves_<scenario>.npz);
tests/test_oracle_solve.py and tests/test_gpu_parity.py prefer those files when they exist and then assert the north
star's tolerances against REAL IPOPT output (u0* 1e-6, f* 1e-8, identical status).  It also prints the reference's
own throughput (solves/s of the recorded loop), the number BASELINE.md's table row 1 asks for.
Without CasADi it prints "CasADi unavailable" and exits 0.
"""
from __future__ import annotations

import argparse
import runpy
import sys
import time
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SCRIPTS = {"NMPC_TT.py": "nmpc_tt", "T_Trajectory.py": "t_trajectory", "Plus Trajectory.py": "plus_trajectory",
           "Race Trajectory 1.py": "race_trajectory_1", "Race Track 2.py": "race_track_2", "10_obstacles.py": "10_obstacles"}
# CasADi's return_status strings -> the codes of include/nmpc_b200.h
STATUS = {"Solve_Succeeded": 0, "Solved_To_Acceptable_Level": 0, "Maximum_Iterations_Exceeded": 1, "Restoration_Failed": 2,
          "Search_Direction_Becomes_Too_Small": 3, "Invalid_Number_Detected": 4, "Error_In_Step_Computation": 5,
          "Infeasible_Problem_Detected": 6}


class _Stop(Exception):
    pass


class _Stub(types.ModuleType):
    """Inert stand-in for a plotting module: every attribute is a callable that returns another stub."""
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(name)

    def __call__(self, *a, **k):
        return _Stub("call")

    def __iter__(self):
        return iter(())

    def __getitem__(self, i):
        return _Stub("item")


def record_script(ca, path: Path, max_steps: int):
    rec = {k: [] for k in ("p", "x0", "x", "f", "g", "lam_x", "lam_g", "status", "iters", "status_name", "seconds")}
    real_nlpsol = ca.nlpsol

    def nlpsol(name, plugin, *a, **k):
        solver = real_nlpsol(name, plugin, *a, **k)

        class Recording:
            def __call__(self_, **kw):
                t0 = time.perf_counter()
                sol = solver(**kw)
                dt = time.perf_counter() - t0
                st = solver.stats()
                vec = lambda v: np.asarray(ca.DM(v).full(), dtype=np.float64).reshape(-1)
                rec["p"].append(vec(kw["p"])); rec["x0"].append(vec(kw["x0"]))
                for q in ("x", "g", "lam_x", "lam_g"):
                    rec[q].append(vec(sol[q]))
                rec["f"].append(float(sol["f"]))
                rec["status_name"].append(st.get("return_status", "?")); rec["status"].append(STATUS.get(st.get("return_status", "?"), -1))
                rec["iters"].append(int(st.get("iter_count", -1))); rec["seconds"].append(dt)
                if max_steps and len(rec["f"]) >= max_steps:
                    raise _Stop()
                return sol

            def __getattr__(self_, n):
                return getattr(solver, n)
        return Recording()

    ca.nlpsol = nlpsol
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation", "mpl_toolkits", "mpl_toolkits.mplot3d",
              "mayavi", "mayavi.mlab", "tvtk", "tvtk.api", "tvtk.tools", "tvtk.tools.visual"):
        sys.modules.setdefault(m, _Stub(m))
    try:
        runpy.run_path(str(path), run_name="__main__")
    except _Stop:
        pass
    finally:
        ca.nlpsol = real_nlpsol
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference-root", default="/root/reference")
    ap.add_argument("--scripts", nargs="*", default=list(SCRIPTS))
    ap.add_argument("--max-steps", type=int, default=60, help="closed-loop steps to record per script (0 = the script's own loop_run)")
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    args = ap.parse_args()
    try:
        import casadi as ca
    except Exception as e:      # noqa: BLE001
        print(f"CasADi unavailable ({type(e).__name__}: {e}) -- reference path not timed, no golden vectors recorded")
        return 0
    out = Path(args.out); out.mkdir(parents=True, exist_ok=True)
    for name in args.scripts:
        path = Path(args.reference_root) / "Python" / name
        if not path.exists():
            print(f"{path}: not found, skipped")
            continue
        rec = record_script(ca, path, args.max_steps)
        n = len(rec["f"])
        if n == 0:
            print(f"{name}: the script made no solver call")
            continue
        scen = SCRIPTS.get(name, Path(name).stem.lower())
        np.savez_compressed(out / f"casadi_{scen}.npz", p=np.array(rec["p"]), x0=np.array(rec["x0"]), x=np.array(rec["x"]), f=np.array(rec["f"]),
                            g=np.array(rec["g"]), lam_x=np.array(rec["lam_x"]), lam_g=np.array(rec["lam_g"]), status=np.array(rec["status"], dtype=np.int32),
                            iters=np.array(rec["iters"], dtype=np.int32), status_name=np.array(rec["status_name"]), seconds=np.array(rec["seconds"]),
                            n_loop=n, casadi_version=ca.__version__)
        sec = float(np.sum(rec["seconds"]))
        conv = int(np.sum(np.array(rec["status"]) == 0))
        print(f"{name}: {n} solves recorded -> {out / f'casadi_{scen}.npz'}; CasADi {ca.__version__} / IPOPT on 1 core: "
              f"{n / sec:.1f} solves/s ({conv / sec:.1f} converged solves/s), mean {np.mean(rec['iters']):.1f} iterations, "
              f"status histogram {dict(zip(*np.unique(rec['status_name'], return_counts=True)))}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
