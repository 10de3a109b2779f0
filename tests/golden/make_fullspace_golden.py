#!/usr/bin/env python
"""Fixture generator: solutions of golden instances computed by the INDEPENDENT full-space restatement
(oracle/ipm_fullspace.py: full-space KKT system, LDL^T inertia count, torch.autograd derivatives) -- not by the C++ oracle.
tests/test_gpu_parity.py::test_fullspace_fixtures compares the CUDA kernel with them directly, so the kernel is checked
against two implementations that share no code.  (CPU only, ~1 min on 8 cores.)

    python tests/golden/make_fullspace_golden.py        ->  tests/golden/fullspace_solves.npz
"""
import concurrent.futures as cf
import multiprocessing as mp
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
CASES = [("t_trajectory", 0), ("t_trajectory", 8), ("nmpc_tt", 0), ("nmpc_tt", 2), ("nmpc_tt", 4), ("nmpc_tt", 10),
         ("race_track_2", 1), ("10_obstacles", 2), ("plus_trajectory", 5), ("gimbal_less", 0)]
STATUS = {"Solve_Succeeded": 0, "Maximum_Iterations_Exceeded": 1}


def work(case):
    name, idx = case
    import torch
    torch.set_num_threads(1)
    import b200nmpc
    from oracle import ipm_fullspace, nlp_ref
    sc = b200nmpc.SCENARIOS[name]
    G = np.load(ROOT / "tests" / "golden" / f"solves_{name}.npz")
    rs = nlp_ref.RefSpec5(sc.T, sc.N) if sc.model == 1 else nlp_ref.RefSpec(T=sc.T, N=sc.N, obstacles=sc.obstacles, uav_r=sc.uav_r, w1=sc.w1, w2=sc.w2)
    q = ipm_fullspace.solve(ipm_fullspace.Problem(rs, G["p"][idx]), G["x0"][idx], *sc.bounds())
    return dict(name=name, idx=idx, status=STATUS[q["status"]], iters=q["iters"], x=q["x"], f=q["f"], lam_x=q["lam_x"], lam_g=q["lam_g"])


def main():
    with cf.ProcessPoolExecutor(max_workers=min(len(CASES), mp.cpu_count()), mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(work, CASES))
    out = dict(names=np.array([r["name"] for r in res]), idx=np.array([r["idx"] for r in res], dtype=np.int32),
               status=np.array([r["status"] for r in res], dtype=np.int32), iters=np.array([r["iters"] for r in res], dtype=np.int32),
               f=np.array([r["f"] for r in res]))
    for k, r in enumerate(res):
        out[f"x_{k}"] = r["x"]; out[f"lam_x_{k}"] = r["lam_x"]; out[f"lam_g_{k}"] = r["lam_g"]
    np.savez_compressed(ROOT / "tests" / "golden" / "fullspace_solves.npz", **out)
    for r in res:
        print(r["name"], r["idx"], "status", r["status"], "iters", r["iters"], "f %.10f" % r["f"])


if __name__ == "__main__":
    main()
