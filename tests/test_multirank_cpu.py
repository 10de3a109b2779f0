"""CPU, world_size 2 over gloo: instance sharding + result gather + counter reduction (the only collectives
the path has).  The per-rank 'solve' is the oracle; sharded results must equal the unsharded ones bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import b200nmpc
    import oracle
    from mpc_implementation_b200 import sharding
    sc = b200nmpc.SCENARIOS["t_trajectory"]
    sp = oracle.make_spec(sc.T, sc.N, sc.n_obs)
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = b200nmpc.random_instances(sc, B, seed=77)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    lo, hi = sharding.shard_range(B, rank, world)
    r = oracle.solve(sp, sc.obstacle_table(), p[lo:hi], x0[lo:hi], lbx, ubx, lbg, ubg, nthreads=2)
    sol = dict(x=torch.from_numpy(r["x"]), f=torch.from_numpy(r["f"]))
    st = dict(return_status=torch.from_numpy(r["status"]), iter_count=torch.from_numpy(r["iters"]))
    rec = sharding.gather_rows(sharding.result_record(sol, st), B, world)
    cnt = sharding.sum_counters([(r["status"] == 0).sum(), r["iters"].sum()], "cpu")
    tmax = sharding.max_over_ranks(float(rank + 1), "cpu")
    if rank == 0:
        np.savez(Path(out_dir) / "gathered.npz", rec=rec.numpy(), cnt=cnt.numpy(), tmax=tmax)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_unsharded(tmp_path, oracle_mod, pkg):
    B, world = 9, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npz")
    sc = pkg.SCENARIOS["t_trajectory"]
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs)
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=77)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    r = oracle_mod.solve(sp, sc.obstacle_table(), p, x0, lbx, ubx, lbg, ubg)
    rec = np.concatenate([r["x"][:, :6], r["f"][:, None], r["status"][:, None].astype(float), r["iters"][:, None].astype(float)], axis=1)
    assert np.array_equal(got["rec"], rec)                      # bit for bit: instances are independent
    assert got["cnt"][0] == (r["status"] == 0).sum() and got["cnt"][1] == r["iters"].sum()
    assert got["tmax"] == 2.0
