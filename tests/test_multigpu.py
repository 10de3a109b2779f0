"""GPU, world size 2 over NCCL (skipped with fewer than two devices): one batch solved on ONE GPU equals the same batch
sharded over two ranks (sharding.shard_range), solved by each rank's own handle and gathered with sharding.gather_rows
(ncclAllGather) -- bit for bit, because instances are independent and the solve path has no collective."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, B, steps, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import b200nmpc
    from mpc_implementation_b200 import sharding
    from mpc_implementation_b200.closed_loop import ClosedLoop
    sc = b200nmpc.SCENARIOS["nmpc_tt"]
    p, vw = b200nmpc.random_instances(sc, B, seed=77)
    lo, hi = sharding.shard_range(B, rank, world)
    s = b200nmpc.nlpsol("solver", "ipm", sc, device=rank, max_batch=hi - lo)
    cl = ClosedLoop(s, sc, p[lo:hi], target_vw=vw[lo:hi], device=f"cuda:{rank}")
    conv = 0
    for _ in range(steps):
        sol = cl.step()
        conv += int(s.stats()["success"].sum())
    rec = torch.cat([sharding.result_record(sol, s.stats()), cl.p, cl.err_sum[:, None]], dim=1)      # [n_local, 9 + 11 + 1]
    allrec = sharding.gather_rows(rec, B, world)                                                   # NCCL all-gather
    cnt = sharding.sum_counters([conv], dev)                                                       # NCCL all-reduce
    if rank == 0:
        np.savez(Path(out_dir) / "gathered.npz", rec=allrec.cpu().numpy(), conv=float(cnt[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_equal_one_gpu(tmp_path, pkg):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from mpc_implementation_b200 import sharding
    from mpc_implementation_b200.closed_loop import ClosedLoop
    B, world, steps = 1001, 2, 4          # odd batch: the two shards differ in size
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, B, steps, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npz")
    sc = pkg.SCENARIOS["nmpc_tt"]
    p, vw = pkg.random_instances(sc, B, seed=77)
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    cl = ClosedLoop(s, sc, p, target_vw=vw)
    conv = 0
    for _ in range(steps):
        sol = cl.step()
        conv += int(s.stats()["success"].sum())
    rec = torch.cat([sharding.result_record(sol, s.stats()), cl.p, cl.err_sum[:, None]], dim=1).cpu().numpy()
    assert got["rec"].shape == rec.shape
    assert np.array_equal(got["rec"], rec)          # u0*, f, status, iterations, plant / target state, FOV error sum
    assert got["conv"] == conv
