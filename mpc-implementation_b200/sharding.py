"""Multi-GPU layout: independent NLP instances are partitioned across ranks; nothing is exchanged on the
solve path.  The reference has a single UAV per process (Python/NMPC_TT.py:153-255 couples nothing across
instances), so the only collectives are a gather of per-instance results and a sum of counters.

One process per GPU (torch.distributed, NCCL on GPUs / gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of B instances owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, B: int, world: int) -> torch.Tensor:
    """All-gather per-instance rows (first dim = local instances) back into global instance order."""
    if world == 1:
        return local
    sizes = [shard_range(B, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def result_record(sol: Dict[str, torch.Tensor], stats: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Per-instance record [u0*(6), f, status, iters] = 9 doubles."""
    return torch.cat([sol["x"][:, :6], sol["f"][:, None], stats["return_status"].to(torch.float64)[:, None],
                      stats["iter_count"].to(torch.float64)[:, None]], dim=1)


def sum_counters(values, device) -> torch.Tensor:
    """All-reduce (sum) of scalar counters, e.g. (converged, iterations, fov_error_sum)."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
