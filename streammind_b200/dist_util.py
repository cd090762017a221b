"""Multi-GPU plumbing: one process per GPU (torchrun), one independent video stream per rank, no
data-path collective.  The only exchanges are the barrier around the timed
region, a MAX-reduce of the device time and a gather of per-rank metrics -- the pattern of the
reference's dist.allgather (/root/reference/streammind/dist.py:109-119).  Backend 'nccl' on GPUs,
'gloo' in the CPU tests."""
from __future__ import annotations

import os
from typing import Dict, List

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str, device=None):
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION; bench.py's stdout must be one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_metrics(metrics: Dict[str, float], device="cpu") -> List[Dict[str, float]]:
    """Every rank contributes a small dict of floats (same keys); every rank receives the list."""
    if not dist.is_initialized():
        return [dict(metrics)]
    keys = sorted(metrics)
    t = torch.tensor([float(metrics[k]) for k in keys], dtype=torch.float64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [dict(zip(keys, o.tolist())) for o in out]


def stream_ids_for_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Stream i -> rank i mod world (the rank-slicing idiom of videollama2_arch.py:239-242)."""
    return [i for i in range(n_streams) if i % world == rank]


def aggregate_throughput(per_rank: List[Dict[str, float]], units_key="frames", ms_key="ms") -> float:
    """Whole-job throughput = units all ranks processed / max-over-ranks time."""
    total = sum(m[units_key] for m in per_rank)
    worst = max(m[ms_key] for m in per_rank)
    return total / (worst / 1e3)
