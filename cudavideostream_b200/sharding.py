"""Multi-GPU partitioning of the hot path.

All state of the path (reference frame, payload counter, histogram) belongs to ONE camera stream and frame t
depends only on frame t-1 of the same stream, so streams are the shard unit: stream s runs on rank s mod G, one
process per GPU, and NO collective touches the data path -- every rank hands its own (pos, xs, diff) payloads to
its own socket-send threads exactly as server/src/threads.cpp:229-231 does.  torch.distributed is only used for the
start/stop barrier and for reducing the timing (max over ranks) and the work counters (sum over ranks).
"""
from __future__ import annotations

from typing import List, Sequence


def assign_streams(n_streams: int, world_size: int) -> List[List[int]]:
    """stream s -> rank s mod world_size.  Returns the stream ids of every rank."""
    if n_streams < 0 or world_size < 1:
        raise ValueError("need n_streams >= 0 and world_size >= 1")
    return [list(range(r, n_streams, world_size)) for r in range(world_size)]


def my_streams(n_streams: int, rank: int, world_size: int) -> List[int]:
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return assign_streams(n_streams, world_size)[rank]


def stream_seed(base_seed: int, stream_id: int) -> int:
    """Per-stream seed of the synthetic camera: seed ^ stream_id-dependent constant."""
    return (base_seed ^ (stream_id * 0x9E3779B9)) & 0xFFFFFFFFFFFFFFFF


def reduce_job(elapsed_s: float, counters: Sequence[int], device=None):
    """Whole-job view of a timed region: (max over ranks of elapsed_s, element-wise sum over ranks of counters).
    Without an initialised process group this is the identity (single GPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_s), [int(c) for c in counters]
    t = torch.tensor([elapsed_s], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor(list(counters), dtype=torch.int64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), [int(v) for v in c.tolist()]
