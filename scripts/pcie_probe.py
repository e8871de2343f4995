"""Where does the end-to-end path stop scaling?  Separates the host side of cvs_submit_io into its parts on an
N-GPU box: pinned H2D only, D2H only, both directions at once (two streams), and the same with the per-frame call
pattern of the drop-in path (6.2 MB frames, one cudaMemcpyAsync per frame).  One rank per GPU:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/pcie_probe.py

Rank 0 prints one JSON line with aggregate GB/s (sum over ranks of bytes / max over ranks of time)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

N = 3 * 1920 * 1080
FRAMES = 64
host_in = torch.empty(FRAMES * N, dtype=torch.uint8).pin_memory()
host_out = torch.empty(FRAMES * N, dtype=torch.uint8).pin_memory()
host_in.fill_(7)
d_in = torch.empty(FRAMES * N, dtype=torch.uint8, device=dev)
d_out = torch.ones(FRAMES * N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=6):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / reps


def h2d_bulk():
    with torch.cuda.stream(s1):
        d_in.copy_(host_in, non_blocking=True)


def d2h_bulk():
    with torch.cuda.stream(s2):
        host_out.copy_(d_out, non_blocking=True)


def both_bulk():
    h2d_bulk()
    d2h_bulk()


def h2d_frames():
    with torch.cuda.stream(s1):
        for f in range(FRAMES):
            d_in[f * N:(f + 1) * N].copy_(host_in[f * N:(f + 1) * N], non_blocking=True)


def both_frames():
    for f in range(FRAMES):
        with torch.cuda.stream(s1):
            d_in[f * N:(f + 1) * N].copy_(host_in[f * N:(f + 1) * N], non_blocking=True)
        with torch.cuda.stream(s2):
            host_out[f * N:(f + 1) * N].copy_(d_out[f * N:(f + 1) * N], non_blocking=True)


def host_memcpy():
    # host DRAM bandwidth per rank: what a capture thread does when it fills the pinned ring
    host_out.copy_(host_in)


res = {"n_gpus": world, "bytes_per_rank_per_rep": FRAMES * N}
for name, fn, nbytes in (("h2d_bulk", h2d_bulk, FRAMES * N), ("d2h_bulk", d2h_bulk, FRAMES * N),
                         ("both_bulk", both_bulk, 2 * FRAMES * N), ("h2d_per_frame_calls", h2d_frames, FRAMES * N),
                         ("both_per_frame_calls", both_frames, 2 * FRAMES * N), ("host_memcpy", host_memcpy, 2 * FRAMES * N)):
    dt = timed(fn)
    res[name + "_GBps_aggregate"] = round(world * nbytes / dt / 1e9, 2)
    res[name + "_GBps_per_gpu"] = round(nbytes / dt / 1e9, 2)
if rank == 0:
    try:
        res["nproc"] = os.cpu_count()
        res["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except Exception:
        pass
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
