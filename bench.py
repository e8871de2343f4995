#!/usr/bin/env python
"""bench.py -- throughput of the CUDAVideoStream hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through libcvs_b200.so)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores

Workload (BASELINE.json configs[1]): 1080p BGR24 synthetic 300-frame sequences at change densities
1 % / 10 % / 50 %, thresholded diff + negative feedback + ordered compaction.  One STEP = one pass over the
three 300-frame sequences (900 frames, 5.6 GB of frames resident in HBM, so every frame is L2-cold).
`value` = frames/s with the frames already in HBM (device-resident sequence API); `e2e` = frames/s through the
pipelined drop-in call (cvs_submit/cvs_wait) with frames in pinned HOST memory, H2D and payload D2H inside
the timed region.  Multi-GPU: one process per GPU, independent camera streams per rank, no collective on the
data path (weak scaling); torch.distributed is used only for the barrier and the max-over-ranks time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
N = 3 * W * H
DENSITIES_PPM = (10000, 100000, 500000)
SEQ_FRAMES = 300
METRIC = "1080p frames/sec (diff+compact, 300-frame sequences at 1%/10%/50% change density)"
THR = 20


def algorithmic_bytes(nframes: int, sum_pos: int) -> int:
    """(2 + 6c) N + 4 per frame = read cur N + read ref N + write ref cN + payload 5cN + count."""
    return nframes * (2 * N + 4) + 6 * sum_pos


def hbm_model_bytes(nframes: int, sum_pos: int) -> int:
    """Bytes that must cross HBM when the reference stays on chip: read cur N + payload 5 pos + count."""
    return nframes * (N + 4) + 5 * sum_pos


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU loop (oracle restatement of tests/cuda_streaming/test.cu:560-576), one
# independent camera stream per host thread.
# ------------------------------------------------------------------------------------------------------
def host_ring(nring: int, density_ppm: int, seed: int):
    from cudavideostream_b200 import synth
    base = synth.base_frame(W, H, seed)
    ring = np.empty((nring, N), dtype=np.uint8)
    prev = base
    for t in range(nring):
        prev = synth.next_frame(prev, seed, t, density_ppm)
        ring[t] = prev
    # walked back and forth, so that consecutive frames are always one synthetic step apart
    ring = np.ascontiguousarray(np.concatenate([ring, ring[-2:0:-1]]))
    return base, ring


_RINGS = None


def cpu_rings():
    """Host frame rings of the three densities (generated once per process: the numpy camera is slow)."""
    global _RINGS
    if _RINGS is None:
        _RINGS = [host_ring(4, d, 0xC0DA5EED) for d in DENSITIES_PPM]
    return _RINGS


def cpu_run(target_seconds: float, threads: int):
    """Times the CPU path on a bounded sample.  Returns dict(value=frames/s, cores, sample, ...)."""
    from oracle import oracle as orc
    orc.build()
    rings = cpu_rings()
    # calibrate: one frame per thread per density
    t_cal = sum(orc.bench_diff_compact(r, b, THR, 1, threads)[0] for b, r in rings)
    iters = max(1, int(target_seconds / max(t_cal, 1e-3)))
    iters = min(iters, 50)
    sec = 0.0
    frames = 0
    for b, r in rings:
        s, nf, _ = orc.bench_diff_compact(r, b, THR, iters, threads)
        sec += s
        frames += nf
    return {"value": frames / sec, "seconds": sec, "frames": frames, "cores": threads, "kind": "port",
            "unit": "frames/s",
            "sample": f"{iters} frames x {threads} independent streams x 3 densities (1/10/50 %) of the 1080p "
                      f"workload, oracle/cvs_oracle.c orc_diff_compact -O2, frames in host RAM"}


def cpu_single_thread(frames_per_density: int = 4):
    """the reference's compute runs on ONE thread (server.cpp:70-146) and its Makefile ships
    -O0 (server/Makefile:13).  A few frames per density of the same workload, one thread, -O2 and -O0 builds."""
    from oracle import oracle as orc
    orc.build()
    out = {}
    for name, o0 in (("O2", False), ("O0", True)):
        sec = 0.0
        for b, r in cpu_rings():
            sec += orc.bench_diff_compact(r, b, THR, frames_per_density, 1, o0=o0)[0]
        out["frames_per_s_1thread_" + name] = 3 * frames_per_density / sec
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    t0 = time.time()
    vals = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_run(2.0, threads)
    for _ in range(args.steps):
        vals.append(cpu_run(max(args.ref_seconds / 7.0, args.ref_seconds / args.steps), threads))
    frames = sum(v["frames"] for v in vals)
    sec = sum(v["seconds"] for v in vals)
    value = frames / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s_seq%d_d1_10_50" % ("1080p" if (W, H) == (1920, 1080) else "%dx%d" % (W, H), args.frames),
                       "width": W, "height": H, "threshold": THR},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": vals[0]["sample"], "single_thread": cpu_single_thread()},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    import cudavideostream_b200 as cvs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # one host thread + one context per GPU, pinned buffers NUMA-local to it: bind this rank
        # to the CPUs nearest its GPU before any pinned allocation is first touched
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        except Exception:
            pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cvs.load_library()
    st = torch.cuda.current_stream().cuda_stream
    T = args.frames
    seed = 0xC0DA5EED ^ (rank * 0x9E3779B9)
    cap = (N + 3) // 4 * 4  # worst case: every byte of a frame changes

    # ---- device-resident sequences, one Stream (reference state) per density
    seqs = []
    for d in DENSITIES_PPM:
        frames = torch.empty((T + 1) * N, dtype=torch.uint8, device=dev)
        cvs.synth.base_frame_device(frames.data_ptr(), W, H, seed, st)
        for t in range(T):
            cvs.synth.next_frame_device(frames.data_ptr() + t * N, frames.data_ptr() + (t + 1) * N, W, H, seed, t, d, st)
        torch.cuda.synchronize()
        base = frames[:N].cpu().numpy()
        c = cap
        s = cvs.Stream(W, H, base, threshold=THR, device=local, max_sequence=max(T, 16))
        seqs.append({"d": d, "frames": frames, "stream": s, "cap": c,
                     "pos": torch.zeros(T, dtype=torch.int32, device=dev),
                     "xs": torch.empty(T * c, dtype=torch.int32, device=dev),
                     "diff": torch.empty(T * c, dtype=torch.uint8, device=dev)})

    def run_seq(q):
        q["stream"].run_sequence_device(q["frames"].data_ptr() + N, N, T, q["pos"].data_ptr(), q["xs"].data_ptr(),
                                        q["diff"].data_ptr(), q["cap"], cuda_stream=st)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        for q in seqs:
            run_seq(q)
    torch.cuda.synchronize()
    for q in seqs:
        q["stream"].sequence_status()
    launches0 = sum(q["stream"].launch_count() for q in seqs)

    sampler = ClockSampler(local)
    sampler.start()
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in seqs]
           for _ in range(args.steps)]
    sum_pos = [0 for _ in seqs]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        for i, q in enumerate(seqs):
            evs[k][i][0].record()
            run_seq(q)
            evs[k][i][1].record()
    e1.record()
    barrier()
    clocks = sampler.result()
    elapsed_ms = e0.elapsed_time(e1)
    for q in seqs:
        q["stream"].sequence_status()
    launches = sum(q["stream"].launch_count() for q in seqs) - launches0
    # per-launch durations of the stream kernel and the algorithmic bytes each launch moved
    per_density = []
    tot_alg = tot_hbm = 0
    tot_ms = 0.0
    for i, q in enumerate(seqs):
        ms = [evs[k][i][0].elapsed_time(evs[k][i][1]) for k in range(args.steps)]
        sp = int(q["pos"].to(torch.int64).sum().item())  # identical in every timed step? (ring restarts): last step
        alg, hbm = algorithmic_bytes(T, sp), hbm_model_bytes(T, sp)
        mean_ms = float(np.mean(ms))
        per_density.append({"density_ppm": q["d"], "realised_c": sp / (T * N), "ms_per_launch": mean_ms,
                            "frames_per_s": T / (mean_ms * 1e-3), "algorithmic_GBps": alg / (mean_ms * 1e-3) / 1e9,
                            "hbm_model_GBps": hbm / (mean_ms * 1e-3) / 1e9})
        tot_alg += alg
        tot_hbm += hbm
        tot_ms += mean_ms
    frames_done = args.steps * T * len(seqs)

    # whole-job view: max over ranks of the device time, sum over ranks of the work (no data-path collective)
    elapsed_s, (frames_total, launches_total) = cvs.sharding.reduce_job(elapsed_ms * 1e-3, [frames_done, launches], dev)
    elapsed_ms = elapsed_s * 1e3
    value = frames_total / elapsed_s

    # ---- end to end through the drop-in call: frames in pinned host memory, pipelined submit/wait
    e2e = e2e_run(cvs, torch, dist, args, seqs, local, world, barrier)

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = tot_alg / (tot_ms * 1e-3) / 1e9
    achieved_hbm = tot_hbm / (tot_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            # ncu dram__bytes_read.sum + dram__bytes_write.sum per frame (mean of the three densities) x frames per launch
            traffic = json.load(fh)["dram_bytes_per_frame_mean"] * T if (W, H) == (1920, 1080) else None
    except Exception:
        pass

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "%s_seq%d_d1_10_50" % ("1080p" if (W, H) == (1920, 1080) else "%dx%d" % (W, H), T), "width": W, "height": H, "threshold": THR,
                           "frames_per_sequence": T, "sequences_per_step": len(seqs), "streams_per_gpu": len(seqs),
                           "l2": "inputs larger than L2: 3 x %.2f GB device-resident frame sequences per step" % (T * N / 1e9),
                           "per_density": per_density},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0, "traffic": traffic,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                             "kernel": "cvs::k_stream<0,false,%s> (one launch = one %d-frame sequence)" % ("true" if (N + 95) // 96 <= 512 * torch.cuda.get_device_properties(dev).multi_processor_count else "false", T),
                             "achieved_hbm_model": achieved_hbm, "frac_hbm_model": achieved_hbm / peak,
                             "note": "achieved = SURVEY 8(d) algorithmic bytes (2+6c)N+4 per frame / event-timed launch; "
                                     "hbm_model counts only bytes that must cross HBM when the reference stays on chip "
                                     "((1+5c)N+4)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_total}
        if world == 1 and not args.no_cpu:
            cb = cpu_run(args.cpu_seconds, os.cpu_count() or 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["single_thread"] = cpu_single_thread()
        print(json.dumps(line), flush=True)
    for q in seqs:
        q["stream"].close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def e2e_run(cvs, torch, dist, args, seqs, local, world, barrier):
    """frames/s through cvs_submit/cvs_wait with HOST frames: per frame an H2D of N bytes from pinned memory and a
    D2H of the count + payload.  The host ring of each density holds R consecutive frames walked back and forth so
    that consecutive submissions are always one synthetic step apart."""
    R = args.e2e_ring
    frames_per_density = args.e2e_frames
    rings = []
    for q in seqs:
        hb = cvs.alloc_host(R * N)
        arr = hb.array()
        src = q["frames"]
        for t in range(R):
            arr[t * N:(t + 1) * N] = src[(t + 1) * N:(t + 2) * N].cpu().numpy()
        base = src[:N].cpu().numpy()
        s = cvs.Stream(W, H, base, threshold=THR, device=local)
        out = [(cvs.alloc_host(N + 32), cvs.alloc_host(4 * N + 32), (C.c_uint * 1)()) for _ in range(4)]
        rings.append({"hb": hb, "stream": s, "out": out})
    order = list(range(R)) + list(range(R - 2, 0, -1))

    def run(nframes, count):
        d2h = 0
        for q in rings:
            s, hb, out = q["stream"], q["hb"], q["out"]
            pending = []
            for i in range(nframes):
                fb, xb, pb = out[i % 4]
                if len(pending) == 4:
                    tk, pp = pending.pop(0)
                    s.wait(tk)
                    d2h += 4 + 5 * pp[0]
                # frames stay in the pinned capture ring; the payload bytes go to the slot's own pinned buffer
                src = order[i % len(order)]
                pending.append((s.submit_io_raw(hb.ptr + src * N, fb.ptr, None, "", C.addressof(pb), xb.ptr), pb))
            for tk, pp in pending:
                s.wait(tk)
                d2h += 4 + 5 * pp[0]
        return d2h

    run(min(8, frames_per_density), False)
    launches0 = sum(q["stream"].launch_count() for q in rings)
    barrier()
    t0 = time.perf_counter()
    d2h = run(frames_per_density, True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nfr = frames_per_density * len(rings)
    launches = sum(q["stream"].launch_count() for q in rings) - launches0
    dt, (nfr, d2h, launches) = cvs.sharding.reduce_job(dt, [nfr, d2h, launches], torch.device("cuda", local))
    tm = rings[1]["stream"].timing()
    res = {"value": nfr / dt, "unit": "frames/s", "h2d_bytes_per_step": N * nfr, "d2h_bytes_per_step": d2h,
           "frames": nfr, "seconds": dt, "gpu_launches": launches,
           "note": "one e2e step = %d frames per density x 3 densities per GPU through cvs_submit_io/cvs_wait from a "
                   "pinned host ring" % frames_per_density,
           "last_frame_us": tm}
    for q in rings:
        q["stream"].close()
    return res


def main():
    global W, H, N, METRIC
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=SEQ_FRAMES, help="frames per sequence")
    ap.add_argument("--e2e-frames", type=int, default=300, help="frames per density in the end-to-end leg")
    ap.add_argument("--e2e-ring", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-seconds", type=float, default=20.0, help="--impl reference: CPU seconds over all steps")
    ap.add_argument("--width", type=int, default=W, help="frame width (default: the 1080p headline workload)")
    ap.add_argument("--height", type=int, default=H)
    args = ap.parse_args()
    if (args.width, args.height) != (W, H):
        # side workloads (e.g. BASELINE config 5's 3840x2160 streams); the headline stays 1080p
        W, H = args.width, args.height
        N = 3 * W * H
        METRIC = "%dx%d frames/sec (diff+compact, %d-frame sequences at 1%%/10%%/50%% change density)" % (W, H, args.frames)
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
