"""Host-side logic that needs no GPU: the synthetic camera twin, the exact integer form of the weighted gray, the
heat-map table, and the multi-process (world_size 2, gloo) job reduction used by bench.py."""
import os
import socket

import numpy as np
import pytest

from cudavideostream_b200 import sharding, synth


def test_synth_is_deterministic_and_hits_the_density():
    b1, f1 = synth.sequence(160, 90, 3, 100000, seed=7)
    b2, f2 = synth.sequence(160, 90, 3, 100000, seed=7)
    assert np.array_equal(b1, b2) and np.array_equal(f1, f2)
    assert not np.array_equal(synth.base_frame(160, 90, 8), b1)
    d = np.abs(f1[0].astype(int) - b1.astype(int))
    changed = (d > 20).mean()
    assert 0.09 < changed < 0.11          # +-U[21,80] with probability 10 %
    assert d[d <= 20].max() <= 3          # everything else drifts by at most 3
    assert len(np.unique(b1)) > 100       # non-degenerate histogram


def test_weighted_gray_integer_shortcut_is_exact(oracle):
    # csrc/cvs_pixel.cuh gray_weighted(): trunc(0.114 B + 0.587 G + 0.299 R in double) == (114 B + 587 G + 299 R) / 1000
    # whenever the sum is not a multiple of 1000.  Exhaustive over all 2^24 pixels against the oracle's C loop.
    b, g, r = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8),
                          np.arange(256, dtype=np.uint8), indexing="ij")
    px = np.stack([b.ravel(), g.ravel(), r.ravel()], axis=1)
    want = oracle.gray_weighted1(px.reshape(-1), 4096, 4096)
    s = 114 * px[:, 0].astype(np.int64) + 587 * px[:, 1].astype(np.int64) + 299 * px[:, 2].astype(np.int64)
    fast = (s // 1000).astype(np.uint8)
    exact = (s % 1000) == 0
    assert np.array_equal(fast[~exact], want[~exact])
    # the multiples of 1000 are the only pixels where double rounding decides (these take the double path on the GPU)
    assert (fast[exact] != want[exact]).sum() > 0
    assert np.all(fast[exact].astype(int) - want[exact].astype(int) >= 0)
    assert np.all(fast[exact].astype(int) - want[exact].astype(int) <= 1)


def test_weighted_gray_magic_multiply_is_exact():
    # csrc/cvs_pixel.cuh gray4(): with M = 4,294,968 the 64-bit product s * M carries s / 1000 in its upper word and
    # "s is a multiple of 1000" in its lower word (< 1,000,000), for every sum s = 114 B + 587 G + 299 R <= 255,000.
    s = np.arange(0, 255001, dtype=np.uint64)
    m = s * np.uint64(4294968)
    hi, lo = (m >> np.uint64(32)).astype(np.int64), (m & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert np.array_equal(hi, s.astype(np.int64) // 1000)
    assert np.array_equal(lo < 1000000, s.astype(np.int64) % 1000 == 0)
    assert lo[s.astype(np.int64) % 1000 == 0].max() <= 179520 and lo[s.astype(np.int64) % 1000 != 0].min() >= 4294967


def test_heat_table_is_monotone_blue_to_red(oracle):
    t = np.array([oracle.heat_pixel(d) for d in range(766)])
    r, g, b = t[:, 0], t[:, 1], t[:, 2]
    assert np.all(np.diff(b[:256]) <= 0) and b[0] == 255 and b[255] <= 1
    assert np.all(np.diff(r[255:511]) >= 0) and r[510] == 255 and np.all(r[:255] == 0)
    assert g[255] == 255 and g[0] == 0
    assert np.all((t >= 0) & (t <= 255))


def test_assign_streams_partitions_exactly():
    for n in (0, 1, 3, 8, 13):
        for w in (1, 2, 4, 8):
            parts = sharding.assign_streams(n, w)
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
            for r in range(w):
                assert all(s % w == r for s in parts[r])
    assert sharding.stream_seed(5, 0) == 5 and sharding.stream_seed(5, 1) != sharding.stream_seed(5, 2)
    with pytest.raises(ValueError):
        sharding.my_streams(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        streams = sharding.my_streams(5, rank, world)
        # every rank "processes" its own streams; rank r takes (r + 1) seconds-equivalents
        frames = 300 * len(streams)
        elapsed, (tot_frames, tot_launches) = sharding.reduce_job(0.5 * (rank + 1), [frames, len(streams)])
        out.put((rank, streams, elapsed, tot_frames, tot_launches))
    finally:
        dist.destroy_process_group()


def test_job_reduction_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    for _, _, elapsed, frames, launches in res:
        assert elapsed == 1.0          # max over ranks
        assert frames == 1500          # sum over ranks
        assert launches == 5


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) works without a GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-seconds", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "frames/s" and j["value"] > 0
    assert j["config"]["workload"] == "1080p_seq300_d1_10_50"
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["cpu_baseline"]["kind"] in ("port", "reference")
