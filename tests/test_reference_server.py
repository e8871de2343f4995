"""The reference's UNMODIFIED server/src/server.cpp, compiled from /root/reference by oracle/build_ref.py behind the
file-fed ThreadsCore stub (tests/host/threads_stub.cpp):

* with -DCPU it is the reference's own CPU filter chain (server.cpp:96-135) -> pins the oracle's A3/A5/A6/A7;
* as shipped (GPU branch) and linked against libcvs_b200.so it is the drop-in boundary exercised by the real caller
  (server.cpp:53 constructor, :139 exec_core, threads.cpp:95 alloc_arrays);
* with the reference's own kernels.cu built for sm_100a it cross-checks the oracle's A1 against the reference's GPU
  code on the same frames (payload compared as a set: kernel2 emits in atomicInc order).

The binaries are built in the build container (where /root/reference exists) and travel to the GPU box.
"""
import os
import subprocess

import numpy as np
import pytest

from util import random_sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def refsrv():
    from oracle import build_ref, ref_server
    if build_ref.available():
        build_ref.build()
    return ref_server


def _need(refsrv, name):
    if refsrv.binary(name) is None:
        pytest.skip(f"oracle/_ref/{name} not built (needs /root/reference in the build container)")


def _oracle_cpu_chain(oracle, frame):
    g3 = oracle.gray_avg3(frame)                       # server.cpp:96-101
    hist = oracle.histogram3(g3)                       # server.cpp:103-106
    thr = oracle.threshold_twomax(hist, 50, 200)       # server.cpp:108-127
    return oracle.binarize(g3, thr), thr               # server.cpp:129-135


@pytest.mark.parametrize("w,h,density", [(7, 5, 0.3), (64, 48, 0.1), (250, 130, 0.05), (1920, 1080, 0.1)])
def test_reference_cpu_chain_matches_oracle(oracle, refsrv, w, h, density):
    _need(refsrv, "ref_server_cpu")
    base, frames = random_sequence(w, h, 2 if w > 1000 else 5, density, seed=w)
    out = refsrv.run("ref_server_cpu", w, h, base, frames)
    for t, f in enumerate(frames):
        want, _ = _oracle_cpu_chain(oracle, f)
        assert np.array_equal(out[t]["data"], want), f"frame {t}"


@pytest.mark.parametrize("levels", [[0], [1], [100], [255], [10, 200], [200, 10], [50, 50, 90], [0, 255, 255], [30, 31, 32, 220]])
def test_reference_cpu_twomax_quirks(oracle, refsrv, levels):
    # histograms that hit the corners of the reference's two-max loop: arg-max at bin 0 (index_sec_max = -1), ties,
    # records before the maximum, both clamps -- through the reference's own compiled loop
    _need(refsrv, "ref_server_cpu")
    w, h = 64, 48
    p = w * h
    px = np.zeros((p, 3), dtype=np.uint8)
    bounds = np.linspace(0, p, len(levels) + 1).astype(int)
    if len(levels) == 3:
        bounds = np.array([0, p // 4, p // 2, p])
    for i, g in enumerate(levels):
        px[bounds[i]:bounds[i + 1]] = g
    frame = px.reshape(-1)
    out = refsrv.run("ref_server_cpu", w, h, frame, frame[None, :])
    want, thr = _oracle_cpu_chain(oracle, frame)
    assert np.array_equal(out[0]["data"], want), f"levels {levels}: oracle threshold {thr}"
    o0 = refsrv.binary("ref_server_cpu_O0")
    if o0:
        assert np.array_equal(refsrv.run("ref_server_cpu_O0", w, h, frame, frame[None, :])[0]["data"], want)


def test_unmodified_server_links_against_the_library(refsrv):
    # the four CUDACore members server.cpp / the stub call are undefined in the binary and defined by libcvs_b200.so
    _need(refsrv, "ref_server_dropin")
    exe = refsrv.binary("ref_server_dropin")
    und = subprocess.run(["nm", "-D", "--undefined-only", exe], stdout=subprocess.PIPE, check=True).stdout.decode()
    for sym in ("_ZN4diff4cuda8CUDACoreC1EPhRNS_5utils5matszEPfiS2_S5_", "_ZN4diff4cuda8CUDACore9exec_coreEPhS2_R",
                "_ZN4diff4cuda8CUDACore12alloc_arraysEPPhS3_S3_PPiii"):
        assert sym in und, sym
    needed = subprocess.run(["readelf", "-d", exe], stdout=subprocess.PIPE, check=True).stdout.decode()
    assert "libcvs_b200.so" in needed


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,mode", [(64, 48, 0), (250, 130, 1), (250, 130, 5), (1920, 1080, 0), (1920, 1080, 2)])
def test_unmodified_server_through_the_library(oracle, refsrv, w, h, mode):
    _need(refsrv, "ref_server_dropin")
    T = 4 if w > 1000 else 6
    base, frames = random_sequence(w, h, T, 0.07, seed=w + mode)
    out = refsrv.run("ref_server_dropin", w, h, base, frames, want_show=bool(mode),
                     env={"CVS_NOISE_VISUALIZER": str(mode)})
    oc = oracle.OracleCore(w, h, base, mode=mode)
    for t, f in enumerate(frames):
        pos, xs, diff, show, _ = oc.exec_core(f)
        assert out[t]["pos"] == pos, f"frame {t}"
        assert np.array_equal(out[t]["xs"], xs) and np.array_equal(out[t]["diff"], diff), f"frame {t}"
        if mode:
            assert np.array_equal(out[t]["show"], show), f"frame {t}: show"
    oc.close()


@pytest.mark.gpu
def test_reference_gpu_kernels_agree_with_the_oracle_as_a_set(oracle, refsrv):
    # the reference's own kernel2 (kernels.cu:289-334) on the same 1080p frames: same count and the same (index,
    # value) pairs as the oracle's A1 once sorted -- kernel2's order is whatever atomicInc gave, and it also
    # compares 5,120 bytes past the end of its buffers (entries with index >= N are its own artefact and are dropped)
    _need(refsrv, "ref_server_refgpu")
    w, h, T = 1920, 1080, 3
    n = 3 * w * h
    base, frames = random_sequence(w, h, T, 0.03, seed=77)
    try:
        out = refsrv.run("ref_server_refgpu", w, h, base, frames, timeout=120)
    except RuntimeError as e:
        pytest.skip(f"the reference's kernels faulted on this GPU (known out-of-bounds accesses): {e}")
    oc = oracle.OracleCore(w, h, base)
    for t, f in enumerate(frames):
        pos, xs, diff, _, _ = oc.exec_core(f)
        rx, rd = out[t]["xs"], out[t]["diff"]
        keep = (rx >= 0) & (rx < n)
        rx, rd = rx[keep], rd[keep]
        order = np.argsort(rx, kind="stable")
        assert rx.size == pos, f"frame {t}: {rx.size} in-range entries vs {pos}"
        assert np.array_equal(rx[order], xs) and np.array_equal(rd[order], diff), f"frame {t}"
    oc.close()
