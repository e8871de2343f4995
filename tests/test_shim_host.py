"""The C++ drop-in: a host program written against diff::cuda::CUDACore (include/cvs_cuda_core.hpp) the way
server/src/server.cpp and threads.cpp use it, linked against libcvs_b200.so."""
import os
import struct
import subprocess

import numpy as np
import pytest

from util import CHARS_STR, glyph_atlas, random_sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "shim_host.cpp")


def _build(tmpdir) -> str:
    import cudavideostream_b200 as cvs
    exe = os.path.join(str(tmpdir), "shim_host")
    libdir = os.path.dirname(cvs.library_path())
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", libdir, "-l:libcvs_b200.so", "-Wl,-rpath," + libdir])
    return exe


def test_reference_style_host_compiles_and_links(cvs, tmp_path):
    # a plain g++ translation unit (no nvcc, no CUDA headers) sees only the four reference signatures
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    # with no arguments the program exits before touching CUDA
    assert subprocess.run([exe]).returncode == 2


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 3])
def test_reference_style_host_matches_oracle(cvs, oracle, tmp_path, mode):
    w, h, nframes = 200, 120, 4
    gw, gh = 7, 5
    atlas = glyph_atlas(gw, gh)
    base, frames = random_sequence(w, h, nframes, 0.07, seed=31 + mode)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<3i", w, h, nframes))
        f.write(base.tobytes())
        f.write(frames.tobytes())
        f.write(struct.pack("<3i", gw, gh, len(CHARS_STR)))
        f.write(atlas.tobytes())
    exe = _build(tmp_path)
    env = dict(os.environ, CVS_NOISE_VISUALIZER=str(mode))
    text = "FPS 26 BW:3/1 kbps"
    subprocess.check_call([exe, fin, fout, text], env=env)
    oc = oracle.OracleCore(w, h, base, mode=mode, glyphs=atlas, gw=gw, gh=gh, chars=CHARS_STR)
    n = 3 * w * h
    with open(fout, "rb") as f:
        for t in range(nframes):
            opos, oxs, odiff, oshow, _ = oc.exec_core(frames[t], text)
            (pos,) = struct.unpack("<I", f.read(4))
            xs = np.frombuffer(f.read(4 * pos), dtype=np.int32)
            diff = np.frombuffer(f.read(pos), dtype=np.uint8)
            show = np.frombuffer(f.read(n), dtype=np.uint8)
            assert pos == opos and np.array_equal(xs, oxs) and np.array_equal(diff, odiff), f"frame {t}"
            if mode:
                assert np.array_equal(show, oshow), f"frame {t}"


JPEG_SRC = os.path.join(ROOT, "tests", "host", "jpeg_host.c")


def _build_jpeg_host(tmpdir) -> str:
    import cudavideostream_b200 as cvs
    exe = os.path.join(str(tmpdir), "jpeg_host")
    libdir = os.path.dirname(cvs.library_path())
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), JPEG_SRC, "-o", exe,
                           "-L", libdir, "-l:libcvs_b200.so", "-Wl,-rpath," + libdir])
    return exe


def test_c_capture_thread_compiles_against_the_c_abi(cvs, tmp_path):
    # plain C (gcc -std=c99, no CUDA headers): include/cvs_b200.h is a C header
    exe = _build_jpeg_host(tmp_path)
    assert subprocess.run([exe]).returncode == 2


@pytest.mark.gpu
def test_c_capture_thread_feeds_the_cameras_jpeg_frames(cvs, oracle, tmp_path):
    """tests/host/jpeg_host.c: the reference's own camera frames as MJPG buffers through cvs_submit_jpeg from a C program,
    four tickets in flight; the payload file equals the oracle's diff of the oracle's (= OpenCV's) pixels."""
    golden = os.path.join(ROOT, "tests", "golden")
    names = ["k1_f1.jpg", "k1_f2.jpg", "k1_f1.jpg", "k1_f2.jpg", "k1_f2.jpg", "k1_f1.jpg"]
    exe = _build_jpeg_host(tmp_path)
    fout = str(tmp_path / "payload.bin")
    env = dict(os.environ, CVS_JPEG_DECODER="own")
    subprocess.check_call([exe, "1920", "1080", fout] + [os.path.join(golden, n) for n in names], env=env)
    frames = {}
    for n in set(names):
        with open(os.path.join(golden, n), "rb") as f:
            frames[n] = oracle.jpeg_decode_bgr(f.read()).reshape(-1)
    ref = np.zeros(3 * 1920 * 1080, dtype=np.uint8)
    with open(fout, "rb") as f:
        for t, n in enumerate(names):
            opos, oxs, odiff, ref, _ = oracle.diff_compact(frames[n], ref, 20)
            (pos,) = struct.unpack("<I", f.read(4))
            xs = np.frombuffer(f.read(4 * pos), dtype=np.int32)
            diff = np.frombuffer(f.read(pos), dtype=np.uint8)
            assert pos == opos and np.array_equal(xs, oxs) and np.array_equal(diff, odiff), f"frame {t} ({n})"
        assert f.read() == b""
