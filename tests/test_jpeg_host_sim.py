"""The GPU JPEG decoder's entropy stage, executed on the CPU: tests/host/jpeg_sim.cpp compiles the very same
__host__ __device__ token functions and table builder as the CUDA kernels (cudavideostream_b200/csrc/cvs_jpeg.cuh,
cvs_jpeg_host.hpp) with plain g++ and runs k_entropy's round structure sequentially -- guessed entry states, exit states
handed to the successor, re-decode until nothing changes, prefix sums, write pass.  Its coefficients must equal the
sequential decode of the oracle (oracle/jpeg_oracle.c: jdhuff.c's decode_mcu) for every subsequence length."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("jpeg_sim") / "jpeg_sim")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "host", "jpeg_sim.cpp")])
    return exe


def _run(sim, jpg_path, sub_bits, tmp_path, hypotheses=1):
    out = str(tmp_path / "coef.bin")
    r = subprocess.run([sim, jpg_path, str(sub_bits), out, str(hypotheses)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return r.returncode, r.stdout.decode(), out


@pytest.mark.parametrize("sub_bits,hypotheses", [(64, 0), (1024, 0), (8192, 0), (256, 1), (1024, 1), (2048, 1)])
def test_parallel_entropy_decode_of_a_camera_frame_equals_the_sequential_one(sim, oracle, tmp_path, sub_bits, hypotheses):
    path = os.path.join(GOLDEN, "k1_f1.jpg")
    rc, log, out = _run(sim, path, sub_bits, tmp_path, hypotheses)
    assert rc == 0, log
    with open(path, "rb") as f:
        ref = oracle.jpeg_coefficients(f.read())
    got = np.fromfile(out, dtype=np.int16).reshape(-1, 64)
    assert got.shape == ref.shape == (120 * 68 * 6, 64)
    assert np.array_equal(got, ref)
    # the rounds until no state changes stay far below the number of subsequences (self-synchronisation works)
    rounds = int(log.strip().splitlines()[-1].split()[1])
    nsub = int(log.strip().splitlines()[-1].split()[3])
    assert rounds < max(16, nsub // 20), log.strip().splitlines()[-1]
    if hypotheses and sub_bits >= 1024:
        # the phase hypotheses find every entry state of the camera frame: the first round only confirms them
        assert rounds == 1 and "hypotheses: 0 of" in log, log


def test_other_samplings_and_sizes(sim, oracle, tmp_path):
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    for n in sorted({k.split("/")[0] for k in z.files}):
        jp = str(tmp_path / (n + ".jpg"))
        z[n + "/jpg"].tofile(jp)
        for sub_bits, hyp in ((32, 0), (96, 1), (1024, 1), (1024, 0)):
            rc, log, out = _run(sim, jp, sub_bits, tmp_path, hyp)
            if "_rst" in n:
                assert rc == 3, f"{n}: restart intervals are reported as unsupported ({log})"
                continue
            assert rc == 0, f"{n} S={sub_bits}: {log}"
            ref = oracle.jpeg_coefficients(z[n + "/jpg"].tobytes())
            got = np.fromfile(out, dtype=np.int16).reshape(-1, 64)
            assert np.array_equal(got, ref), f"{n} S={sub_bits}"


def test_frame_without_huffman_tables(sim, oracle, tmp_path):
    from util import strip_dht
    with open(os.path.join(GOLDEN, "k1_f2.jpg"), "rb") as f:
        bare = strip_dht(f.read())
    jp = str(tmp_path / "bare.jpg")
    with open(jp, "wb") as f:
        f.write(bare)
    rc, log, out = _run(sim, jp, 1024, tmp_path)
    assert rc == 0, log
    assert np.array_equal(np.fromfile(out, dtype=np.int16).reshape(-1, 64), oracle.jpeg_coefficients(bare))


@pytest.mark.parametrize("name,dc_len,ac_len", [("q95_420_200x150", 11, 10), ("q90_gray_123x77", 10, 12), ("q90_422_320x240", 16, 16),
                                                ("q85_444_161x97", 4, 9)])
def test_uniform_long_codes(sim, oracle, tmp_path, name, dc_len, ac_len):
    """The same coefficients re-encoded with Huffman tables whose codes are all dc_len / ac_len bits long: dozens of long
    prefixes -- more than the decoder has second-level tables, so the canonical slow path decodes most tokens."""
    from util import transcode_huffman
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    jpg = z[name + "/jpg"].tobytes()
    info, coef = oracle.jpeg_info(jpg), oracle.jpeg_coefficients(jpg)
    t = transcode_huffman(jpg, coef, info["hs"], info["vs"], info["ncomp"], dc_len, ac_len)
    assert np.array_equal(oracle.jpeg_coefficients(t), coef)  # the transcoder itself
    jp = str(tmp_path / "t.jpg")
    with open(jp, "wb") as f:
        f.write(t)
    for sub_bits in (256, 1024):
        rc, log, out = _run(sim, jp, sub_bits, tmp_path)
        assert rc == 0, log
        assert np.array_equal(np.fromfile(out, dtype=np.int16).reshape(-1, 64), coef), f"S={sub_bits}"
