"""Damaged camera frames against the host side of the capture-side decode, under AddressSanitizer / UBSan.

The bytes cvs_submit_jpeg receives come from a camera (threads.cpp:32-41 hands OpenCV whatever the device delivered), so
the marker parser and the table builder (cudavideostream_b200/csrc/cvs_jpeg_host.hpp) must answer every input with a
status and never read outside the buffer:

* tests/host/jpeg_parse_fuzz.cpp: every truncation of the header, truncations followed by an EOI, and seeded random
  damage of header bytes (segment lengths, table classes, counts, sampling factors ...), each call on an exactly-sized
  heap copy;
* tests/host/jpeg_sim.cpp (the decoder's own __host__ __device__ token functions, executed on the CPU): frames whose
  damaged header still parses -- wrong Huffman / quantisation tables, other sampling -- are decoded; the token loop
  must stay inside its tables and buffers and end with "ok", "unsupported", "not a JPEG" or "wrong number of blocks".
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SAN = ["-fsanitize=address,undefined", "-fno-sanitize-recover=all"]


def _build(tmp, name):
    exe = str(tmp / name)
    src = os.path.join(ROOT, "tests", "host", name + ".cpp")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-Wall", *SAN, "-o", exe, src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0 and b"sanitize" in r.stdout.lower() or r.returncode != 0 and b"asan" in r.stdout.lower():
        pytest.skip("this g++ has no sanitizer runtime")
    assert r.returncode == 0, r.stdout.decode()
    return exe


@pytest.fixture(scope="module")
def fuzz(tmp_path_factory):
    return _build(tmp_path_factory.mktemp("jpeg_fuzz"), "jpeg_parse_fuzz")


@pytest.fixture(scope="module")
def sim_asan(tmp_path_factory):
    return _build(tmp_path_factory.mktemp("jpeg_sim_asan"), "jpeg_sim")


def _cases(tmp_path):
    yield os.path.join(GOLDEN, "k1_f1.jpg")  # the reference's camera frame
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    for n in ("q90_420_opt_300x200", "q90_gray_123x77", "q85_422_rst8_640x480", "q85_444_161x97"):
        p = str(tmp_path / (n + ".jpg"))
        z[n + "/jpg"].tofile(p)
        yield p


def test_parser_answers_every_damaged_header_with_a_status(fuzz, tmp_path):
    for i, path in enumerate(_cases(tmp_path)):
        r = subprocess.run([fuzz, path, "30000", str(0xC0DA5EED + i)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
        log = r.stdout.decode()
        assert r.returncode == 0, f"{os.path.basename(path)}: {log[-2000:]}"
        calls, ok, notjpeg, unsupported = (int(v) for v in log.split()[1::2])
        # the damage is real (a good share of the calls is refused) and not everything is refused
        assert calls == ok + notjpeg + unsupported and notjpeg > calls // 10 and ok > calls // 10, log


def test_token_loop_stays_in_bounds_with_damaged_tables(sim_asan, tmp_path):
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    rng = np.random.default_rng(11)
    seen = {}
    for name in ("q95_420_8x8", "q100_420_37x29", "q90_gray_123x77"):
        good = z[name + "/jpg"].tobytes()
        sos = good.index(b"\xff\xda")
        for trial in range(60):
            m = bytearray(good)
            for k in rng.integers(2, sos, size=int(rng.integers(1, 4))):
                m[k] = int(rng.integers(0, 256))
            p = str(tmp_path / "m.jpg")
            with open(p, "wb") as f:
                f.write(m)
            r = subprocess.run([sim_asan, p, "256", str(tmp_path / "o.bin"), "1"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
            # 0 decoded, 3 unsupported form, 4 refused by the parser, 7 the stream does not hold the image's blocks
            assert r.returncode in (0, 3, 4, 7), f"{name} trial {trial}: rc {r.returncode}\n{r.stdout.decode()[-1500:]}"
            seen[r.returncode] = seen.get(r.returncode, 0) + 1
    assert seen.get(0, 0) > 20 and seen.get(4, 0) > 10, seen
