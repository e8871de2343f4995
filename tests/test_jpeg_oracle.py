"""Pins oracle/jpeg_oracle.c (the CPU restatement of the capture-side decode: OpenCV's libjpeg-turbo defaults, which is
what server/src/threads.cpp:32-41 and tests/noise_filter_benchmark/v2.cu:195-198 run) against outputs of the real decoder:
digests of cv2's pixels for the reference's own camera frames (tests/golden/k1_f1_f2.json, made from the reference's
files by make_golden.py) and for the re-encodings of make_jpeg_cases.py; where cv2 is importable, against cv2 live."""
import hashlib
import json
import os

import numpy as np
import pytest

from util import strip_dht as _strip_dht

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _cases():
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    return z, sorted({k.split("/")[0] for k in z.files})


def test_oracle_decodes_the_reference_camera_frames_like_opencv(oracle):
    with open(os.path.join(GOLDEN, "k1_f1_f2.json")) as f:
        k1 = json.load(f)
    frames = []
    for name, key in (("k1_f1.jpg", "sha256_f1"), ("k1_f2.jpg", "sha256_f2")):
        with open(os.path.join(GOLDEN, name), "rb") as f:
            img = oracle.jpeg_decode_bgr(f.read())
        assert img.shape == (1080, 1920, 3)
        assert hashlib.sha256(img.tobytes()).hexdigest() == k1[key], name
        frames.append(img)
    # K1 (REPORT/report.tex:2594): 369,350 changed bytes between the two frames -- now from the bitstreams alone
    assert oracle.count_difference(frames[0], frames[1], 20) == k1["changed_bytes"] == 369350


def test_oracle_matches_opencv_on_other_samplings_qualities_sizes_and_restart_intervals(oracle):
    z, names = _cases()
    assert len(names) >= 13
    for n in names:
        img = oracle.jpeg_decode_bgr(z[n + "/jpg"].tobytes())
        w, h = (int(v) for v in z[n + "/wh"])
        assert img.shape == (h, w, 3), n
        assert hashlib.sha256(img.tobytes()).digest() == z[n + "/sha"].tobytes(), n


def test_oracle_matches_opencv_live(oracle):
    cv2 = pytest.importorskip("cv2")
    z, names = _cases()
    for n in names:
        ref = cv2.imdecode(z[n + "/jpg"], cv2.IMREAD_COLOR)
        assert np.array_equal(oracle.jpeg_decode_bgr(z[n + "/jpg"].tobytes()), ref), n


def test_oracle_rejects_what_it_does_not_cover(oracle):
    with pytest.raises(ValueError):
        oracle.jpeg_decode_bgr(b"\x55" * 64)
    with open(os.path.join(GOLDEN, "k1_f1.jpg"), "rb") as f:
        j = bytearray(f.read())
    j[j.index(b"\xff\xc0") + 1] = 0xC2  # progressive SOF
    with pytest.raises(ValueError):
        oracle.jpeg_decode_bgr(bytes(j))


def test_frames_without_dht_use_the_standard_tables(oracle):
    """A camera's MJPG frame may carry no Huffman tables (the standard ones are implied); OpenCV decodes such frames, and
    so must the oracle -- to the same pixels as the frame that spells the standard tables out."""
    with open(os.path.join(GOLDEN, "k1_f2.jpg"), "rb") as f:
        j = f.read()
    bare = _strip_dht(j)
    assert len(bare) < len(j) and b"\xff\xc4" not in bare[:600]
    with open(os.path.join(GOLDEN, "k1_f1_f2.json")) as f:
        k1 = json.load(f)
    assert hashlib.sha256(oracle.jpeg_decode_bgr(bare).tobytes()).hexdigest() == k1["sha256_f2"]
    cv2 = pytest.importorskip("cv2")
    assert np.array_equal(cv2.imdecode(np.frombuffer(bare, np.uint8), cv2.IMREAD_COLOR), oracle.jpeg_decode_bgr(bare))
