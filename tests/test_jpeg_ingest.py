"""Capture-side decode on the GPU: cvs_submit_jpeg / cvs_decode_jpeg_device take the camera's
JPEG bitstream (the reference's camera delivers MJPG and OpenCV decodes it, server/src/threads.cpp:32-41) and decode it
on the device with the library's own kernels (cudavideostream_b200/csrc/cvs_jpeg.cuh: parallel Huffman decode through
self-synchronisation, jidctint IDCT, fancy upsampling, fixed-point colour conversion).

Checked bit for bit: the decoded pixels against the digests of what OpenCV (libjpeg-turbo) makes of the reference's own
camera frames and of 13 re-encodings, and against the CPU oracle (oracle/jpeg_oracle.c, itself pinned to cv2 in
tests/test_jpeg_oracle.py); the payload of cvs_submit_jpeg against the oracle's diff of the oracle's pixels -- K1
(369,350 changed bytes, REPORT/report.tex:2594) comes out of the two bitstreams alone."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _jpeg(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def _decode(cvs, jpg, w, h):
    import torch
    n = 3 * w * h
    s = cvs.Stream(w, h, np.zeros(n, dtype=np.uint8))
    d = torch.full((n + 64,), 0xA5, dtype=torch.uint8, device="cuda")
    try:
        s.decode_jpeg_device(jpg, d.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        s.sequence_status()
        out = d.cpu().numpy()
        assert np.all(out[n:] == 0xA5), "the decoder wrote past the frame"
        return out[:n].copy()
    finally:
        s.close()


def test_decoded_pixels_of_the_reference_camera_frames_are_opencvs(cvs, oracle, monkeypatch):
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    with open(os.path.join(GOLDEN, "k1_f1_f2.json")) as f:
        k1 = json.load(f)
    frames = []
    for name, key in (("k1_f1.jpg", "sha256_f1"), ("k1_f2.jpg", "sha256_f2")):
        g = _decode(cvs, _jpeg(name), 1920, 1080)
        ref = oracle.jpeg_decode_bgr(_jpeg(name)).reshape(-1)
        bad = np.flatnonzero(g != ref)
        assert bad.size == 0, f"{name}: {bad.size} bytes differ from the oracle, first at {bad[:5]}"
        assert hashlib.sha256(g.tobytes()).hexdigest() == k1[key], name
        frames.append(g)
    assert oracle.count_difference(frames[0], frames[1], 20) == k1["changed_bytes"] == 369350


def test_camera_frames_without_huffman_tables(cvs, oracle, monkeypatch):
    """MJPG as cameras send it: no DHT segment, the standard tables are implied (OpenCV decodes it; so does the library)."""
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    from util import strip_dht as _strip_dht
    with open(os.path.join(GOLDEN, "k1_f1_f2.json")) as f:
        k1 = json.load(f)
    g = _decode(cvs, _strip_dht(_jpeg("k1_f1.jpg")), 1920, 1080)
    assert hashlib.sha256(g.tobytes()).hexdigest() == k1["sha256_f1"]


@pytest.mark.parametrize("sub_bits", [128, 1024, 4096])
def test_other_samplings_qualities_and_sizes(cvs, oracle, monkeypatch, sub_bits):
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    monkeypatch.setenv("CVS_JPEG_SUB_BITS", str(sub_bits))
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    done = 0
    for n in names:
        w, h = (int(v) for v in z[n + "/wh"])
        jpg = z[n + "/jpg"].tobytes()
        g = _decode(cvs, jpg, w, h)
        assert hashlib.sha256(g.tobytes()).digest() == z[n + "/sha"].tobytes(), n
        done += 1
    assert done >= 16  # includes the streams with restart intervals (k_entropy_restart) and the one-component frame


def test_restart_intervals_at_camera_size(cvs, oracle, monkeypatch):
    """The reference's camera frame re-encoded with restart intervals of 1 MCU, 7 MCUs and one MCU row: every interval is
    decoded by its own thread; the pixels are still OpenCV's."""
    cv2 = pytest.importorskip("cv2")
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    src = cv2.imdecode(np.frombuffer(_jpeg("k1_f1.jpg"), np.uint8), cv2.IMREAD_COLOR)
    for interval, extra in ((1, []), (7, [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]), (120, [])):
        ok, enc = cv2.imencode(".jpg", src, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, interval] + extra)
        assert ok and b"\xff\xdd" in enc.tobytes()[:700]
        ref = cv2.imdecode(enc, cv2.IMREAD_COLOR).reshape(-1)
        assert np.array_equal(oracle.jpeg_decode_bgr(enc.tobytes()).reshape(-1), ref)
        g = _decode(cvs, enc.tobytes(), 1920, 1080)
        bad = np.flatnonzero(g != ref)
        assert bad.size == 0, f"interval {interval}: {bad.size} bytes differ, first at {bad[:5]}"


def test_other_forms_fall_back_to_nvjpeg(cvs, monkeypatch):
    """Progressive JPEG is not the built-in decoder's form: nvJPEG takes it (another decoder: the same picture, not the
    same bits); CVS_JPEG_DECODER=own refuses instead."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = np.clip(np.add.outer(np.arange(144) * 1.2, np.arange(256) * 0.7)[:, :, None] + rng.integers(0, 30, (144, 256, 3)), 0, 255)
    ok, enc = cv2.imencode(".jpg", img.astype(np.uint8), [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert ok
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    with pytest.raises(cvs.CVSError) as e:
        _decode(cvs, enc.tobytes(), 256, 144)
    assert e.value.status == 1
    monkeypatch.delenv("CVS_JPEG_DECODER", raising=False)
    try:
        g = _decode(cvs, enc.tobytes(), 256, 144)
    except cvs.CVSError as e2:
        if e2.status == 6:
            pytest.skip(f"nvJPEG is not available on this box: {e2}")
        raise
    ref = cv2.imdecode(enc, cv2.IMREAD_COLOR).reshape(-1)
    assert np.abs(g.astype(np.int16) - ref.astype(np.int16)).mean() < 2.0


def test_submit_jpeg_payload_is_the_oracles(cvs, oracle, monkeypatch):
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    w, h = 1920, 1080
    n = 3 * w * h
    j1, j2 = _jpeg("k1_f1.jpg"), _jpeg("k1_f2.jpg")
    g1 = oracle.jpeg_decode_bgr(j1).reshape(-1)
    g2 = oracle.jpeg_decode_bgr(j2).reshape(-1)
    s = cvs.Stream(w, h, g1)
    bufs = []
    for j in (j2, j1):
        jb = cvs.alloc_host(len(j) + 64)
        jb.array()[:len(j)] = np.frombuffer(j, dtype=np.uint8)
        bufs.append((jb, len(j)))
    dout, xout, pb = cvs.alloc_host(n + 32), cvs.alloc_host(4 * n + 32), (C.c_uint * 1)()
    oref = g1
    for rep, (cur, (jb, jl)) in enumerate(((g2, bufs[0]), (g1, bufs[1]), (g2, bufs[0]))):
        tk = s.submit_jpeg_raw(jb.ptr, jl, dout.ptr, None, "", C.addressof(pb), xout.ptr)
        s.wait(tk)
        opos, oxs, odiff, oref, _ = oracle.diff_compact(cur, oref, 20)
        if rep == 0:
            assert opos == 369350  # K1, from the bitstreams alone
        assert pb[0] == opos, f"submission {rep}"
        assert np.array_equal(dout.array()[:opos], odiff) and np.array_equal(xout.array(np.int32)[:opos], oxs)
        assert np.array_equal(s.reference(), oref)
    # several tickets in flight: the decoder's scratch is shared by the slots of a handle
    tks = []
    outs = [(cvs.alloc_host(n + 32), cvs.alloc_host(4 * n + 32), (C.c_uint * 1)()) for _ in range(4)]
    for q in range(4):
        jb, jl = bufs[q & 1]
        d, x, pp = outs[q]
        tks.append(s.submit_jpeg_raw(jb.ptr, jl, d.ptr, None, "", C.addressof(pp), x.ptr))
    for q in range(4):
        s.wait(tks[q])
        cur = g2 if (q & 1) == 0 else g1
        opos, oxs, odiff, oref, _ = oracle.diff_compact(cur, oref, 20)
        d, x, pp = outs[q]
        assert pp[0] == opos and np.array_equal(d.array()[:opos], odiff) and np.array_equal(x.array(np.int32)[:opos], oxs)
    s.close()


def test_jpeg_in_compact_wire_format_out(cvs, oracle, monkeypatch):
    """Both opt-ins together (cvs_submit_jpeg_wire): the CVW1 frame holds the oracle's payload of the oracle's pixels."""
    from cudavideostream_b200 import wire
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    w, h = 1920, 1080
    n = 3 * w * h
    j1, j2 = _jpeg("k1_f1.jpg"), _jpeg("k1_f2.jpg")
    g1, g2 = oracle.jpeg_decode_bgr(j1).reshape(-1), oracle.jpeg_decode_bgr(j2).reshape(-1)
    s = cvs.Stream(w, h, g1)
    wout = cvs.alloc_host(wire.bound(w, h) + 64)
    oref = g1
    for cur, j in ((g2, j2), (g1, j1)):
        jb = cvs.alloc_host(len(j) + 64)
        jb.array()[:len(j)] = np.frombuffer(j, dtype=np.uint8)
        s.wait(s.submit_jpeg_wire_raw(jb.ptr, len(j), wout.ptr, None, ""))
        opos, oxs, odiff, oref, _ = oracle.diff_compact(cur, oref, 20)
        enc = wout.array()
        pos, xs, diff = wire.parse(enc[:wire.size_of(enc)])
        assert pos == opos and np.array_equal(xs, oxs) and np.array_equal(diff, odiff)
        assert wire.size_of(enc) < (4 + 5 * opos) // 2
    assert np.array_equal(s.reference(), oref)
    s.close()


def test_camera_frame_3840x2160(cvs, oracle, monkeypatch):
    """A 4K frame (13 k subsequences, four times the blocks): still OpenCV's pixels."""
    cv2 = pytest.importorskip("cv2")
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    src = cv2.imdecode(np.frombuffer(_jpeg("k1_f2.jpg"), np.uint8), cv2.IMREAD_COLOR)
    big = cv2.resize(src, (3840, 2160), interpolation=cv2.INTER_CUBIC)
    ok, enc = cv2.imencode(".jpg", big, [cv2.IMWRITE_JPEG_QUALITY, 93])
    assert ok
    ref = cv2.imdecode(enc, cv2.IMREAD_COLOR).reshape(-1)
    g = _decode(cvs, enc.tobytes(), 3840, 2160)
    bad = np.flatnonzero(g != ref)
    assert bad.size == 0, f"{bad.size} bytes differ, first at {bad[:5]}"


def test_rejects_other_sizes_garbage_and_damaged_streams(cvs, monkeypatch):
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    w, h = 640, 360
    s = cvs.Stream(w, h, np.zeros(3 * w * h, dtype=np.uint8))
    j = _jpeg("k1_f1.jpg")  # 1920x1080
    jb = cvs.alloc_host(len(j) + 64)
    jb.array()[:len(j)] = np.frombuffer(j, dtype=np.uint8)
    n = 3 * w * h
    dout, xout, pb = cvs.alloc_host(n + 32), cvs.alloc_host(4 * n + 32), (C.c_uint * 1)()
    with pytest.raises(cvs.CVSError) as e:
        s.submit_jpeg_raw(jb.ptr, len(j), dout.ptr, None, "", C.addressof(pb), xout.ptr)
    assert e.value.status == 1  # CVS_ERR_INVALID: the JPEG is 1920x1080, the stream 640x360
    jb.array()[:64] = 0x55
    with pytest.raises(cvs.CVSError):
        s.submit_jpeg_raw(jb.ptr, 64, dout.ptr, None, "", C.addressof(pb), xout.ptr)
    # the handle is still usable afterwards
    f = np.full(n, 77, dtype=np.uint8)
    pos, xs, diff, _ = s.exec(f)
    assert pos == n
    s.close()
    # a truncated frame through the pipelined call: cvs_wait reports it (the stream holds fewer blocks than the image)
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    name = "q50_420_641x359"
    w, h = (int(v) for v in z[name + "/wh"])
    n = 3 * w * h
    good = z[name + "/jpg"].tobytes()
    cutj = good[:len(good) // 2] + b"\xff\xd9"
    s = cvs.Stream(w, h, np.zeros(n, dtype=np.uint8))
    jb = cvs.alloc_host(len(good) + 64)
    dout, xout, pb = cvs.alloc_host(n + 32), cvs.alloc_host(4 * n + 32), (C.c_uint * 1)()
    jb.array()[:len(cutj)] = np.frombuffer(cutj, dtype=np.uint8)
    tk = s.submit_jpeg_raw(jb.ptr, len(cutj), dout.ptr, None, "", C.addressof(pb), xout.ptr)
    with pytest.raises(cvs.CVSError) as e:
        s.wait(tk)
    assert e.value.status == 1
    s.reset(np.zeros(n, dtype=np.uint8))  # the damaged frame's pixels went into the reference: re-seed it
    jb.array()[:len(good)] = np.frombuffer(good, dtype=np.uint8)
    s.wait(s.submit_jpeg_raw(jb.ptr, len(good), dout.ptr, None, "", C.addressof(pb), xout.ptr))
    assert hashlib.sha256(s.reference().tobytes()).digest() != b"" and pb[0] > 0
    s.close()
    # a damaged entropy-coded segment: never a crash, a hang or a write outside the frame; either an error status or a picture
    rng = np.random.default_rng(5)
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    name = "q50_420_641x359"
    w, h = (int(v) for v in z[name + "/wh"])
    for trial in range(6):
        jpg = bytearray(z[name + "/jpg"].tobytes())
        sos = jpg.index(b"\xff\xda")
        lo = sos + 14
        if trial < 3:
            for k in rng.integers(lo, len(jpg) - 2, size=8):
                jpg[k] = int(rng.integers(0, 255))
                if jpg[k] == 0xFF:
                    jpg[k] = 0x7F
        else:
            cut = int(rng.integers(lo + 100, len(jpg) - 100))
            jpg = jpg[:cut] + b"\xff\xd9"
        try:
            g = _decode(cvs, bytes(jpg), w, h)
            assert g.size == 3 * w * h
        except cvs.CVSError as e:
            assert e.status == 1


def test_damaged_headers_are_contained(cvs, oracle, monkeypatch):
    """A frame whose header was damaged on the way (wrong Huffman / quantisation tables, other sampling factors, a
    restart interval that is not there) either is refused or decodes to SOME picture: no crash, no hang, nothing written
    outside the frame, and the library decodes the next good frame exactly (the CPU side of the same damage:
    tests/test_jpeg_parse_fuzz.py)."""
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    rng = np.random.default_rng(23)
    outcomes = {"decoded": 0, "refused": 0}
    for name in ("q100_420_37x29", "q90_gray_123x77", "q85_444_161x97", "q95_420_rst4_256x144"):
        good = z[name + "/jpg"].tobytes()
        w, h = (int(v) for v in z[name + "/wh"])
        sos = good.index(b"\xff\xda")
        for trial in range(40):
            m = bytearray(good)
            for k in rng.integers(2, sos + 10, size=int(rng.integers(1, 4))):
                m[k] = int(rng.integers(0, 256))
            try:
                g = _decode(cvs, bytes(m), w, h)  # checks the guard band behind the frame itself
                assert g.size == 3 * w * h
                outcomes["decoded"] += 1
            except cvs.CVSError as e:
                assert e.status == 1, f"{name} trial {trial}: status {e.status}"
                outcomes["refused"] += 1
        g = _decode(cvs, good, w, h)
        assert np.array_equal(g, oracle.jpeg_decode_bgr(good).reshape(-1)), f"{name}: good frame after the damaged ones"
    assert outcomes["decoded"] > 20 and outcomes["refused"] > 20, outcomes


@pytest.mark.parametrize("name,dc_len,ac_len", [("q95_420_200x150", 11, 10), ("q90_gray_123x77", 10, 12), ("q92_420_1280x720", 16, 16)])
def test_uniform_long_codes(cvs, oracle, monkeypatch, name, dc_len, ac_len):
    """Worst case for the decoder's look-up tables: every code dc_len / ac_len bits long (tests/util.transcode_huffman)."""
    from util import transcode_huffman
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    z = np.load(os.path.join(GOLDEN, "jpeg_cases.npz"))
    jpg = z[name + "/jpg"].tobytes()
    w, h = (int(v) for v in z[name + "/wh"])
    info, coef = oracle.jpeg_info(jpg), oracle.jpeg_coefficients(jpg)
    t = transcode_huffman(jpg, coef, info["hs"], info["vs"], info["ncomp"], dc_len, ac_len)
    g = _decode(cvs, t, w, h)
    assert hashlib.sha256(g.tobytes()).digest() == z[name + "/sha"].tobytes()


def test_random_streams_against_opencv(cvs, oracle, monkeypatch):
    """Random pictures (smooth, noisy, saturated), sizes, qualities 3..100, samplings, optimised Huffman tables (unusual code
    length distributions: the second-level tables and the canonical slow path of the decoder's look-up) and restart
    intervals, encoded by OpenCV on the spot: the GPU decoder, the oracle and cv2.imdecode must agree bit for bit."""
    cv2 = pytest.importorskip("cv2")
    monkeypatch.setenv("CVS_JPEG_DECODER", "own")
    rng = np.random.default_rng(20261018)
    samplings = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    for case in range(28):
        w, h = int(rng.integers(1, 420)), int(rng.integers(1, 300))
        kind = case % 4
        yy, xx = np.mgrid[0:h, 0:w]
        if kind == 0:      # smooth
            img = np.stack([(xx * 3 + yy) % 256, (yy * 5) % 256, (xx + yy * 2) % 256], axis=2)
        elif kind == 1:    # noise: large coefficients, long codes
            img = rng.integers(0, 256, size=(h, w, 3))
        elif kind == 2:    # saturated rectangles on noise
            img = rng.integers(100, 156, size=(h, w, 3))
            for _ in range(5):
                x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
                img[y0:y0 + h // 3 + 1, x0:x0 + w // 3 + 1] = rng.integers(0, 2, size=3) * 255
        else:              # flat
            img = np.full((h, w, 3), int(rng.integers(0, 256)))
        img = img.astype(np.uint8)
        gray = case % 7 == 3
        params = [cv2.IMWRITE_JPEG_QUALITY, int(rng.integers(3, 101))]
        if not gray:
            params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samplings[int(rng.integers(0, 3))]]
        if rng.integers(0, 2):
            params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
        if case % 3 == 2:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(rng.integers(1, 9))]
        ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(img[:, :, 1]) if gray else img, params)
        assert ok
        ref = cv2.imdecode(enc, cv2.IMREAD_COLOR).reshape(-1)
        assert np.array_equal(oracle.jpeg_decode_bgr(enc.tobytes()).reshape(-1), ref), f"case {case}: oracle vs cv2 ({w}x{h}, {params})"
        g = _decode(cvs, enc.tobytes(), w, h)
        bad = np.flatnonzero(g != ref)
        assert bad.size == 0, f"case {case} ({w}x{h}, {params}): {bad.size} bytes differ, first at {bad[:5]}"


def test_many_tickets_on_several_streams(cvs):
    """scripts/jpeg_soak.py in small: 3 streams x 120 tickets through cvs_submit_jpeg with four tickets in flight each (the
    tickets of a stream alternate between two decode streams with their own scratch): every ticket delivers the same
    payload on every stream, and the payloads repeat with the two camera frames."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "jpeg_soak.py"), "3", "120"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=170, env=dict(os.environ, CVS_JPEG_DECODER="own"))
    assert r.returncode == 0, r.stdout.decode()[-800:]
