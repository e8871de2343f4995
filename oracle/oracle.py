"""ctypes binding of the CPU oracle (oracle/cvs_oracle.c).  TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (cudavideostream_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_O0.so with the Makefile next to this file."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, "cvs_oracle.c"), os.path.join(_HERE, "jpeg_oracle.c"), os.path.join(_HERE, "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])


def _load(name: str = "liboracle.so") -> C.CDLL:
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    lib.orc_diff_compact.restype = C.c_uint
    lib.orc_diff_compact.argtypes = [_u8p, _u8p, _i32p, C.c_int, C.c_int]
    lib.orc_client_apply.argtypes = [_u8p, _i32p, _u8p, C.c_uint]
    lib.orc_count_difference.restype = C.c_int
    lib.orc_count_difference.argtypes = [_u8p, _u8p, C.c_int, C.c_int]
    lib.orc_gray_avg3.argtypes = [_u8p, C.c_int]
    lib.orc_gray_avg1.argtypes = [_u8p, _u8p, C.c_int, C.c_int]
    lib.orc_gray_weighted1.argtypes = [_u8p, _u8p, C.c_int, C.c_int]
    lib.orc_gray_weighted3.argtypes = [_u8p, _u8p, C.c_int]
    lib.orc_histogram3.argtypes = [_u8p, C.c_int, _i32p]
    lib.orc_histogram1.argtypes = [_u8p, C.c_int, _i32p]
    lib.orc_threshold_twomax.restype = C.c_int
    lib.orc_threshold_twomax.argtypes = [_i32p, C.c_int, C.c_int]
    lib.orc_binarize.argtypes = [_u8p, C.c_int, C.c_int]
    lib.orc_heat_pixel.argtypes = [C.c_int, _i32p, _i32p, _i32p]
    lib.orc_heat_map.argtypes = [_u8p, _u8p, _u8p, C.c_int, C.c_int]
    lib.orc_red_map.argtypes = [_u8p, _u8p, _u8p, C.c_int, C.c_int, C.c_int]
    lib.orc_red_overlap_from_xs.argtypes = [_u8p, _i32p, C.c_uint]
    lib.orc_gaussian_kernel.argtypes = [_f32p, C.c_int, C.c_float]
    lib.orc_mean_kernel.argtypes = [_f32p, C.c_int]
    lib.orc_noise_filter.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, _f32p]
    lib.orc_text_overlay.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
    lib.orc_jpeg_info.argtypes = [_u8p, C.c_size_t, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, C.POINTER(C.c_long)]
    lib.orc_jpeg_coefficients.argtypes = [_u8p, C.c_size_t, C.POINTER(C.c_int16), C.c_long]
    lib.orc_jpeg_decode_bgr.argtypes = [_u8p, C.c_size_t, _u8p, C.c_int, C.c_int]
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _u8p,
                               _u8p, C.c_int, C.c_int, C.c_char_p]
    lib.orc_destroy.argtypes = [C.c_void_p]
    lib.orc_reference.restype = _u8p
    lib.orc_reference.argtypes = [C.c_void_p]
    lib.orc_exec_core.argtypes = [C.c_void_p, _u8p, _u8p, C.c_char_p, C.POINTER(C.c_uint), _i32p]
    lib.orc_synth_base.argtypes = [_u8p, C.c_int, C.c_int, C.c_uint64]
    lib.orc_synth_next.argtypes = [_u8p, _u8p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
    lib.orc_bench_diff_compact.restype = C.c_double
    lib.orc_bench_diff_compact.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int,
                                           C.POINTER(C.c_ulonglong)]
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(_u8p)


def _pi(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_i32p)


def _pf(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_f32p)


# ------------------------------------------------------------------------------------------
# functional wrappers (inputs are never modified; outputs are fresh arrays)
# ------------------------------------------------------------------------------------------
def diff_compact(cur: np.ndarray, prev: np.ndarray, thr: int = 20):
    """A1.  Returns (pos, xs[pos], diff[pos], new_reference[N], frame_after[N])."""
    frame = np.ascontiguousarray(cur, dtype=np.uint8).reshape(-1).copy()
    ref = np.ascontiguousarray(prev, dtype=np.uint8).reshape(-1).copy()
    xs = np.empty(max(frame.size, 1), dtype=np.int32)
    pos = lib().orc_diff_compact(_p8(frame), _p8(ref), _pi(xs), frame.size, thr)
    return pos, xs[:pos].copy(), frame[:pos].copy(), ref, frame


def client_apply(frame: np.ndarray, xs: np.ndarray, diff: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(frame, dtype=np.uint8).reshape(-1).copy()
    xs = np.ascontiguousarray(xs, dtype=np.int32)
    diff = np.ascontiguousarray(diff, dtype=np.uint8)
    lib().orc_client_apply(_p8(out), _pi(xs) if xs.size else None, _p8(diff) if diff.size else None, xs.size)
    return out


def count_difference(a: np.ndarray, b: np.ndarray, thr: int = 20) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
    return lib().orc_count_difference(_p8(a), _p8(b), a.size, thr)


def gray_avg3(frame: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(frame, dtype=np.uint8).reshape(-1).copy()
    lib().orc_gray_avg3(_p8(out), out.size)
    return out


def gray_avg1(frame: np.ndarray, width: int, height: int) -> np.ndarray:
    f = np.ascontiguousarray(frame, dtype=np.uint8).reshape(-1)
    out = np.empty(width * height, dtype=np.uint8)
    lib().orc_gray_avg1(_p8(f), _p8(out), width, height)
    return out


def gray_weighted1(frame: np.ndarray, width: int, height: int) -> np.ndarray:
    f = np.ascontiguousarray(frame, dtype=np.uint8).reshape(-1)
    out = np.empty(width * height, dtype=np.uint8)
    lib().orc_gray_weighted1(_p8(f), _p8(out), width, height)
    return out


def gray_weighted3(frame: np.ndarray) -> np.ndarray:
    f = np.ascontiguousarray(frame, dtype=np.uint8).reshape(-1)
    out = np.empty(f.size, dtype=np.uint8)
    lib().orc_gray_weighted3(_p8(f), _p8(out), f.size)
    return out


def histogram3(gray3: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(gray3, dtype=np.uint8).reshape(-1)
    h = np.zeros(256, dtype=np.int32)
    lib().orc_histogram3(_p8(g), g.size, _pi(h))
    return h


def histogram1(gray1: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(gray1, dtype=np.uint8).reshape(-1)
    h = np.zeros(256, dtype=np.int32)
    lib().orc_histogram1(_p8(g), g.size, _pi(h))
    return h


def threshold_twomax(hist: np.ndarray, clamp_lo: int = 50, clamp_hi: int = 200) -> int:
    h = np.ascontiguousarray(hist, dtype=np.int32)
    return lib().orc_threshold_twomax(_pi(h), clamp_lo, clamp_hi)


def binarize(data: np.ndarray, thr: int) -> np.ndarray:
    out = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1).copy()
    lib().orc_binarize(_p8(out), out.size, thr)
    return out


def heat_pixel(d: int):
    r, g, b = C.c_int(), C.c_int(), C.c_int()
    lib().orc_heat_pixel(d, C.byref(r), C.byref(g), C.byref(b))
    return r.value, g.value, b.value


def heat_map(prev: np.ndarray, cur: np.ndarray, width: int, height: int) -> np.ndarray:
    a = np.ascontiguousarray(prev, dtype=np.uint8).reshape(-1)
    b = np.ascontiguousarray(cur, dtype=np.uint8).reshape(-1)
    out = np.empty(a.size, dtype=np.uint8)
    lib().orc_heat_map(_p8(a), _p8(b), _p8(out), width, height)
    return out


def red_map(prev: np.ndarray, cur: np.ndarray, width: int, height: int, thr: int = 20) -> np.ndarray:
    a = np.ascontiguousarray(prev, dtype=np.uint8).reshape(-1)
    b = np.ascontiguousarray(cur, dtype=np.uint8).reshape(-1)
    out = np.empty(a.size, dtype=np.uint8)
    lib().orc_red_map(_p8(a), _p8(b), _p8(out), width, height, thr)
    return out


def red_overlap_from_xs(base: np.ndarray, xs: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(base, dtype=np.uint8).reshape(-1).copy()
    xs = np.ascontiguousarray(xs, dtype=np.int32)
    if xs.size:
        lib().orc_red_overlap_from_xs(_p8(out), _pi(xs), xs.size)
    return out


def gaussian_kernel(K: int, sigma: float) -> np.ndarray:
    k = np.empty(K * K, dtype=np.float32)
    lib().orc_gaussian_kernel(_pf(k), K, sigma)
    return k


def mean_kernel(K: int) -> np.ndarray:
    k = np.empty(K * K, dtype=np.float32)
    lib().orc_mean_kernel(_pf(k), K)
    return k


def noise_filter(image: np.ndarray, width: int, height: int, K: int, k: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(image, dtype=np.uint8).reshape(-1)
    k = np.ascontiguousarray(k, dtype=np.float32).reshape(-1)
    out = np.empty(img.size, dtype=np.uint8)
    lib().orc_noise_filter(_p8(img), _p8(out), width, height, K, _pf(k))
    return out


def jpeg_info(jpg) -> dict:
    """Geometry of a baseline JPEG (oracle/jpeg_oracle.c)."""
    b = np.ascontiguousarray(np.frombuffer(bytes(jpg), dtype=np.uint8))
    v = [C.c_int() for _ in range(6)]
    nb = C.c_long()
    e = lib().orc_jpeg_info(_p8(b), b.size, *[C.byref(x) for x in v], C.byref(nb))
    if e:
        raise ValueError(f"jpeg oracle: not a supported baseline JPEG ({e})")
    return dict(width=v[0].value, height=v[1].value, hs=v[2].value, vs=v[3].value, ncomp=v[4].value,
                restart_interval=v[5].value, nblocks=nb.value)


def jpeg_coefficients(jpg) -> np.ndarray:
    """Quantised DCT coefficients [nblocks, 64] in scan order, natural order inside a block, DC absolute."""
    info = jpeg_info(jpg)
    b = np.ascontiguousarray(np.frombuffer(bytes(jpg), dtype=np.uint8))
    out = np.zeros((info["nblocks"], 64), dtype=np.int16)
    e = lib().orc_jpeg_coefficients(_p8(b), b.size, out.ctypes.data_as(C.POINTER(C.c_int16)), info["nblocks"])
    if e:
        raise ValueError(f"jpeg oracle: entropy decode failed ({e})")
    return out


def jpeg_decode_bgr(jpg) -> np.ndarray:
    """What cv2.imread / VideoCapture (libjpeg-turbo defaults) make of a baseline JPEG: [h, w, 3] BGR."""
    info = jpeg_info(jpg)
    b = np.ascontiguousarray(np.frombuffer(bytes(jpg), dtype=np.uint8))
    out = np.zeros((info["height"], info["width"], 3), dtype=np.uint8)
    e = lib().orc_jpeg_decode_bgr(_p8(b), b.size, _p8(out), info["width"], info["height"])
    if e:
        raise ValueError(f"jpeg oracle: decode failed ({e})")
    return out


def text_overlay(frame: np.ndarray, width: int, height: int, glyphs: np.ndarray, gw: int, gh: int,
                 chars: str, text: str) -> np.ndarray:
    out = np.ascontiguousarray(frame, dtype=np.uint8).reshape(-1).copy()
    g = np.ascontiguousarray(glyphs, dtype=np.uint8).reshape(-1)
    lib().orc_text_overlay(_p8(out), width, height, _p8(g), gw, gh, chars.encode(), text.encode())
    return out


class OracleCore:
    """CPU twin of diff::cuda::CUDACore (server/include/kernels.cuh:13-43) built from the CPU loops."""

    def __init__(self, width, height, base_frame, thr=20, mode=0, noise_filter=0, K=3, k=None,
                 glyphs=None, gw=0, gh=0, chars=""):
        self.width, self.height, self.total = width, height, 3 * width * height
        base = np.ascontiguousarray(base_frame, dtype=np.uint8).reshape(-1)
        assert base.size == self.total
        self._k = None if k is None else np.ascontiguousarray(k, dtype=np.float32).reshape(-1)
        self._glyphs = None if glyphs is None else np.ascontiguousarray(glyphs, dtype=np.uint8).reshape(-1)
        self._h = lib().orc_create(width, height, thr, mode, noise_filter, K,
                                   _pf(self._k) if self._k is not None else None, _p8(base),
                                   _p8(self._glyphs) if self._glyphs is not None else None,
                                   gw, gh, chars.encode())
        self.mode = mode

    def exec_core(self, frame: np.ndarray, text: str = ""):
        """Returns (pos, xs, diff, show or None, frame_after)."""
        f = np.ascontiguousarray(frame, dtype=np.uint8).reshape(-1).copy()
        show = np.zeros(max(self.total, 1), dtype=np.uint8)[: self.total] if self.mode else None
        xs = np.empty(max(self.total, 1), dtype=np.int32)
        pos = C.c_uint(0)
        lib().orc_exec_core(self._h, _p8(f), _p8(show) if show is not None and show.size else None,
                            text.encode(), C.byref(pos), _pi(xs))
        p = pos.value
        return p, xs[:p].copy(), f[:p].copy(), show, f

    def reference(self) -> np.ndarray:
        ptr = lib().orc_reference(self._h)
        return np.ctypeslib.as_array(ptr, shape=(max(self.total, 1),))[: self.total].copy()

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bench_diff_compact(frames: np.ndarray, base: np.ndarray, thr: int, iters_per_thread: int,
                       nthreads: int, o0: bool = False):
    """Times A1 on `nthreads` independent streams.  Returns (seconds, total_frames, sum_pos)."""
    L = _load("liboracle_O0.so") if o0 else lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    base = np.ascontiguousarray(base, dtype=np.uint8).reshape(-1)
    nring, total = frames.shape[0], base.size
    s = C.c_ulonglong(0)
    sec = L.orc_bench_diff_compact(_p8(frames.reshape(-1)), nring, total, thr, _p8(base),
                                   iters_per_thread, nthreads, C.byref(s))
    return sec, iters_per_thread * nthreads, s.value


def synth_sequence(width: int, height: int, nframes: int, density_ppm: int, seed: int):
    """C twin of cudavideostream_b200.synth.sequence (fast enough for full-size frames).  Returns (base, frames)."""
    n = 3 * width * height
    base = np.empty(n, dtype=np.uint8)
    lib().orc_synth_base(_p8(base), width, height, seed & 0xFFFFFFFFFFFFFFFF)
    frames = np.empty((nframes, n), dtype=np.uint8)
    prev = base
    for t in range(nframes):
        lib().orc_synth_next(_p8(prev), _p8(frames[t]), n, seed & 0xFFFFFFFFFFFFFFFF, t, density_ppm)
        prev = frames[t]
    return base, frames
