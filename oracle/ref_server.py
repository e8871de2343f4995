"""Runs the binaries of oracle/_ref/ (the reference's unmodified server.cpp behind the file-fed ThreadsCore stub,
see oracle/build_ref.py and tests/host/threads_stub.cpp).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
NGLYPHS = 22  # len(CHARS_STR), server/include/common.h:13


def binary(name: str):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def run(name: str, width: int, height: int, base: np.ndarray, frames: np.ndarray, *, glyphs=None, gw: int = 0, gh: int = 0,
        want_show: bool = False, env=None, timeout: float = 300.0, want_times: bool = False):
    """Feeds `frames` ([T, N] uint8) to oracle/_ref/<name>.  Returns a list with one dict per frame:
    {"pos", "xs", "diff"} (GPU branch) or {"data"} (CPU branch), plus "show" when want_show; and, with want_times,
    the per-frame nanoseconds between readCap() returning and writeShow() being called."""
    exe = binary(name)
    if exe is None:
        raise FileNotFoundError(f"oracle/_ref/{name} has not been built (python oracle/build_ref.py)")
    n = 3 * width * height
    frames = np.ascontiguousarray(frames, dtype=np.uint8).reshape(-1, n)
    base = np.ascontiguousarray(base, dtype=np.uint8).reshape(-1)
    assert base.size == n
    g = np.zeros(0, np.uint8) if glyphs is None else np.ascontiguousarray(glyphs, dtype=np.uint8).reshape(-1)
    assert g.size == NGLYPHS * 3 * gw * gh
    with tempfile.TemporaryDirectory() as td:
        fin, fout, ftimes = os.path.join(td, "in.bin"), os.path.join(td, "out.bin"), os.path.join(td, "times.txt")
        with open(fin, "wb") as f:
            f.write(struct.pack("<5i", width, height, frames.shape[0], gw, gh))
            f.write(g.tobytes())
            f.write(base.tobytes())
            f.write(frames.tobytes())
        e = dict(os.environ)
        e.update({"CVS_STUB_IN": fin, "CVS_STUB_OUT": fout, "CVS_STUB_SHOW": "1" if want_show else "0"})
        if want_times:
            e["CVS_STUB_TIMES"] = ftimes
        e.update(env or {})
        r = subprocess.run([exe], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"{name} exited with {r.returncode}: {r.stderr.decode(errors='replace')[-800:]}")
        raw = open(fout, "rb").read()
        times = [int(x) for x in open(ftimes).read().split()] if want_times else None
    out, off = [], 0
    for _ in range(frames.shape[0]):
        pos, kind = struct.unpack_from("<II", raw, off)
        off += 8
        rec = {}
        if kind == 0:
            rec["pos"] = pos
            rec["xs"] = np.frombuffer(raw, dtype=np.int32, count=pos, offset=off).copy()
            off += 4 * pos
            rec["diff"] = np.frombuffer(raw, dtype=np.uint8, count=pos, offset=off).copy()
            off += pos
        else:
            rec["data"] = np.frombuffer(raw, dtype=np.uint8, count=n, offset=off).copy()
            off += n
        if want_show:
            rec["show"] = np.frombuffer(raw, dtype=np.uint8, count=n, offset=off).copy()
            off += n
        out.append(rec)
    assert off == len(raw), "trailing bytes in the stub's output"
    return (out, times) if want_times else out
