"""Shared helpers for the parity tests."""
import numpy as np

CHARS_STR = "0123456789BFPSWbkps :/"  # server/include/common.h:13


def glyph_atlas(gw=7, gh=5, seed=3):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(len(CHARS_STR), gh, gw, 3), dtype=np.uint8)


def random_sequence(width, height, nframes, density, seed=0, drift=True):
    """numpy-only frame generator for small shapes: base + frames with `density` of bytes changed by a
    super-threshold delta and (optionally) sub-threshold drift on the rest."""
    rng = np.random.default_rng(seed)
    n = 3 * width * height
    base = rng.integers(0, 256, size=n, dtype=np.uint8)
    frames = np.empty((nframes, n), dtype=np.uint8)
    prev = base.astype(np.int64)
    for t in range(nframes):
        big = rng.random(n) < density
        delta = rng.integers(21, 120, size=n) * rng.choice([-1, 1], size=n)
        small = rng.integers(-3, 4, size=n) if drift else np.zeros(n, dtype=np.int64)
        cur = np.where(big, prev + delta, prev + small)
        cur = np.where((cur < 0) | (cur > 255), prev - (cur - prev), cur)
        cur = np.clip(cur, 0, 255)
        frames[t] = cur.astype(np.uint8)
        prev = cur
    return base, frames


def run_oracle_sequence(oracle, width, height, base, frames, **kw):
    """Runs OracleCore frame by frame.  Returns list of (pos, xs, diff, show) and the final reference."""
    core = oracle.OracleCore(width, height, base, **kw)
    text = kw.pop("text", "") if "text" in kw else ""
    out = []
    for f in frames:
        pos, xs, diff, show, _ = core.exec_core(f, text)
        out.append((pos, xs, diff, show))
    ref = core.reference()
    core.close()
    return out, ref


def strip_dht(jpg: bytes) -> bytes:
    """The JPEG without its DHT segments (MJPG as cameras send it: the standard Huffman tables are implied)."""
    out, i = bytearray(jpg[:2]), 2
    while i < len(jpg):
        m, ln = jpg[i + 1], (jpg[i + 2] << 8) | jpg[i + 3]
        if m == 0xDA:
            out += jpg[i:]
            break
        if m != 0xC4:
            out += jpg[i:i + 2 + ln]
        i += 2 + ln
    return bytes(out)
