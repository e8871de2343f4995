"""Generates the committed golden fixtures.  Run in the build container (needs /root/reference + cv2):

    python tests/golden/make_golden.py

k1_f1_f2.json : facts about the reference's own fixture pair tests/noise_filter_benchmark/f{1,2}.jpg
                (REPORT/report.tex:2594 states 369,350 changed bytes) plus digests of the oracle's payload on
                it, so the GPU box -- which has no /root/reference -- can still check the CUDA path against
                the real webcam frames through the 128x72 crop stored in k1_crop.npz.
small_payload.json : oracle payloads for tiny hand-checkable frame pairs.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402

REF = "/root/reference/tests/noise_filter_benchmark"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import cv2
    orc.build()
    a = cv2.imread(os.path.join(REF, "f1.jpg"))
    b = cv2.imread(os.path.join(REF, "f2.jpg"))
    pos, xs, diff, ref, _ = orc.diff_compact(b, a, 20)
    changed_px = int(np.any(np.abs(a.astype(int) - b.astype(int)) > 20, axis=2).sum())
    rec = {"source": "tests/noise_filter_benchmark/f1.jpg,f2.jpg (reference)", "threshold": 20,
           "total_bytes": int(a.size), "changed_bytes": int(pos), "changed_pixels": changed_px,
           "report_line": "REPORT/report.tex:2594", "sha256_f1": sha(a), "sha256_f2": sha(b),
           "sha256_xs": sha(xs), "sha256_diff": sha(diff), "sha256_new_reference": sha(ref)}
    # a crop of the real frames (the busiest 128x72 window) so that real-camera statistics travel
    best, by, bx = -1, 0, 0
    ch = np.abs(a.astype(int) - b.astype(int)) > 20
    for y in range(0, 1080 - 72, 36):
        for x in range(0, 1920 - 128, 64):
            c = int(ch[y:y + 72, x:x + 128].sum())
            if c > best:
                best, by, bx = c, y, x
    ca, cb = np.ascontiguousarray(a[by:by + 72, bx:bx + 128]), np.ascontiguousarray(b[by:by + 72, bx:bx + 128])
    cpos, cxs, cdiff, cref, _ = orc.diff_compact(cb, ca, 20)
    np.savez_compressed(os.path.join(HERE, "k1_crop.npz"), f1=ca, f2=cb, xs=cxs, diff=cdiff, new_reference=cref)
    rec["crop"] = {"y": by, "x": bx, "h": 72, "w": 128, "changed_bytes": int(cpos)}
    with open(os.path.join(HERE, "k1_f1_f2.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(rec)


if __name__ == "__main__":
    main()
