"""Generates tests/golden/jpeg_cases.npz: small baseline JPEGs (other qualities, chroma samplings, sizes that are not
multiples of the MCU, restart intervals, optimised Huffman tables) together with the SHA-256 of what cv2.imdecode
(OpenCV's libjpeg-turbo, the reference's decoder: server/src/threads.cpp:32-41) makes of them.  Run in the build
container (needs cv2):

    python tests/golden/make_jpeg_cases.py
"""
import hashlib
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def picture(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x + y) * 3) % 256], axis=2).astype(np.int32)
    img += rng.integers(-40, 41, size=img.shape)
    # a few saturated rectangles: hard edges exercise the range limiter and large AC coefficients
    for _ in range(6):
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        img[y0:y0 + h // 5 + 1, x0:x0 + w // 5 + 1] = rng.integers(0, 2, size=3) * 255
    return np.clip(img, 0, 255).astype(np.uint8)


CASES = [
    # name, w, h, params
    ("q95_420_200x150", 200, 150, [cv2.IMWRITE_JPEG_QUALITY, 95]),
    ("q50_420_641x359", 641, 359, [cv2.IMWRITE_JPEG_QUALITY, 50]),
    ("q100_420_37x29", 37, 29, [cv2.IMWRITE_JPEG_QUALITY, 100]),
    ("q10_420_128x128", 128, 128, [cv2.IMWRITE_JPEG_QUALITY, 10]),
    ("q90_422_320x240", 320, 240, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]),
    ("q75_422_333x77", 333, 77, [cv2.IMWRITE_JPEG_QUALITY, 75, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]),
    ("q85_444_161x97", 161, 97, [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]),
    ("q95_420_rst4_256x144", 256, 144, [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_RST_INTERVAL, 4]),
    ("q80_422_rst1_100x60", 100, 60, [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
                                    cv2.IMWRITE_JPEG_RST_INTERVAL, 1]),
    ("q90_420_opt_300x200", 300, 200, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_OPTIMIZE, 1]),
    ("q95_420_8x8", 8, 8, [cv2.IMWRITE_JPEG_QUALITY, 95]),
    ("q95_420_3x5", 3, 5, [cv2.IMWRITE_JPEG_QUALITY, 95]),
    ("q92_420_1280x720", 1280, 720, [cv2.IMWRITE_JPEG_QUALITY, 92]),
    ("q90_gray_123x77", 123, 77, [cv2.IMWRITE_JPEG_QUALITY, 90]),            # one component (encoded from a gray image)
    ("q85_422_rst8_640x480", 640, 480, [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
                                       cv2.IMWRITE_JPEG_RST_INTERVAL, 8]),
    ("q30_444_rst3_97x55", 97, 55, [cv2.IMWRITE_JPEG_QUALITY, 30, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
                                   cv2.IMWRITE_JPEG_RST_INTERVAL, 3]),
]


def main():
    out = {}
    for i, (name, w, h, params) in enumerate(CASES):
        pic = picture(w, h, 1000 + i)
        if "_gray_" in name:
            pic = np.ascontiguousarray(pic[:, :, 1])
        ok, enc = cv2.imencode(".jpg", pic, params)
        assert ok
        dec = cv2.imdecode(enc, cv2.IMREAD_COLOR)
        assert dec.shape == (h, w, 3)
        out[name + "/jpg"] = enc.reshape(-1)
        out[name + "/sha"] = np.frombuffer(hashlib.sha256(dec.tobytes()).digest(), dtype=np.uint8)
        out[name + "/wh"] = np.array([w, h], dtype=np.int32)
        print(name, enc.size, "bytes")
    np.savez_compressed(os.path.join(HERE, "jpeg_cases.npz"), **out)


if __name__ == "__main__":
    main()
