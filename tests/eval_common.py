"""Shared helpers of the evaluation-protocol tests: golden images and the oracle-side inputs."""
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_png_rgb(path):
    """Lossless PNG -> HWC uint8 RGB (Pillow-free: cv2 is part of the image; falls back to a tiny stdlib reader)."""
    import cv2
    im = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    return np.ascontiguousarray(im[:, :, ::-1])


def kat():
    return json.loads((GOLDEN / "eval_kat.json").read_text())


def cbsd68_images(k):
    return [load_png_rgb(GOLDEN / "images" / f"cbsd68_{e['name']}.png") for e in k["denoise"]]


def set5_images(k):
    return [load_png_rgb(GOLDEN / "images" / f"set5_{e['name']}.png") for e in k["sisr"]]


def checksum_close(arr, ck, stride, rtol):
    a = np.asarray(arr, dtype=np.float64)
    assert abs(a.sum() - ck["sum"]) <= rtol * max(1.0, abs(ck["sum"])), (a.sum(), ck["sum"])
    assert abs((a * a).sum() - ck["sumsq"]) <= rtol * max(1.0, ck["sumsq"]), ((a * a).sum(), ck["sumsq"])
    s = a.flatten()[::stride][:64]
    np.testing.assert_allclose(s, np.array(ck["sample"]), rtol=0, atol=rtol * 10)
