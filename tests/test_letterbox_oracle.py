"""Pins oracle/letterbox_oracle.py (integer restatement of LetterBox + cv2.resize INTER_LINEAR + the predictor's BGR->RGB /
HWC->CHW) to the REAL reference pipeline (fixtures: tests/golden/make_golden.py letterbox, cv2 4.13): bit-exact (CRC) for
downscales / identity, within 1 LSB for the two-axis upscales where this cv2 wheel's SIMD dispatch departs from the generic kernel."""
import os
import zlib

import numpy as np
import pytest
import torch

from _util import GOLD
from oracle import letterbox_oracle as L

GOLDEN = torch.load(os.path.join(GOLD, "letterbox.golden.pt"))


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"{c['h']}x{c['w']}")
def test_letterbox_oracle_matches_reference(case):
    h, w = case["h"], case["w"]
    img = np.random.default_rng(case["seed"]).integers(0, 256, (h, w, 3), dtype=np.uint8)
    out = L.preprocess_frame(img, (640, 640))
    assert list(out.shape) == case["shape"] and out.dtype == np.uint8
    upscale = min(640 / h, 640 / w) > 1.0
    if not upscale:
        assert zlib.crc32(out.tobytes()) == case["crc"]                      # bit-exact
    thumb = out.reshape(3, 40, 16, 40, 16).astype(np.float32).mean((2, 4))
    assert float(np.abs(thumb - case["thumb"].numpy()).max()) <= (0.25 if upscale else 0.0)   # 16x16 block means; 1-LSB pixels cluster


def test_letterbox_params_known_answers():
    """LetterBox.get_params augment.py:1742-1786: 480x640 -> no resize, 80 px of padding top and bottom; 1080x1920 -> 640x360."""
    assert L.letterbox_params((480, 640)) == ((640, 480), 80, 80, 0, 0)
    assert L.letterbox_params((1080, 1920)) == ((640, 360), 140, 140, 0, 0)
    assert L.letterbox_params((333, 500)) == ((640, 426), 107, 107, 0, 0)
    assert L.letterbox_params((100, 37)) == ((237, 640), 0, 0, 201, 202)


def test_two_axis_upscale_is_within_one_lsb_of_cv2():
    cv2 = pytest.importorskip("cv2")
    img = np.random.default_rng(5).integers(0, 256, (64, 64, 3), dtype=np.uint8)
    ref = cv2.resize(img, (640, 640), interpolation=cv2.INTER_LINEAR)
    d = np.abs(ref.astype(int) - L.resize_linear_u8(img, 640, 640).astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
    big = np.random.default_rng(6).integers(0, 256, (300, 400, 3), dtype=np.uint8)
    for dw, dh in ((320, 240), (400, 100), (137, 300), (200, 150), (399, 299), (57, 31)):   # downscale / single-axis / exact 2x
        assert np.array_equal(cv2.resize(big, (dw, dh), interpolation=cv2.INTER_LINEAR), L.resize_linear_u8(big, dw, dh)), (dw, dh)
