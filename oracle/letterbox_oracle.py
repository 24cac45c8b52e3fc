"""CPU oracle for the predictor's pre- and post-processing - TEST INFRASTRUCTURE ONLY: the checker of
`ym_letterbox_u8` / `ym_scale_boxes` and of the host mirror in yolo-master_b200/{data,engine,utils/ops.py}.

Restates, in integer arithmetic, what `LetterBox.__call__` (ultralytics/data/augment.py:1705-1800: `get_params` :1742-1786, then
cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder(value=114)) and `BasePredictor.preprocess` (engine/predictor.py:155-176: BGR->RGB,
HWC->CHW) do to one uint8 frame.  cv2.resize on 8-bit data is third-party code (opencv-python, version floor only in the
reference's pyproject.toml); its generic kernel (modules/imgproc/src/resize.cpp: `HResizeLinear` + `VResizeLinear<uchar,int,short>`)
is restated here from its published algorithm:
  * source coordinate fx = (float)((dx + 0.5) * scale - 0.5), floor; at the border x clamps index AND weight (fx = 0), y clips only
    the two row indices and keeps the weights (the border row is blended with itself);
  * 11-bit fixed-point coefficients a = saturate_cast<short>(w * 2048) (round half to even), horizontal pass in int32;
  * vertical pass  dst = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
  * an exact 2x downscale in both axes takes the INTER_AREA fast path: (a + b + c + d + 2) >> 2.
Pinned (tests/test_letterbox_oracle.py, fixtures generated with cv2 4.13 by tests/golden/make_golden.py): bit-exact (CRC) for
every fixture - downscales, identity, single-axis and two-axis upscales.
"""
from __future__ import annotations

import numpy as np


def letterbox_params(shape_hw, new_shape=(640, 640), scaleup=True, center=True, auto=False, scale_fill=False, stride=32):
    """`LetterBox.get_params` augment.py:1742-1790: returns (new_unpad (w, h), top, bottom, left, right)."""
    h, w = shape_hw
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = (round(w * r), round(h * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = dw % stride, dh % stride
    elif scale_fill:
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
    if center:
        dw /= 2
        dh /= 2
    top, bottom = (round(dh - 0.1) if center else 0), round(dh + 0.1)
    left, right = (round(dw - 0.1) if center else 0), round(dw + 0.1)
    return new_unpad, top, bottom, left, right


def _coeffs(dn, sn, scale, clamp_weights):
    """Source indices and 11-bit weights of one axis.  OpenCV clamps the WEIGHTS at the image border only along x
    (resize.cpp: `if (sx < 0) fx = 0, sx = 0`); along y the two row indices are clipped (`clip(sy0 + k, 0, height)`) but the
    fractional weights are kept, so a border row is blended with itself through two separately truncated products."""
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= sn - 1
        f[hi], s[hi] = 0, sn - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
    a1 = np.rint(f * np.float32(2048)).astype(np.int32)
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), a0, a1


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC images (generic OpenCV kernel)."""
    sh, sw = src.shape[:2]
    if (dw, dh) == (sw, sh):
        return src.copy()
    sx, sy = 1.0 / (dw / sw), 1.0 / (dh / sh)
    if sx == 2.0 and sy == 2.0:
        s = src.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, ax0, ax1 = _coeffs(dw, sw, sx, True)
    y0, y1, ay0, ay1 = _coeffs(dh, sh, sy, False)
    s = src.astype(np.int32)
    rows = s[:, x0] * ax0[None, :, None] + s[:, x1] * ax1[None, :, None]
    out = ((((ay0[:, None, None] * (rows[y0] >> 4)) >> 16) + ((ay1[:, None, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2)
    return np.clip(out, 0, 255).astype(np.uint8)


def preprocess_frame(img_bgr_hwc: np.ndarray, new_shape=(640, 640), pad_value=114, **lb) -> np.ndarray:
    """One frame through LetterBox + the predictor's layout change: uint8 HWC BGR -> uint8 CHW RGB (`lb`: LetterBox options)."""
    out = letterbox_frame(img_bgr_hwc, new_shape, pad_value, **lb)
    return np.ascontiguousarray(out[..., ::-1].transpose(2, 0, 1))


def letterbox_frame(img: np.ndarray, new_shape=(640, 640), pad_value=114, **lb) -> np.ndarray:
    """`LetterBox.apply_image` augment.py:1792-1822: resize if needed, constant border; HWC in, HWC out."""
    (nw, nh), top, bottom, left, right = letterbox_params(img.shape[:2], new_shape, **lb)
    if img.shape[:2] != (nh, nw):
        img = resize_linear_u8(img, nw, nh)
    out = np.full((nh + top + bottom, nw + left + right, img.shape[2]), pad_value, dtype=np.uint8)
    out[top:top + nh, left:left + nw] = img
    return out


def scale_boxes(img1_shape, boxes: np.ndarray, img0_shape, padding=True, xywh=False) -> np.ndarray:
    """`ops.scale_boxes` + `clip_boxes` utils/ops.py:119-158,174-201 on fp32 rows (n, >= 4); returns a new array.
    fp32 arithmetic in the reference's order: subtract the integer padding, divide by float32(gain), clamp to the original frame."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - round(img0_shape[1] * gain)) / 2 - 0.1)
    pad_y = round((img1_shape[0] - round(img0_shape[0] * gain)) / 2 - 0.1)
    b = boxes.astype(np.float32).copy()
    if padding:
        b[..., 0] -= np.float32(pad_x)
        b[..., 1] -= np.float32(pad_y)
        if not xywh:
            b[..., 2] -= np.float32(pad_x)
            b[..., 3] -= np.float32(pad_y)
    b[..., :4] = b[..., :4] / np.float32(gain)
    if not xywh:
        h, w = img0_shape[:2]
        b[..., [0, 2]] = b[..., [0, 2]].clip(0, w)
        b[..., [1, 3]] = b[..., [1, 3]].clip(0, h)
    return b


def scale_coords(img1_shape, coords: np.ndarray, img0_shape, normalize=False, padding=True) -> np.ndarray:
    """`ops.scale_coords` + `clip_coords` utils/ops.py:596-631,204-225 on fp32 points (..., >= 2); returns a new array (fp32 arithmetic
    in the reference's order: subtract the integer padding, divide by float32(gain), clamp, optional normalisation)."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - round(img0_shape[1] * gain)) / 2 - 0.1)
    pad_y = round((img1_shape[0] - round(img0_shape[0] * gain)) / 2 - 0.1)
    c = coords.astype(np.float32).copy()
    if padding:
        c[..., 0] -= np.float32(pad_x)
        c[..., 1] -= np.float32(pad_y)
    c[..., 0] = c[..., 0] / np.float32(gain)
    c[..., 1] = c[..., 1] / np.float32(gain)
    h, w = img0_shape[:2]
    c[..., 0] = c[..., 0].clip(0, w)
    c[..., 1] = c[..., 1].clip(0, h)
    if normalize:
        c[..., 0] = c[..., 0] / np.float32(w)
        c[..., 1] = c[..., 1] / np.float32(h)
    return c
