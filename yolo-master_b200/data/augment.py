"""`LetterBox` with the reference's constructor and `get_params` (ultralytics/data/augment.py:1646-1822), executed on the GPU.

The reference resizes with cv2 (INTER_LINEAR on uint8 = 11-bit fixed point) and pads with cv2.copyMakeBorder on the host; here
the host only derives the geometry and two small tap tables per (source shape, target shape), and one kernel
(`ym_letterbox_u8`) writes the padded frame - optionally already channel-reversed, planar and scaled, which is what
`BasePredictor.preprocess` (engine/predictor.py:155-176) does next.  Results are bit-identical to the cv2 pipeline.
Only the image path of `__call__` exists (labels / instances / masks belong to training, which is out of scope).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .. import ops

INTER_LINEAR = 1  # cv2.INTER_LINEAR


def _axis_taps(dn: int, sn: int, clamp_weights: bool) -> np.ndarray:
    """Tap table of one axis of cv2's 8-bit bilinear resize: (dn, 2) uint32 rows {i0 | i1 << 16, a0 | a1 << 16}.

    Source coordinate in float32 as cv2 computes it; 11-bit weights rounded half-to-even.  cv2 treats the borders of the two
    axes differently: along x an out-of-range tap is moved onto the border pixel with weight 2048 / 0, along y only the two
    row indices are clipped and the fractional weights stay."""
    scale = 1.0 / (dn / sn)
    pos = ((np.arange(dn, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    i0 = np.floor(pos).astype(np.int64)
    frac = (pos - i0.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        below, above = i0 < 0, i0 >= sn - 1
        frac[below | above] = 0
        i0[below] = 0
        i0[above] = sn - 1
    a1 = np.rint(frac * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1) - frac) * np.float32(2048)).astype(np.int64)
    i1 = np.clip(i0 + 1, 0, sn - 1)
    i0 = np.clip(i0, 0, sn - 1)
    return np.stack([i0 | (i1 << 16), a0 | (a1 << 16)], 1).astype(np.uint32)


class _Plan:
    """Geometry + device tables of one (source shape -> letterboxed shape) mapping."""

    __slots__ = ("params", "nw", "nh", "top", "left", "H", "W", "area2x", "xtab", "ytab")

    def __init__(self, params, shape, device):
        self.params = params
        self.nw, self.nh = params["new_unpad"]
        if self.nw < 1 or self.nh < 1:   # cv2.resize raises on an empty destination as well
            raise ValueError(f"LetterBox: a {shape[0]}x{shape[1]} frame resizes to an empty {self.nh}x{self.nw} image")
        self.top, self.left = params["top"], params["left"]
        self.H, self.W = self.nh + params["top"] + params["bottom"], self.nw + params["left"] + params["right"]
        sh, sw = shape
        self.area2x = 1.0 / (self.nw / sw) == 2.0 and 1.0 / (self.nh / sh) == 2.0   # cv2 switches to its 2x2 area average
        self.xtab = self.ytab = None
        if not self.area2x:
            tabs = np.concatenate([_axis_taps(self.nw, sw, True), _axis_taps(self.nh, sh, False)]).view(np.int32)
            dev = torch.from_numpy(tabs).to(device)
            self.xtab, self.ytab = dev[: self.nw], dev[self.nw:]


class LetterBox:
    """Resize-and-pad to `new_shape` keeping the aspect ratio; same arguments as the reference (augment.py:1672-1706)."""

    _MAX_PLANS = 64

    def __init__(self, new_shape=(640, 640), auto: bool = False, scale_fill: bool = False, scaleup: bool = True,
                 center: bool = True, stride: int = 32, padding_value: int = 114, interpolation: int = INTER_LINEAR):
        if interpolation != INTER_LINEAR:
            raise NotImplementedError("LetterBox: only cv2.INTER_LINEAR is implemented on the B200 path")
        self.new_shape = new_shape
        self.auto = auto
        self.scale_fill = scale_fill
        self.scaleup = scaleup
        self.stride = stride
        self.center = center
        self.padding_value = padding_value
        self.interpolation = interpolation
        self._plans: OrderedDict = OrderedDict()

    def get_params(self, labels) -> dict:
        """augment.py:1742-1790; `labels["img"]` only needs a `.shape`."""
        shape = tuple(labels["img"].shape[:2])
        new_shape = labels.pop("rect_shape", self.new_shape)
        if isinstance(new_shape, int):
            new_shape = (new_shape, new_shape)
        r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
        if not self.scaleup:
            r = min(r, 1.0)
        ratio = r, r
        new_unpad = round(shape[1] * r), round(shape[0] * r)
        dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
        if self.auto:
            dw, dh = int(np.mod(dw, self.stride)), int(np.mod(dh, self.stride))
        elif self.scale_fill:
            dw, dh = 0.0, 0.0
            new_unpad = (new_shape[1], new_shape[0])
            ratio = new_shape[1] / shape[1], new_shape[0] / shape[0]
        if self.center:
            dw /= 2
            dh /= 2
        top, bottom = (round(dh - 0.1) if self.center else 0), round(dh + 0.1)
        left, right = (round(dw - 0.1) if self.center else 0), round(dw + 0.1)
        return {"orig_shape": shape, "new_shape": new_shape, "ratio": ratio, "new_unpad": new_unpad,
                "top": top, "bottom": bottom, "left": left, "right": right}

    def plan(self, shape, device) -> _Plan:
        key = (tuple(shape[:2]), str(device))
        p = self._plans.get(key)
        if p is None:
            class _Shape:   # get_params reads nothing but the shape
                pass
            probe = _Shape()
            probe.shape = tuple(shape[:2])
            p = _Plan(self.get_params({"img": probe}), shape[:2], device)
            self._plans[key] = p
            if len(self._plans) > self._MAX_PLANS:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        return p

    def apply_batch(self, frames: torch.Tensor, swap_rb: bool = False, chw: bool = False, dtype=torch.uint8, out=None) -> torch.Tensor:
        """frames: uint8 CUDA (B, h, w, 3), all of one shape.  -> (B, H, W, 3), or (B, 3, H, W) when `chw`."""
        p = self.plan(frames.shape[1:3], frames.device)
        return ops.letterbox(frames, p.xtab, p.ytab, p.area2x, p.nw, p.nh, p.top, p.left, p.H, p.W, self.padding_value,
                             swap_rb, chw, dtype, out)

    def __call__(self, labels=None, image=None):
        """Image-only form of augment.py:1708-1740: uint8 HWC frame (numpy or tensor) -> letterboxed uint8 HWC CUDA tensor."""
        if labels:
            raise NotImplementedError("LetterBox: the labels / instances path (training) is not on the B200 path")
        if image is None:
            raise ValueError("LetterBox: pass image=<uint8 HWC frame>")
        frame = torch.from_numpy(np.ascontiguousarray(image)) if isinstance(image, np.ndarray) else image
        if frame.dtype != torch.uint8 or frame.dim() != 3 or frame.shape[2] != 3:
            raise ValueError(f"LetterBox: expected a uint8 (H, W, 3) frame, got {tuple(frame.shape)} {frame.dtype}")
        return self.apply_batch(frame.cuda().contiguous()[None])[0]
