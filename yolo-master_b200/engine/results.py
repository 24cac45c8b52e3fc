"""`Results` / `Boxes` containers with the attribute names of ultralytics/engine/results.py:184-300,860-1100 (detection, segmentation and oriented-box subset).
Plain holders: the tensors stay wherever the predictor produced them (CUDA) until `.cpu()` / `.numpy()` is asked for."""
from __future__ import annotations

import torch


def _xyxy2xywh(x):
    y = torch.empty_like(x)
    y[..., 0] = (x[..., 0] + x[..., 2]) / 2
    y[..., 1] = (x[..., 1] + x[..., 3]) / 2
    y[..., 2] = x[..., 2] - x[..., 0]
    y[..., 3] = x[..., 3] - x[..., 1]
    return y


class Boxes:
    """(n, 6) rows (x1, y1, x2, y2, conf, cls) in pixels of the original frame (results.py:860-1100)."""

    def __init__(self, boxes, orig_shape):
        if boxes.ndim == 1:
            boxes = boxes[None, :]
        if boxes.shape[-1] not in (6, 7):
            raise ValueError(f"expected 6 or 7 values but got {boxes.shape[-1]}")
        self.data = boxes
        self.orig_shape = tuple(orig_shape)
        self.is_track = boxes.shape[-1] == 7

    @property
    def shape(self):
        return self.data.shape

    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, -2]

    @property
    def cls(self):
        return self.data[:, -1]

    @property
    def id(self):
        return self.data[:, -3] if self.is_track else None

    @property
    def xywh(self):
        return _xyxy2xywh(self.xyxy)

    @property
    def xyxyn(self):
        xyxy = self.xyxy.clone()
        xyxy[..., [0, 2]] /= self.orig_shape[1]
        xyxy[..., [1, 3]] /= self.orig_shape[0]
        return xyxy

    @property
    def xywhn(self):
        xywh = self.xywh
        xywh[..., [0, 2]] /= self.orig_shape[1]
        xywh[..., [1, 3]] /= self.orig_shape[0]
        return xywh

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return Boxes(self.data[idx], self.orig_shape)

    def cpu(self):
        return Boxes(self.data.cpu(), self.orig_shape)

    def cuda(self):
        return Boxes(self.data.cuda(), self.orig_shape)

    def to(self, *args, **kwargs):
        return Boxes(self.data.to(*args, **kwargs), self.orig_shape)

    def numpy(self):
        b = Boxes.__new__(Boxes)
        b.data, b.orig_shape, b.is_track = self.data.cpu().numpy(), self.orig_shape, self.is_track
        return b


class Masks:
    """(n, H, W) uint8 masks at the network input size (results.py:1080-1150): `data`, `orig_shape`, `shape`."""

    def __init__(self, masks, orig_shape):
        if masks.ndim == 2:
            masks = masks[None, :]
        self.data = masks
        self.orig_shape = tuple(orig_shape)

    @property
    def shape(self):
        return self.data.shape

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return Masks(self.data[idx], self.orig_shape)

    def cpu(self):
        return Masks(self.data.cpu(), self.orig_shape)

    def cuda(self):
        return Masks(self.data.cuda(), self.orig_shape)

    def to(self, *args, **kwargs):
        return Masks(self.data.to(*args, **kwargs), self.orig_shape)

    def numpy(self):
        return Masks(self.data.cpu().numpy(), self.orig_shape)


class Keypoints:
    """(n, nk, 2 | 3) keypoints in pixels of the original frame (results.py:1180-1290): `data`, `xy`, `conf`, `has_visible`."""

    def __init__(self, keypoints, orig_shape):
        if keypoints.ndim == 2:
            keypoints = keypoints[None, :]
        self.data = keypoints
        self.orig_shape = tuple(orig_shape)
        self.has_visible = self.data.shape[-1] == 3

    @property
    def shape(self):
        return self.data.shape

    @property
    def xy(self):
        return self.data[..., :2]

    @property
    def conf(self):
        return self.data[..., 2] if self.has_visible else None

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return Keypoints(self.data[idx], self.orig_shape)

    def cpu(self):
        return Keypoints(self.data.cpu(), self.orig_shape)

    def cuda(self):
        return Keypoints(self.data.cuda(), self.orig_shape)

    def to(self, *args, **kwargs):
        return Keypoints(self.data.to(*args, **kwargs), self.orig_shape)

    def numpy(self):
        return Keypoints(self.data.cpu().numpy(), self.orig_shape)


class OBB:
    """(n, 7) rows (x, y, w, h, angle, conf, cls) in pixels of the original frame (results.py:1380-1560)."""

    def __init__(self, boxes, orig_shape):
        if boxes.ndim == 1:
            boxes = boxes[None, :]
        if boxes.shape[-1] not in (7, 8):
            raise ValueError(f"expected 7 or 8 values but got {boxes.shape[-1]}")
        self.data = boxes
        self.orig_shape = tuple(orig_shape)
        self.is_track = boxes.shape[-1] == 8

    @property
    def shape(self):
        return self.data.shape

    @property
    def xywhr(self):
        return self.data[:, :5]

    @property
    def conf(self):
        return self.data[:, -2]

    @property
    def cls(self):
        return self.data[:, -1]

    @property
    def id(self):
        return self.data[:, -3] if self.is_track else None

    @property
    def xyxyxyxy(self):
        """utils/ops.py xywhr2xyxyxyxy: the four corners, (n, 4, 2)."""
        ctr, w, h, ang = self.data[:, :2], self.data[:, 2:3], self.data[:, 3:4], self.data[:, 4:5]
        cos, sin = torch.cos(ang), torch.sin(ang)
        v1 = torch.cat([w / 2 * cos, w / 2 * sin], -1)
        v2 = torch.cat([-h / 2 * sin, h / 2 * cos], -1)
        return torch.stack([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2], -2)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return OBB(self.data[idx], self.orig_shape)

    def cpu(self):
        return OBB(self.data.cpu(), self.orig_shape)

    def cuda(self):
        return OBB(self.data.cuda(), self.orig_shape)

    def to(self, *args, **kwargs):
        return OBB(self.data.to(*args, **kwargs), self.orig_shape)

    def numpy(self):
        return OBB(self.data.cpu().numpy(), self.orig_shape)


class Results:
    """One image's detections (results.py:184-300): `orig_img`, `orig_shape`, `boxes`, `masks`, `obb`, `names`, `path`, `speed`."""

    def __init__(self, orig_img, path=None, names=None, boxes=None, speed=None, masks=None, obb=None, keypoints=None):
        self.orig_img = orig_img
        self.orig_shape = tuple(orig_img.shape[:2])
        self.boxes = Boxes(boxes, self.orig_shape) if boxes is not None else None
        self.probs = None
        self.keypoints = Keypoints(keypoints, self.orig_shape) if keypoints is not None else None
        self.masks = Masks(masks, self.orig_shape) if masks is not None else None
        self.obb = OBB(obb, self.orig_shape) if obb is not None else None
        self.speed = speed if speed is not None else {"preprocess": None, "inference": None, "postprocess": None}
        self.names = names
        self.path = path
        self.save_dir = None
        self._keys = ("boxes",)

    def __len__(self):
        for k in ("boxes", "masks", "obb"):                           # results.py:262-274: the first non-empty field
            v = getattr(self, k)
            if v is not None:
                return len(v)
        return 0

    def __getitem__(self, idx):
        return Results(self.orig_img, self.path, self.names, None if self.boxes is None else self.boxes.data[idx], self.speed,
                       None if self.masks is None else self.masks.data[idx], None if self.obb is None else self.obb.data[idx],
                       None if self.keypoints is None else self.keypoints.data[idx])

    def _apply(self, fn, *args, **kwargs):
        r = Results(self.orig_img, self.path, self.names, None, self.speed)
        for k in ("boxes", "masks", "obb", "keypoints"):
            v = getattr(self, k)
            if v is not None:
                setattr(r, k, getattr(v, fn)(*args, **kwargs))
        return r

    def cpu(self):
        return self._apply("cpu")

    def numpy(self):
        return self._apply("numpy")

    def cuda(self):
        return self._apply("cuda")

    def to(self, *args, **kwargs):
        return self._apply("to", *args, **kwargs)

    def summary(self, normalize: bool = False, decimals: int = 5):
        """results.py:606-660 (detection rows): list of {name, class, confidence, box}."""
        out = []
        if self.boxes is None:
            return out
        h, w = self.orig_shape if normalize else (1, 1)
        for row in self.boxes.data.cpu().tolist():
            c = int(row[5])
            out.append({"name": (self.names or {}).get(c, str(c)), "class": c, "confidence": round(row[4], decimals),
                        "box": {"x1": round(row[0] / w, decimals), "y1": round(row[1] / h, decimals),
                                "x2": round(row[2] / w, decimals), "y2": round(row[3] / h, decimals)}})
        return out
