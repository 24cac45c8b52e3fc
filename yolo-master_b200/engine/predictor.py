"""`DetectionPredictor`: the steps either side of the forward pass, device resident.

Mirrors the method names of ultralytics/engine/predictor.py (`preprocess` :155-176, `pre_transform` :186-204, `inference`
:178-184) and ultralytics/models/yolo/detect/predict.py (`postprocess` :32-74, `construct_results` :91-105,
`construct_result` :107-125).  The reference letterboxes every frame with cv2 on the host, stacks, transposes, uploads the
padded fp32/fp16 batch and divides by 255 on the device; here the RAW uint8 frames are uploaded and one kernel per group of
same-sized frames writes the letterboxed, channel-reversed planar batch the first convolution reads (`ym_letterbox_u8`); box
rescaling of the whole batch is one launch (`ym_scale_boxes`).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops as _k
from ..data.augment import LetterBox
from ..utils import nms, ops
from .results import Results


class DetectionPredictor:
    """`DetectionPredictor(model, imgsz=640, conf=0.25, iou=0.7, max_det=300, ...)`; call it with a list of BGR uint8 frames.

    Arguments follow the keys of ultralytics/cfg/default.yaml (`imgsz`, `conf`, `iou`, `max_det`, `classes`, `agnostic_nms`,
    `rect`, `half`, `cluster`, `sigma`).  `half=None` (default) feeds the model the uint8 batch - its first convolution scales
    by 1/255 while loading (same arithmetic as `im.half() / 255`); `half=True/False` reproduce the reference's fp16/fp32 tensor.
    """

    task = "detect"

    def __init__(self, model, imgsz=640, conf: float = 0.25, iou: float = 0.7, max_det: int = 300, classes=None,
                 agnostic_nms: bool = False, rect: bool = False, half=None, cluster: bool = False, sigma: float = 0.1, device=None):
        self.model = model
        end2end = bool(getattr(model, "end2end", False))
        if (classes is not None or agnostic_nms) and not end2end:
            # `non_max_suppression` would only refuse after the letterbox and the forward have run: fail at construction instead
            raise NotImplementedError("DetectionPredictor: classes= / agnostic_nms= are not on the B200 NMS path (end2end heads filter classes)")
        if not 1 <= int(max_det) <= 512:
            raise ValueError(f"DetectionPredictor: max_det must be in 1..512 (the NMS kernels keep a 512-entry survivor table), got {max_det}")
        self.imgsz = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        self.conf, self.iou, self.max_det = conf, iou, max_det
        self.classes, self.agnostic_nms, self.rect = classes, agnostic_nms, rect
        self.cluster, self.sigma = cluster, sigma
        self.half = half
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.stride = int(max(model.stride.tolist())) if hasattr(model, "stride") else 32
        self._letterbox = {}
        self._staging = {}
        self.batch = None

    # ------------------------------------------------------------------------------------------ before the forward pass
    def _copy_pool(self):
        pool = self.__dict__.get("_pool")
        if pool is None:
            from concurrent.futures import ThreadPoolExecutor
            pool = self.__dict__["_pool"] = ThreadPoolExecutor(max_workers=8, thread_name_prefix="ym-stage")
        return pool

    def _get_letterbox(self, auto: bool) -> LetterBox:
        lb = self._letterbox.get(auto)
        if lb is None:
            lb = self._letterbox[auto] = LetterBox(self.imgsz, auto=auto, stride=self.stride)
        return lb

    def pre_transform(self, im):
        """predictor.py:186-204 on the device: list of uint8 HWC frames -> list of letterboxed uint8 HWC CUDA tensors (BGR)."""
        same_shapes = len({tuple(x.shape) for x in im}) == 1
        lb = self._get_letterbox(bool(same_shapes and self.rect))
        return [lb(image=x) for x in im]

    def preprocess(self, im):
        """predictor.py:155-176.  im: (N, 3, H, W) tensor (moved / cast only, as in the reference) or a list of BGR uint8 HWC
        frames -> letterboxed RGB planar batch on the device."""
        out_dtype = torch.uint8 if self.half is None else (torch.float16 if self.half else torch.float32)
        if isinstance(im, torch.Tensor):
            im = im.to(self.device)
            return im if self.half is None else (im.half() if self.half else im.float())
        if len(im) == 0:
            raise ValueError("preprocess: empty frame list")
        shapes = [tuple(x.shape) for x in im]
        for s in shapes:
            if len(s) != 3 or s[2] != 3:
                raise ValueError(f"preprocess: expected (H, W, 3) frames, got {s}")
        same_shapes = len(set(shapes)) == 1
        lb = self._get_letterbox(bool(same_shapes and self.rect))
        plans = [lb.plan(s, self.device) for s in shapes]
        H, W = plans[0].H, plans[0].W
        if any((p.H, p.W) != (H, W) for p in plans):
            raise ValueError("preprocess: frames letterbox to different shapes (rect=True needs same-sized frames)")
        out = torch.empty((len(im), 3, H, W), dtype=out_dtype, device=self.device)
        groups = {}
        for i, s in enumerate(shapes):
            groups.setdefault(s, []).append(i)
        for s, idxs in groups.items():   # one upload + one launch per group of same-sized frames
            # pinned staging buffer, kept per (count, frame shape): allocating 88 MB of page-locked memory per call costs more than
            # the copy it serves.  Reuse is safe: every call ends with a host read of the detection counts, i.e. after its H2D.
            key = (len(idxs), s)
            host = self._staging.get(key)
            if host is None:
                if len(self._staging) >= 8:
                    self._staging.clear()
                host = self._staging[key] = torch.empty((len(idxs), *s), dtype=torch.uint8, pin_memory=self.device.type == "cuda")
            def stage(ji, host=host):
                j, i = ji
                frame = im[i]
                host[j].copy_(torch.from_numpy(np.ascontiguousarray(frame)) if isinstance(frame, np.ndarray) else frame)

            if len(idxs) >= 8 and host[0].numel() >= (1 << 20):
                # 32 frames of 720p are 88 MB of pageable -> pinned copies: one thread moves ~10 GB/s, i.e. as long as the whole forward.
                # The copies are independent and release the GIL: spread them over a small pool.
                list(self._copy_pool().map(stage, enumerate(idxs)))
            else:
                for ji in enumerate(idxs):
                    stage(ji)
            dev = host.to(self.device, non_blocking=True)
            contiguous_run = idxs == list(range(idxs[0], idxs[0] + len(idxs)))
            res = lb.apply_batch(dev, swap_rb=True, chw=True, dtype=out_dtype, out=out[idxs[0]:idxs[0] + len(idxs)] if contiguous_run else None)
            if not contiguous_run:
                out[torch.tensor(idxs, device=self.device)] = res
        return out

    def inference(self, im):
        return self.model(im)

    # ------------------------------------------------------------------------------------------ after the forward pass
    def _nc(self) -> int:
        """detect/predict.py:62: 0 for the detect task, else the class count (rows past 4 + nc ride along with the kept anchors)."""
        if self.task == "detect":
            return 0
        names = getattr(self.model, "names", None)
        return len(names) if names else int(self.model.model[-1].nc)

    def postprocess(self, preds, img, orig_imgs, **kwargs):
        """detect/predict.py:32-74: NMS (or the end2end confidence filter), then boxes back to the original frames."""
        if isinstance(preds, (list, tuple)):
            preds = preds[0]
        end2end = (bool(getattr(self.model, "end2end", False)) or preds.shape[-1] == 6) and self.task == "detect"
        frame_wh = (img.shape[3], img.shape[2]) if self.cluster else None
        dets = nms.non_max_suppression(preds, self.conf, kwargs.pop("iou", self.iou), self.classes, self.agnostic_nms,
                                       max_det=self.max_det, nc=self._nc(), end2end=end2end, rotated=self.task == "obb",
                                       cluster=self.cluster and not end2end, sigma=self.sigma, frame_wh=frame_wh)
        if self.cluster and not end2end:   # CW-NMS rows are (x, y, w, h): back to corners before rescaling
            for d in dets:
                d[:, 2:4] += d[:, 0:2]
        if isinstance(orig_imgs, torch.Tensor):   # tensor source: the "original" frames are the batch itself (predict.py:63-64)
            as_float = orig_imgs.float() / 255 if orig_imgs.dtype == torch.uint8 else orig_imgs.float()
            orig_imgs = list(ops.convert_torch2numpy_batch(as_float)[..., ::-1])
        return self.construct_results(dets, img, orig_imgs, **kwargs)

    def construct_results(self, preds, img, orig_imgs):
        """detect/predict.py:91-125 with the per-image `scale_boxes` calls folded into one launch over the ragged rows."""
        counts = [int(p.shape[0]) for p in preds]
        paths = self.batch[0] if self.batch else [None] * len(preds)
        if sum(counts):
            flat = torch.cat([p[:, :6] for p in preds]).float().contiguous()
            row_img = torch.repeat_interleave(torch.arange(len(preds), dtype=torch.int32, device=flat.device),
                                              torch.tensor(counts, device=flat.device))
            for s in range(0, len(preds), 128):
                lo, hi = sum(counts[:s]), sum(counts[:s + 128])
                if hi > lo:
                    ops.scale_boxes_batch(img.shape[2:], flat[lo:hi], [o.shape for o in orig_imgs[s:s + 128]],
                                          row_img=(row_img[lo:hi] - s).contiguous())
            rows = list(flat.split(counts))
        else:
            rows = [p[:, :6].float() for p in preds]
        names = getattr(self.model, "names", None)
        return [Results(o, path=pth, names=names, boxes=r) for r, o, pth in zip(rows, orig_imgs, paths)]

    def construct_result(self, pred, img, orig_img, img_path):
        """detect/predict.py:107-125 (single image)."""
        pred = pred[:, :6].float().contiguous()
        ops.scale_boxes(img.shape[2:], pred, orig_img.shape)
        return Results(orig_img, path=img_path, names=getattr(self.model, "names", None), boxes=pred)

    def __call__(self, source, paths=None):
        """list of BGR uint8 HWC frames -> list of `Results` (preprocess -> inference -> postprocess)."""
        frames = list(source) if not isinstance(source, torch.Tensor) else source
        self.batch = (paths or [None] * len(frames), frames, None)
        with torch.no_grad():
            im = self.preprocess(frames)
            preds = self.inference(im)
            return self.postprocess(preds, im, frames)


class SegmentationPredictor(DetectionPredictor):
    """ultralytics/models/yolo/segment/predict.py:11-113: NMS carrying the mask coefficients, then `ops.process_mask` per image
    (`ym_process_mask`: logits, x4 bilinear upsampling, crop and threshold without a float (n, H, W) intermediate in HBM)."""

    task = "segment"

    def __init__(self, *args, retina_masks: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        if retina_masks:
            raise NotImplementedError("SegmentationPredictor(retina_masks=True) (ops.process_mask_native) is not on the B200 path")

    def postprocess(self, preds, img, orig_imgs, **kwargs):
        """segment/predict.py:50-66: the model returns ((y, proto), aux) (or (y, proto) when exported)."""
        protos = preds[0][1] if isinstance(preds[0], (tuple, list)) else preds[1]
        return super().postprocess(preds[0], img, orig_imgs, protos=protos)

    def construct_results(self, preds, img, orig_imgs, protos):
        """segment/predict.py:68-84."""
        paths = self.batch[0] if self.batch else [None] * len(preds)
        return [self.construct_result(p, img, o, pth, proto) for p, o, pth, proto in zip(preds, orig_imgs, paths, protos)]

    def construct_result(self, pred, img, orig_img, img_path, proto):
        """segment/predict.py:86-111."""
        pred = pred.float().contiguous()
        if pred.shape[0] == 0:
            masks = None
        else:
            masks = _k.process_mask(proto, pred, img.shape[2:], upsample=True, coef_col=6)   # boxes / coefficients read in place
            ops.scale_boxes(img.shape[2:], pred[:, :4], orig_img.shape)
            keep = masks.amax((-2, -1)) > 0                              # only keep predictions with masks
            if not bool(keep.all()):
                pred, masks = pred[keep], masks[keep]
        return Results(orig_img, path=img_path, names=getattr(self.model, "names", None), boxes=pred[:, :6], masks=masks)


class OBBPredictor(DetectionPredictor):
    """ultralytics/models/yolo/obb/predict.py:10-58: rotated NMS (`ym_nms_rotated`), centres / sizes back to the original frame."""

    task = "obb"

    def construct_results(self, preds, img, orig_imgs):
        paths = self.batch[0] if self.batch else [None] * len(preds)
        return [self.construct_result(p, img, o, pth) for p, o, pth in zip(preds, orig_imgs, paths)]

    def construct_result(self, pred, img, orig_img, img_path):
        """obb/predict.py:41-58.  pred rows: (x, y, w, h, conf, cls, angle) -> obb rows (x, y, w, h, angle, conf, cls)."""
        pred = pred.float()
        rboxes = torch.cat([pred[:, :4], pred[:, -1:]], dim=-1).contiguous()
        ops.scale_boxes(img.shape[2:], rboxes[:, :4], orig_img.shape, xywh=True)
        obb = torch.cat([rboxes, pred[:, 4:6]], dim=-1)
        return Results(orig_img, path=img_path, names=getattr(self.model, "names", None), obb=obb)


class PosePredictor(DetectionPredictor):
    """ultralytics/models/yolo/pose/predict.py:9-66: NMS carrying the keypoint columns, boxes and keypoints back to the original frame
    (`ym_scale_boxes`, `ym_scale_coords`)."""

    task = "pose"

    def construct_results(self, preds, img, orig_imgs):
        paths = self.batch[0] if self.batch else [None] * len(preds)
        return [self.construct_result(p, img, o, pth) for p, o, pth in zip(preds, orig_imgs, paths)]

    def construct_result(self, pred, img, orig_img, img_path):
        """pose/predict.py:43-66."""
        result = super().construct_result(pred, img, orig_img, img_path)
        kpt_shape = tuple(getattr(self.model, "kpt_shape", None) or self.model.model[-1].kpt_shape)
        kpts = pred[:, 6:].float().reshape(pred.shape[0], *kpt_shape).contiguous()
        ops.scale_coords(img.shape[2:], kpts, orig_img.shape)
        result.keypoints = type(result)(orig_img, keypoints=kpts).keypoints
        return result
