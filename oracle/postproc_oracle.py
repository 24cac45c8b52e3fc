"""CPU oracles for the task post-processing after the Segment and OBB heads.  TEST INFRASTRUCTURE ONLY (see yolo_master_oracle.py):
imported by tests/ alone, never by the product path.

* `process_mask` restates `ultralytics/utils/ops.py:500-528` (with `crop_mask` :477-497) in fp32 torch: coefficients @ prototypes,
  `F.interpolate(..., mode="bilinear")` to the network input size, separable crop, `> 0`, uint8.
* `non_max_suppression_rotated` restates the `rotated=True` path of `ultralytics/utils/nms.py:13-171`: xywh boxes kept as they
  are (:90), best class (:133-134), top `max_nms` by score (:142-146), class offset on the centre (:148,151),
  `TorchNMS.fast_nms` (:193-242) with `batch_probiou` (`utils/metrics.py:293-326`, covariance :224-242), `[:max_det]` (:160).
  Score ties are broken towards the lower anchor index (the reference's `argsort(descending=True)` leaves them unspecified).

Both are pinned against the real reference functions by tests/golden/make_golden.py -> postproc.golden.pt
(tests/test_postproc_oracle.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def mask_field(protos: torch.Tensor, masks_in: torch.Tensor, shape, upsample: bool) -> torch.Tensor:
    """The fp32 field process_mask thresholds (before the crop): (n, H, W) with upsample else (n, mh, mw).  ops.py:517-521."""
    c, mh, mw = protos.shape
    masks = (masks_in.float() @ protos.float().view(c, -1)).view(-1, mh, mw)
    if upsample:
        masks = F.interpolate(masks[None], tuple(shape), mode="bilinear")[0]
    return masks


def crop_keep(boxes: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """crop_mask ops.py:489-496 as a boolean (n, h, w) grid: x1 <= col < x2 and y1 <= row < y2."""
    x1, y1, x2, y2 = torch.chunk(boxes[:, :, None].float(), 4, 1)
    r = torch.arange(w, dtype=torch.float32)[None, None, :]
    c = torch.arange(h, dtype=torch.float32)[None, :, None]
    return ((r >= x1) & (r < x2)) & ((c >= y1) & (c < y2))


def process_mask(protos, masks_in, bboxes, shape, upsample: bool = False) -> torch.Tensor:
    """ops.py:500-528."""
    c, mh, mw = protos.shape
    if masks_in.shape[0] == 0:
        return torch.zeros((0, *(tuple(shape) if upsample else (mh, mw))), dtype=torch.uint8)
    field = mask_field(protos, masks_in, shape, upsample)
    if upsample:
        keep = crop_keep(bboxes, *field.shape[1:])
    else:
        ratios = torch.tensor([[mw / shape[1], mh / shape[0], mw / shape[1], mh / shape[0]]])   # fp32, ops.py:523-525
        keep = crop_keep(bboxes.float() * ratios, mh, mw)
    return ((field > 0) & keep).to(torch.uint8)


def covariance(boxes: torch.Tensor):
    """metrics.py:224-242 (_get_covariance_matrix, floor 0)."""
    a = boxes[:, 2:3].pow(2) / 12
    b = boxes[:, 3:4].pow(2) / 12
    c = boxes[:, 4:5]
    cos, sin = c.cos(), c.sin()
    cos2, sin2 = cos.pow(2), sin.pow(2)
    return a * cos2 + b * sin2, a * sin2 + b * cos2, (a - b) * cos * sin


def batch_probiou(obb1: torch.Tensor, obb2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """metrics.py:293-326: (N, 5) x (M, 5) xywhr -> (N, M)."""
    x1, y1 = obb1[..., :2].split(1, dim=-1)
    x2, y2 = (x.squeeze(-1)[None] for x in obb2[..., :2].split(1, dim=-1))
    a1, b1, c1 = covariance(obb1)
    a2, b2, c2 = (x.squeeze(-1)[None] for x in covariance(obb2))
    den = (a1 + a2) * (b1 + b2) - (c1 + c2).pow(2) + eps
    t1 = (((a1 + a2) * (y1 - y2).pow(2) + (b1 + b2) * (x1 - x2).pow(2)) / den) * 0.25
    t2 = (((c1 + c2) * (x2 - x1) * (y1 - y2)) / den) * 0.5
    t3 = (((a1 + a2) * (b1 + b2) - (c1 + c2).pow(2))
          / (4 * ((a1 * b1 - c1.pow(2)).clamp(0) * (a2 * b2 - c2.pow(2)).clamp(0)).sqrt() + eps) + eps).log() * 0.5
    bd = (t1 + t2 + t3).clamp(eps, 100.0)
    hd = (1.0 - (-bd).exp() + eps).sqrt()
    return 1 - hd


def fast_nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, iou_thr: float):
    """TorchNMS.fast_nms nms.py:193-242 with iou_func=batch_probiou.  Returns (kept indices in score order, margin) where margin
    is the smallest |ProbIoU - iou_thr| over the pairs that decide (i < j): decisions closer than the arithmetic noise of another
    implementation are not comparable."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64), float("inf")
    order = torch.sort(scores, descending=True, stable=True).indices
    b = boxes[order]
    ious = batch_probiou(b, b).triu_(diagonal=1)
    pick = torch.nonzero((ious >= iou_thr).sum(0) <= 0).squeeze(-1)
    n = b.shape[0]
    upper = torch.triu(torch.ones((n, n), dtype=torch.bool), diagonal=1)
    margin = (ious - iou_thr).abs()[upper].min().item() if n > 1 else float("inf")
    return order[pick], margin


def non_max_suppression_rotated(prediction: torch.Tensor, conf_thres=0.25, iou_thres=0.45, max_det=300, max_nms=30000, max_wh=7680):
    """prediction: (B, 4 + nc + 1, A) = xywh, class scores, angle.  Returns (rows [(n, 7) x, y, w, h, conf, cls, angle], kept anchor
    indices, smallest decision margin)."""
    B, no, A = prediction.shape
    nc = no - 5
    outs, keeps, margin = [], [], float("inf")
    for xi in range(B):
        x = prediction[xi].transpose(0, 1).float()                    # (A, 4 + nc + 1)
        anchors = torch.arange(A)
        xc = x[:, 4:4 + nc].amax(1) > conf_thres                      # nms.py:76
        x, anchors = x[xc], anchors[xc]
        if not x.shape[0]:
            outs.append(torch.zeros((0, 7)))
            keeps.append(torch.zeros((0,), dtype=torch.int64))
            continue
        box, cls, ang = x[:, :4], x[:, 4:4 + nc], x[:, 4 + nc:]
        conf, j = cls.max(1, keepdim=True)                            # nms.py:133
        x = torch.cat((box, conf, j.float(), ang), 1)
        if x.shape[0] > max_nms:                                      # nms.py:142-146
            top = torch.sort(x[:, 4], descending=True, stable=True).indices[:max_nms]
            x, anchors = x[top], anchors[top]
        c = x[:, 5:6] * max_wh                                        # nms.py:148
        boxes = torch.cat((x[:, :2] + c, x[:, 2:4], x[:, -1:]), dim=-1)   # nms.py:151
        i, m = fast_nms_rotated(boxes, x[:, 4], iou_thres)
        i = i[:max_det]
        margin = min(margin, m)
        outs.append(x[i])
        keeps.append(anchors[i])
    return outs, keeps, margin
