"""CPU oracles for NMS and Cluster-Weighted NMS.  TEST INFRASTRUCTURE ONLY (see yolo_master_oracle.py).

* `non_max_suppression` restates `ultralytics/utils/nms.py:13-171` (single-label, class-aware, non-rotated branch) with the
  greedy kernel of `TorchNMS.nms` (:245-302).  Pinned against the real reference function by
  tests/golden/make_golden.py -> nms.golden.pt (tests/test_nms_oracle.py).
* `cw_nms` restates the ONLY executable specification of CW-NMS in the reference snapshot, the C++ deployment code
  `examples/YOLO-Master-Cross-Platform-Edge-Deployment/cpp/src/common.cpp:56-198` (`box_iou`, `nms_greedy`, `nms_and_cap`),
  in float64 like the C++.  **Pinned** to the reference's own C++: oracle/Makefile compiles common.cpp as it lies (against
  oracle/cv_shim, a stand-in for the few cv:: value types; there is no OpenCV in this image) into oracle/_ref/libcwnms_ref.so,
  tests/golden/make_cwnms_golden.py runs its `decode` (= decode_candidates + nms_and_cap) into tests/golden/cwnms.golden.pt and
  tests/test_nms_oracle.py holds `cw_nms` to it: survivor set / order / score / class exact, boxes to one fp32 ulp (incl. the
  > 3000-candidate pool cap, Standard mode, max_det cap and frame clipping).  Hand-derived known answers are kept beside it.
"""
from __future__ import annotations

import numpy as np
import torch


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    xy, wh = x[..., :2], x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def greedy_nms(boxes: torch.Tensor, scores: torch.Tensor, iou_thr: float) -> torch.Tensor:
    """`TorchNMS.nms` nms.py:245-302: score-descending, suppress IoU > thr (boxes with IoU <= thr survive)."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    x1, y1, x2, y2 = boxes.unbind(1)
    areas = (x2 - x1) * (y2 - y1)
    order = scores.argsort(dim=0, descending=True, stable=True)
    keep = []
    while order.numel() > 0:
        i = order[0]
        keep.append(int(i))
        if order.numel() == 1:
            break
        rest = order[1:]
        w = (torch.minimum(x2[i], x2[rest]) - torch.maximum(x1[i], x1[rest])).clamp_(min=0)
        h = (torch.minimum(y2[i], y2[rest]) - torch.maximum(y1[i], y1[rest])).clamp_(min=0)
        inter = w * h
        iou = inter / (areas[i] + areas[rest] - inter)
        order = rest[~(iou > iou_thr)]
    return torch.tensor(keep, dtype=torch.int64)


def non_max_suppression(prediction: torch.Tensor, conf_thres=0.25, iou_thres=0.45, max_det=300, max_nms=30000, max_wh=7680):
    """prediction (B, 4+nc, A) with xywh boxes.  Returns (list of (n,6) [x1,y1,x2,y2,conf,cls], list of anchor indices)."""
    bs, no, A = prediction.shape
    nc = no - 4
    xc = prediction[:, 4:].amax(1) > conf_thres
    pred = prediction.transpose(-1, -2).clone()
    pred[..., :4] = xywh2xyxy(pred[..., :4])
    outs, idxs = [], []
    for xi in range(bs):
        x = pred[xi][xc[xi]]
        xk = torch.arange(A)[xc[xi]]
        if not x.shape[0]:
            outs.append(torch.zeros((0, 6)))
            idxs.append(torch.zeros((0,), dtype=torch.int64))
            continue
        box, cls = x[:, :4], x[:, 4:]
        conf, j = cls.max(1, keepdim=True)
        filt = conf.view(-1) > conf_thres
        x = torch.cat((box, conf, j.float()), 1)[filt]
        xk = xk[filt]
        n = x.shape[0]
        if not n:
            outs.append(torch.zeros((0, 6)))
            idxs.append(torch.zeros((0,), dtype=torch.int64))
            continue
        if n > max_nms:
            f = x[:, 4].argsort(descending=True, stable=True)[:max_nms]
            x, xk = x[f], xk[f]
        c = x[:, 5:6] * max_wh
        i = greedy_nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        outs.append(x[i])
        idxs.append(xk[i])
    return outs, idxs


# ---------------------------------------------------------------------------------------------------------------------
def _iou_xywh(a, b):
    """common.cpp:56-67 `box_iou` on (x, y, w, h) rectangles, float64; `uni > 0 ? inter/uni : 0`."""
    xx1, yy1 = max(a[0], b[0]), max(a[1], b[1])
    xx2, yy2 = min(a[0] + a[2], b[0] + b[2]), min(a[1] + a[3], b[1] + b[3])
    inter = max(0.0, xx2 - xx1) * max(0.0, yy2 - yy1)
    uni = a[2] * a[3] + b[2] * b[3] - inter
    return inter / uni if uni > 0 else 0.0


def cw_nms(boxes_xywh: np.ndarray, scores: np.ndarray, classes: np.ndarray, conf: float, iou_thr: float, sigma: float,
           max_det: int, frame_w: float, frame_h: float, cluster: bool = True):
    """common.cpp:127-198 `nms_and_cap` on one image.

    boxes_xywh: (n,4) top-left x, y, w, h in frame pixels (float32 as produced by decode_candidates :96-125).
    Returns (dets (m,6) float32 [x, y, w, h, conf, cls] clipped to the frame, kept candidate indices in score order).
    """
    n = len(scores)
    idx = [i for i in range(n) if not (scores[i] < conf)]                       # :131-135
    boxes = [tuple(float(v) for v in boxes_xywh[i]) for i in idx]              # Rect2d from the float candidates
    sc = [float(np.float32(scores[i])) for i in idx]
    OFF = 2.0 * max(frame_w, frame_h) + 8192.0                                   # :141
    off = [(b[0] + int(classes[i]) * OFF, b[1] + int(classes[i]) * OFF, b[2], b[3]) for b, i in zip(boxes, idx)]
    # nms_greedy :71-89 (std::sort is not stable: ties are broken towards the lower index here)
    order = sorted([k for k in range(len(sc)) if sc[k] >= conf], key=lambda k: (-sc[k], k))
    dead = [False] * len(sc)
    keep = []
    for m_, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        for j in order[m_ + 1:]:
            if not dead[j] and _iou_xywh(off[i], off[j]) > iou_thr:
                dead[j] = True
    refined = None
    if cluster and keep:                                                          # :150-177
        pool = sorted(range(len(sc)), key=lambda k: (-sc[k], k))[:3000]
        refined = []
        for k in keep:
            sw = ax = ay = ax2 = ay2 = 0.0
            for m_ in pool:
                ov = _iou_xywh(off[k], off[m_])
                if ov <= iou_thr:
                    continue
                w = sc[m_] * np.exp(-((1.0 - ov) ** 2) / sigma)
                sw += w
                ax += w * boxes[m_][0]
                ay += w * boxes[m_][1]
                ax2 += w * (boxes[m_][0] + boxes[m_][2])
                ay2 += w * (boxes[m_][1] + boxes[m_][3])
            if sw > 1e-6:
                x0, y0 = ax / sw, ay / sw
                refined.append((x0, y0, max(0.0, ax2 / sw - x0), max(0.0, ay2 / sw - y0)))
            else:
                refined.append(boxes[k])
    dets, kept = [], []
    for s_, k in enumerate(keep):                                                 # :180-197
        if len(dets) >= max_det:
            break
        raw = boxes[k] if refined is None else refined[s_]
        x0, y0 = max(raw[0], 0.0), max(raw[1], 0.0)                              # cv::Rect2d & frame
        x1, y1 = min(raw[0] + raw[2], float(frame_w)), min(raw[1] + raw[3], float(frame_h))
        w, h = x1 - x0, y1 - y0
        if w > 0 and h > 0:
            dets.append((np.float32(x0), np.float32(y0), np.float32(w), np.float32(h), np.float32(sc[k]), np.float32(classes[idx[k]])))
            kept.append(idx[k])
    return np.array(dets, dtype=np.float32).reshape(-1, 6), kept
