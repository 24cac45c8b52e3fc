"""`non_max_suppression` with the reference's signature (ultralytics/utils/nms.py:13-171) on the batched CUDA kernel, plus
Cluster-Weighted NMS (`cluster=True`, the `cluster` / `sigma` keys of ultralytics/cfg/default.yaml:195-198 whose only
executable specification is the C++ deployment code, examples/.../cpp/src/common.cpp:127-198)."""
from __future__ import annotations

import torch

from .. import ops


def non_max_suppression(prediction, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None, agnostic: bool = False,
                        multi_label: bool = False, labels=(), max_det: int = 300, nc: int = 0, max_time_img: float = 0.05,
                        max_nms: int = 30000, max_wh: int = 7680, rotated: bool = False, end2end: bool = False,
                        return_idxs: bool = False, cluster: bool = False, sigma: float = 0.1, frame_wh=None):
    """Returns a list (one (n,6) tensor per image) like the reference; `return_idxs=True` also returns anchor indices.

    cluster=False: rows are (x1, y1, x2, y2, conf, cls).  cluster=True (CW-NMS): rows are (x, y, w, h, conf, cls) in frame
    pixels, clipped to `frame_wh` (required)."""
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if prediction.shape[-1] == 6 or end2end:  # end-to-end model output (B, N, 6): mask + slice only (nms.py:66-75)
        output, keepi = [], []
        for pred in prediction:
            mask = pred[:, 4] > conf_thres
            if classes is not None:
                mask &= (pred[:, 5:6] == torch.tensor(classes, device=pred.device)).any(1)
            idx = mask.nonzero(as_tuple=False).view(-1)[:max_det]
            output.append(pred[idx])
            keepi.append(idx)
        return (output, keepi) if return_idxs else output
    if classes is not None or agnostic or multi_label or len(labels):
        raise NotImplementedError("non_max_suppression: classes / agnostic / multi_label / labels are not on the B200 path")
    if rotated:
        # OBB rows (B, 4 + nc + 1, A): xywh stays (nms.py:90), fast-NMS on ProbIoU (nms.py:148-152); rows (x, y, w, h, conf, cls, angle)
        if cluster or (nc and prediction.shape[1] - 4 - nc != 1):
            raise NotImplementedError("non_max_suppression(rotated=True): expects (B, 4 + nc + 1, A) predictions, no cluster mode")
        out, cnt, idx = ops.nms_rotated(prediction.float(), conf_thres, iou_thres, max_det, max_nms, float(max_wh))
        counts = cnt.tolist()
        output = [out[b, :n] for b, n in enumerate(counts)]
        return (output, [idx[b, :n].long() for b, n in enumerate(counts)]) if return_idxs else output
    extra = None
    if nc and prediction.shape[1] - 4 != nc:      # rows past 4 + nc (keypoints, mask coefficients) ride along with the kept anchors
        if cluster:
            raise NotImplementedError("non_max_suppression(cluster=True) with extra columns is not on the B200 path")
        extra = prediction[:, 4 + nc:]
        prediction = prediction[:, :4 + nc].contiguous()
    if cluster and frame_wh is None:
        raise ValueError("non_max_suppression(cluster=True) needs frame_wh=(width, height)")
    out, cnt, idx, scratch = ops.nms_batched(prediction.float(), conf_thres, iou_thres, max_det, max_nms, float(max_wh),
                                             1 if cluster else 0, sigma, frame_wh or (0.0, 0.0))
    B, _, A = prediction.shape
    if ops.lib().ym_nms_overflowed(scratch.data_ptr(), B, A, torch.cuda.current_stream().cuda_stream):
        # some image has more candidates than the shared-memory sorter holds (validation at conf 0.001): global-memory path
        if cluster:
            raise RuntimeError("non_max_suppression(cluster=True): more than 16384 candidates above conf_thres in one image")
        out, cnt, idx = ops.nms_batched_large(prediction.float(), conf_thres, iou_thres, max_det, max_nms, float(max_wh))
    counts = cnt.tolist()  # the reference returns ragged Python lists: one host read of B integers
    output = [out[b, :n] for b, n in enumerate(counts)]
    if extra is not None:                       # nms.py:124-131: x = cat(box, conf, cls, mask) - gathered by the kept anchor indices
        output = [torch.cat([o, extra[b][:, idx[b, :n].long()].t().to(o.dtype)], 1) for b, (o, n) in enumerate(zip(output, counts))]
    if return_idxs:
        return output, [idx[b, :n].long() for b, n in enumerate(counts)]
    return output
