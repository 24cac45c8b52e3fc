"""Box utilities of the predictor's post-processing with the reference's signatures (ultralytics/utils/ops.py), in place on
CUDA fp32 tensors through `ym_scale_boxes`."""
from __future__ import annotations

import torch

from .. import ops as _k


def _gain_pad(img1_shape, img0_shape, ratio_pad=None):
    """utils/ops.py:143-149."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad_x = round((img1_shape[1] - round(img0_shape[1] * gain)) / 2 - 0.1)
        pad_y = round((img1_shape[0] - round(img0_shape[0] * gain)) / 2 - 0.1)
    else:
        gain = ratio_pad[0][0]
        pad_x, pad_y = ratio_pad[1]
    return gain, pad_x, pad_y


def box_params(img1_shape, img0_shapes, ratio_pads=None) -> torch.Tensor:
    """Host fp32 (n_img, 5) rows (gain, pad_x, pad_y, w0, h0) for `scale_boxes_batch`."""
    rows = []
    for i, s0 in enumerate(img0_shapes):
        gain, px, py = _gain_pad(img1_shape, s0, None if ratio_pads is None else ratio_pads[i])
        rows.append((gain, px, py, s0[1], s0[0]))
    return torch.tensor(rows, dtype=torch.float32).reshape(-1, 5)


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding: bool = True, xywh: bool = False):
    """utils/ops.py:119-158: rescale (N, >= 4) boxes from the letterboxed `img1_shape` to `img0_shape`, in place; clips unless xywh."""
    if boxes.numel() == 0:
        return boxes
    return _k.scale_boxes(boxes, box_params(img1_shape, [img0_shape], None if ratio_pad is None else [ratio_pad]),
                          rows_per_img=boxes.numel() // boxes.shape[-1], padding=padding, xywh=xywh)


def scale_boxes_batch(img1_shape, boxes, img0_shapes, row_img=None, padding: bool = True, xywh: bool = False):
    """All images of a batch in one launch: boxes (B, K, >= 4) with image = first index, or flat (n, >= 4) rows whose image
    is `row_img[row]` (int32 CUDA; at most 128 images per call - the per-image parameters ride in the kernel's parameter block)."""
    if boxes.numel() == 0:
        return boxes
    params = box_params(img1_shape, img0_shapes)
    if row_img is not None:
        if params.shape[0] > 128:
            raise ValueError("scale_boxes_batch: ragged rows are limited to 128 images per call")
        return _k.scale_boxes(boxes, params, row_img=row_img, padding=padding, xywh=xywh)
    if boxes.dim() != 3 or boxes.shape[0] != len(img0_shapes):
        raise ValueError("scale_boxes_batch: expected (B, K, >= 4) boxes with one original shape per image")
    for s in range(0, params.shape[0], 128):
        _k.scale_boxes(boxes[s:s + 128], params[s:s + 128].contiguous(), rows_per_img=boxes.shape[1], padding=padding, xywh=xywh)
    return boxes


def convert_torch2numpy_batch(batch):
    """utils/ops.py:683-692: (B, C, H, W) float images in [0, 1] -> (B, H, W, C) uint8 numpy."""
    return (batch.permute(0, 2, 3, 1).contiguous() * 255).clamp(0, 255).byte().cpu().numpy()


def clip_boxes(boxes, shape):
    """utils/ops.py:174-201 (tensor branch), in place."""
    if boxes.numel() == 0:
        return boxes
    p = torch.tensor([[1.0, 0.0, 0.0, shape[1], shape[0]]], dtype=torch.float32)
    return _k.scale_boxes(boxes, p, rows_per_img=boxes.numel() // boxes.shape[-1], padding=False, xywh=False)


def process_mask(protos, masks_in, bboxes, shape, upsample: bool = False):
    """utils/ops.py:500-528: (mask_dim, mh, mw) prototypes x (N, mask_dim) coefficients -> uint8 (N, H, W) masks cropped to the
    xyxy `bboxes` (given in `shape` coordinates); H, W = `shape` when upsample else the prototype resolution."""
    dets = torch.cat([bboxes.float(), masks_in.float()], 1).contiguous()
    return _k.process_mask(protos, dets, shape, upsample=upsample, coef_col=4)


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None, normalize: bool = False, padding: bool = True):
    """utils/ops.py:596-631: rescale (..., 2 | 3) points from the letterboxed `img1_shape` to `img0_shape` and clip them, in place."""
    if coords.numel() == 0:
        return coords
    return _k.scale_coords(coords, box_params(img1_shape, [img0_shape], None if ratio_pad is None else [ratio_pad]).reshape(5).contiguous(),
                           padding=padding, normalize=normalize)
