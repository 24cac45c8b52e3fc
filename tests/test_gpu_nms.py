"""GPU parity of batched NMS (bit-exact kept indices vs the REAL reference outputs) and CW-NMS (vs the float64 oracle)."""
import os

import numpy as np
import pytest
import torch

from _util import GOLD
from oracle import nms_oracle as N
from test_nms_oracle import synth_predictions
from yolo_master_b200.utils.nms import non_max_suppression

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_nms_matches_reference_golden_bit_exact():
    g = torch.load(os.path.join(GOLD, "nms.golden.pt"))
    for c in g["cases"]:
        pred = synth_predictions(c["B"], c["nc"], c["A"], c["seed"])
        out, keep = non_max_suppression(pred.to(DEV), c["conf"], c["iou"], max_det=c["max_det"], return_idxs=True)
        for o, k, ro, rk in zip(out, keep, c["out"], c["keep"]):
            assert torch.equal(k.cpu(), rk), "kept anchor indices differ from the reference"
            assert torch.equal(o.cpu(), ro), "NMS rows differ from the reference"


@pytest.mark.parametrize("B,nc,A,conf,iou", [(4, 80, 8400, 0.25, 0.7), (2, 80, 33600, 0.4, 0.5), (3, 1, 300, 0.01, 0.3), (1, 80, 64, 0.999, 0.5)])
def test_nms_matches_oracle(B, nc, A, conf, iou):
    pred = synth_predictions(B, nc, A, 900 + A)
    out, keep = non_max_suppression(pred.to(DEV), conf, iou, return_idxs=True)
    ro, rk = N.non_max_suppression(pred, conf, iou)
    for o, k, a, b in zip(out, keep, ro, rk):
        assert torch.equal(k.cpu(), b) and torch.equal(o.cpu(), a)


@pytest.mark.parametrize("B,nc,A,conf,iou,sigma", [(3, 80, 2100, 0.25, 0.5, 0.1), (2, 20, 8400, 0.3, 0.6, 0.5), (2, 80, 700, 0.25, 0.45, 0.0)])
def test_cw_nms_matches_oracle(B, nc, A, conf, iou, sigma):
    pred = synth_predictions(B, nc, A, 700 + A)
    out, keep = non_max_suppression(pred.to(DEV), conf, iou, cluster=True, sigma=sigma, frame_wh=(640, 640), return_idxs=True)
    for b in range(B):
        p = pred[b]
        sc, cl = p[4:].max(0)
        cx, cy, w, h = p[0], p[1], p[2], p[3]
        boxes = torch.stack([cx - 0.5 * w, cy - 0.5 * h, w, h], 1).numpy()
        dets, kept = N.cw_nms(boxes, sc.numpy(), cl.numpy(), conf, iou, sigma if sigma > 0 else 1.0, 300, 640, 640, cluster=sigma > 0)
        assert keep[b].cpu().tolist() == kept, "CW-NMS survivor set / order differs"
        np.testing.assert_allclose(out[b].cpu().numpy(), dets, rtol=1e-5, atol=1e-4)


def test_nms_past_sorter_capacity_takes_large_path_and_end2end_passthrough():
    """> 16384 candidates above conf in one image: the shared-memory sorter reports an overflow and `non_max_suppression` switches to
    `ym_nms_batched_large` (utils/nms.py) - same kept anchors / rows as the oracle.  CW-NMS (`cluster=True`) has no large path: loud."""
    g = torch.Generator().manual_seed(5)
    pred = torch.rand((1, 6, 20000), generator=g) * 0.5 + 0.5   # 20000 candidates above conf: exceeds the 16384 sorter
    pred[:, :2] *= 600
    pred[:, 2:4] *= 40
    out, keep = non_max_suppression(pred.to(DEV), 0.25, 0.5, return_idxs=True)
    ro, rk = N.non_max_suppression(pred, 0.25, 0.5)
    assert torch.equal(keep[0].cpu(), rk[0]) and torch.equal(out[0].cpu(), ro[0])
    with pytest.raises(RuntimeError, match="16384"):
        non_max_suppression(pred.to(DEV), 0.25, 0.5, cluster=True, frame_wh=(640, 640))
    e2e = torch.rand((2, 300, 6), device=DEV)
    out = non_max_suppression(e2e, 0.5, 0.7)
    assert all((o[:, 4] > 0.5).all() for o in out) and out[0].shape[1] == 6


def test_cw_nms_matches_reference_cpp_golden():
    """`ym_nms_batched(mode=1)` against the reference's own compiled C++ (`decode_candidates` + `nms_and_cap`, common.cpp:93-205;
    tests/golden/make_cwnms_golden.py): survivors / order / score / class exact, boxes to fp32 rounding of the float64 weighted mean."""
    g = torch.load(os.path.join(GOLD, "cwnms.golden.pt"))
    for c in g["cases"]:
        pred = synth_predictions(c["B"], c["nc"], c["A"], c["seed"], c["dense"])
        out = non_max_suppression(pred.to(DEV), c["conf"], c["iou"], max_det=c["max_det"], cluster=True,
                                  sigma=c["sigma"] if c["cluster"] else 0.0, frame_wh=(c["frame_w"], c["frame_h"]))
        for b in range(c["B"]):
            ref = c["dets"][b].numpy()
            o = out[b].cpu().numpy()
            assert o.shape == ref.shape, (c["seed"], b, o.shape, ref.shape)
            assert np.array_equal(o[:, 4:], ref[:, 4:]), "survivor set / order differs from the reference C++"
            np.testing.assert_allclose(o[:, :4], ref[:, :4], rtol=1e-5, atol=1e-4)
