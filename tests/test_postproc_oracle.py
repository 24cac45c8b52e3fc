"""The post-processing oracle (oracle/postproc_oracle.py) against the REAL reference functions' outputs stored by
tests/golden/make_golden.py -> postproc.golden.pt: ops.process_mask (both branches) and non_max_suppression(rotated=True)."""
import os

import numpy as np
import torch

from _util import GOLD
from oracle import postproc_oracle as P

G = torch.load(os.path.join(GOLD, "postproc.golden.pt"))


def unpack(case, key):
    shape = tuple(case[key + "_shape"])
    n = int(np.prod(shape))
    return torch.from_numpy(np.unpackbits(case[key + "_bits"].numpy())[:n].reshape(shape))


def test_process_mask_matches_reference_bit_for_bit():
    for i, c in enumerate(G["masks"]):
        for key, up in (("up", True), ("native", False)):
            want = unpack(c, key)
            got = P.process_mask(c["protos"], c["coef"], c["boxes"], c["shape"], upsample=up)
            assert got.dtype == torch.uint8 and got.shape == want.shape, (i, key)
            assert torch.equal(got, want), (i, key, int((got != want).sum()))


def test_process_mask_empty_and_crop_edges():
    protos = torch.ones((4, 6, 6))
    assert P.process_mask(protos, torch.zeros((0, 4)), torch.zeros((0, 4)), (24, 24), True).shape == (0, 24, 24)
    assert P.process_mask(protos, torch.zeros((0, 4)), torch.zeros((0, 4)), (24, 24), False).shape == (0, 6, 6)
    m = P.process_mask(protos, torch.ones((1, 4)), torch.tensor([[2.0, 3.5, 7.0, 9.0]]), (24, 24), True)[0]
    assert m[:, :2].sum() == 0 and m[:, 7:].sum() == 0 and m[:4].sum() == 0 and m[9:].sum() == 0     # x1 <= col < x2, y1 <= row < y2
    assert m[4:9, 2:7].min() == 1


def test_rotated_nms_matches_reference():
    for i, c in enumerate(G["nms"]):
        outs, keeps, margin = P.non_max_suppression_rotated(c["pred"], c["conf"], c["iou"], c["max_det"], c["max_nms"])
        assert margin > 1e-6, (i, margin)
        for b, (o, k) in enumerate(zip(outs, keeps)):
            assert torch.equal(k, c["keep"][b]), (i, b)
            assert torch.equal(o, c["out"][b]), (i, b)


def test_probiou_known_answers():
    a = torch.tensor([[10.0, 10.0, 8.0, 4.0, 0.3]])
    assert abs(P.batch_probiou(a, a).item() - (1 - (1e-7 + 1e-7) ** 0.5)) < 2e-4          # identical boxes: bd -> eps
    far = torch.tensor([[500.0, 500.0, 8.0, 4.0, 0.3]])
    assert P.batch_probiou(a, far).item() < 1e-3
    rot = torch.tensor([[10.0, 10.0, 4.0, 8.0, 0.3 + torch.pi / 2]])                       # same Gaussian: w/h swapped + 90 degrees
    assert abs(P.batch_probiou(a, rot).item() - P.batch_probiou(a, a).item()) < 1e-4
