"""Pins the oracle's v0 family support (ES_MOE in-model, A2C2f area attention +/- layer-scale residual, C3k, DFL Detect)
to outputs of the REAL reference (tests/golden/make_golden.py: yolo-master-n-v0 / yolo-master-l-v0)."""
import os

import pytest
import torch

from _util import GOLD, synth_sd_from_keys, yaml_of
from oracle import yolo_master_oracle as O
from yolo_master_b200.utils.synth import synth_images

CASES = [("yolo-master-n-v0", "master/v0/det/yolo-master-n.yaml", "b2_128"), ("yolo-master-n-v0", "master/v0/det/yolo-master-n.yaml", "b1_64"),
         ("yolo-master-l-v0", "master/v0/det/yolo-master-l.yaml", "b1_64"),
         # v0_1 zoo: ModularRouterExpertMoE (= OptimizedMOEImproved) as a top-level layer that owns its residual
         ("yolo-master-n-v0_1", "master/v0_1/det/yolo-master-n.yaml", "b2_128"),
         # Pose head on the v0_1 backbone (reference PoseModel): (B, 4 + nc + 17*3, A)
         ("yolo-master-pose-n-v0_1", "master/v0_1/pose/yolo-master-pose-n.yaml", "b2_128")]


@pytest.mark.parametrize("name,cfg,tag", CASES)
def test_oracle_matches_reference_v0(name, cfg, tag):
    c = torch.load(os.path.join(GOLD, f"{name}.golden.pt"))["cases"][tag]
    sd = synth_sd_from_keys(0, name)
    spec = O.parse_spec(yaml_of(cfg))
    x = synth_images(c["B"], c["H"], c["W"], c["seed"])
    y, ys = O.forward(spec, sd, x, return_layers=True)
    for i, ref in c["layers"].items():
        torch.testing.assert_close(ys[i], ref, atol=2e-4, rtol=1e-4, msg=lambda m, i=i: f"layer {i}: {m}")
    braw, sraw, _ = ys["detect_raw"]
    torch.testing.assert_close(braw, c["head_boxes"], atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(sraw, c["head_scores"], atol=2e-4, rtol=1e-4)
    ref = c["final"].float()
    tol = 2e-2 if c["final"].dtype == torch.float16 else 1e-3      # the L fixture stores the dense output in fp16
    rows = 5 + 51 if "pose" in name else 84
    assert y.shape == ref.shape == (c["B"], rows, (c["H"] // 8) ** 2 + (c["H"] // 16) ** 2 + (c["H"] // 32) ** 2)
    torch.testing.assert_close(y[:, :4], ref[:, :4], atol=tol, rtol=1e-3)        # xywh in pixels (DFL expectation)
    nsc = 1 if "pose" in name else 80
    torch.testing.assert_close(y[:, 4:4 + nsc], ref[:, 4:4 + nsc], atol=1e-3 if tol > 1e-3 else 1e-5, rtol=1e-3)
    torch.testing.assert_close(y[:, 4 + nsc:], ref[:, 4 + nsc:], atol=2e-3, rtol=1e-3)     # decoded keypoints (pixels / visibility)


def test_oracle_matches_reference_segment():
    """Segment head + Proto on the v0_1 backbone: dense prediction with the 32 mask coefficients appended, and the prototypes."""
    name, cfg = "yolo-master-seg-n-v0_1", "master/v0_1/seg/yolo-master-seg-n.yaml"
    c = torch.load(os.path.join(GOLD, f"{name}.golden.pt"))["cases"]["b2_96"]
    sd = synth_sd_from_keys(0, name)
    y, ys = O.forward(O.parse_spec(yaml_of(cfg)), sd, synth_images(c["B"], c["H"], c["W"], c["seed"]), return_layers=True)
    ref = c["final"].float()
    assert y.shape == ref.shape == (2, 4 + 80 + 32, 12 * 12 + 6 * 6 + 3 * 3)
    torch.testing.assert_close(y[:, :4], ref[:, :4], atol=1e-3, rtol=1e-3)
    torch.testing.assert_close(y[:, 4:84], ref[:, 4:84], atol=1e-5, rtol=1e-3)
    torch.testing.assert_close(y[:, 84:], ref[:, 84:], atol=3e-4, rtol=1e-3)          # mask coefficients
    assert ys["proto"].shape == c["proto"].shape == (2, 32, 24, 24)
    torch.testing.assert_close(ys["proto"], c["proto"], atol=3e-4, rtol=1e-3)


def test_oracle_matches_reference_obb():
    """OBB head on the v0_1 backbone: rotated (cx, cy, w, h), class scores and the angle row."""
    name, cfg = "yolo-master-obb-n-v0_1", "master/v0_1/obb/yolo-master-obb-n.yaml"
    c = torch.load(os.path.join(GOLD, f"{name}.golden.pt"))["cases"]["b2_96"]
    sd = synth_sd_from_keys(0, name)
    y = O.forward(O.parse_spec(yaml_of(cfg)), sd, synth_images(c["B"], c["H"], c["W"], c["seed"]))
    ref = c["final"].float()
    nc = ref.shape[1] - 5
    assert y.shape == ref.shape and ref.shape[2] == 12 * 12 + 6 * 6 + 3 * 3
    torch.testing.assert_close(y[:, :4], ref[:, :4], atol=1e-3, rtol=1e-3)
    torch.testing.assert_close(y[:, 4:4 + nc], ref[:, 4:4 + nc], atol=1e-5, rtol=1e-3)
    torch.testing.assert_close(y[:, -1], ref[:, -1], atol=1e-5, rtol=1e-4)             # angle in [-pi/4, 3pi/4]
    assert float(y[:, -1].min()) >= -0.7854 and float(y[:, -1].max()) <= 2.3562


def test_oracle_matches_reference_classify():
    """Classify head on the v0_1 backbone (reference ClassificationModel): logits and softmax probabilities over 1000 classes."""
    name, cfg = "yolo-master-cls-n-v0_1", "master/v0_1/cls/yolo-master-cls-n.yaml"
    c = torch.load(os.path.join(GOLD, f"{name}.golden.pt"))["cases"]["b3_64"]
    sd = synth_sd_from_keys(0, name)
    y, ys = O.forward(O.parse_spec(yaml_of(cfg)), sd, synth_images(c["B"], c["H"], c["W"], c["seed"]), return_layers=True)
    assert y.shape == (3, 1000)
    torch.testing.assert_close(ys[11], c["layers"][11], atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(ys["logits"], c["logits"], atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(y, c["final"], atol=1e-6, rtol=1e-4)
    assert torch.equal(y.argmax(1), c["final"].argmax(1))
