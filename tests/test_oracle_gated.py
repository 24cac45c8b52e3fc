"""Pins the oracle's gated-MoE restatement (`VisualEnhancedAdaptiveGateMoE`: SE-gated split, detail gate, dual-stream router,
batch-level complexity gate, low-rank fused / shared-inverted expert groups, pyramid context, refinement) to outputs of the REAL
reference v0_10 model (tests/golden/make_golden.py: yolo-master-n-v0_10).  SURVEY.md §8(f) rank 1: oracle first, CUDA path next."""
import os

import pytest
import torch
import yaml

from _util import GOLD, synth_sd_from_keys
from oracle import yolo_master_oracle as O
from yolo_master_b200.utils.synth import synth_images

V010 = """
nc: 80
scales:
  n: [0.50, 0.25, 1024]
backbone:
  - [-1, 1, Conv, [64, 3, 2]]
  - [-1, 1, Conv, [128, 3, 2]]
  - [-1, 2, C3k2, [256, False, 0.25]]
  - [-1, 1, Conv, [256, 3, 2]]
  - [-1, 2, C3k2, [512, False, 0.25]]
  - [-1, 1, VisualEnhancedAdaptiveGateMoE, [512, 4, 2, 0.5]]
  - [-1, 1, Conv, [512, 3, 2]]
  - [-1, 4, A2C2f, [512, True, 4]]
  - [-1, 1, VisualEnhancedAdaptiveGateMoE, [512, 8, 2, 0.5]]
  - [-1, 1, Conv, [1024, 3, 2]]
  - [-1, 4, A2C2f, [1024, True, 1]]
  - [-1, 1, VisualEnhancedAdaptiveGateMoE, [1024, 16, 2, 0.5]]
head:
  - [-1, 1, nn.Upsample, [None, 2, "nearest"]]
  - [[-1, 8], 1, Concat, [1]]
  - [-1, 2, C3k2, [512, True]]
  - [-1, 1, nn.Upsample, [None, 2, "nearest"]]
  - [[-1, 5], 1, Concat, [1]]
  - [-1, 2, C3k2, [256, True]]
  - [-1, 1, Conv, [256, 3, 2]]
  - [[-1, 14], 1, Concat, [1]]
  - [-1, 2, C3k2, [512, True]]
  - [-1, 1, Conv, [512, 3, 2]]
  - [[-1, 11], 1, Concat, [1]]
  - [-1, 2, C3k2, [512, True]]
  - [[17, 20, 23], 1, Detect, [nc]]
"""   # layer table of ultralytics/cfg/models/master/v0_10/det/yolo-master-n.yaml (the product does not ship this family yet)


@pytest.mark.parametrize("tag", ["b2_160", "b1_128"])
def test_oracle_matches_reference_gated(tag):
    name = "yolo-master-n-v0_10"
    c = torch.load(os.path.join(GOLD, f"{name}.golden.pt"))["cases"][tag]
    sd = synth_sd_from_keys(0, name)
    spec = O.parse_spec(yaml.safe_load(V010))
    x = synth_images(c["B"], c["H"], c["W"], c["seed"])
    y, ys = O.forward(spec, sd, x, return_layers=True)
    for i, ref in c["layers"].items():
        torch.testing.assert_close(ys[i], ref, atol=3e-4, rtol=1e-4, msg=lambda m, i=i: f"layer {i}: {m}")
    braw, sraw, _ = ys["detect_raw"]
    torch.testing.assert_close(braw, c["head_boxes"], atol=3e-4, rtol=1e-4)
    torch.testing.assert_close(sraw, c["head_scores"], atol=3e-4, rtol=1e-4)
    torch.testing.assert_close(y[:, 4:], c["final"][:, 4:], atol=1e-5, rtol=1e-3)
    torch.testing.assert_close(y[:, :4], c["final"][:, :4], atol=2e-3, rtol=1e-3)


def test_complexity_gate_known_answers():
    """gated.py:469-490: k=2 -> keep round(c*2) ranks with c clamped to [0.3, 1.5]: c=0.3 -> 1 rank (weights renormalise to 1, 0);
    c=0.74 -> round(1.48) = 1; c=0.76 -> 2; torch.round is half-to-even: c=0.75 -> round(1.5) = 2."""
    w = torch.tensor([[0.7, 0.3], [0.6, 0.4]])
    assert torch.allclose(O.complexity_gate(w, torch.tensor(0.1)), torch.tensor([[1.0, 0.0], [1.0, 0.0]]))
    assert torch.allclose(O.complexity_gate(w, torch.tensor(0.74)), torch.tensor([[1.0, 0.0], [1.0, 0.0]]))
    assert torch.allclose(O.complexity_gate(w, torch.tensor(0.76)), w)
    assert torch.allclose(O.complexity_gate(w, torch.tensor(0.75)), w)
    assert torch.allclose(O.complexity_gate(w, torch.tensor(float("nan"))), w)


# ----------------------------------------------------------------------------------------------------------------------
# The whole AdaptiveGateMoE line (v0_4 ... v0_10 zoos): module-level goldens of the REAL reference classes
# (tests/golden/make_golden.py gated_family), top-2 of 4 (fused / low-rank back-ends) and of 16 (shared-inverted) experts.
FAMILY = torch.load(os.path.join(GOLD, "gated_family.golden.pt"))


@pytest.mark.parametrize("key", sorted(FAMILY), ids=lambda k: k.replace("AdaptiveGateMoE", "AGM"))
def test_gated_family_oracle_matches_reference_module(key):
    from yolo_master_b200.utils.synth import fill_state_dict_
    name, E = key.split("/E")
    c = FAMILY[key]
    sd = {k: torch.zeros(shape, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32) for k, shape in c["keys"].items()}
    fill_state_dict_(sd, c["seed"])
    sd.update(c["scalars"])                                      # 0-dim parameters (alpha, *_scale) stay at their init values
    sd = {"m." + k: v for k, v in sd.items()}
    x = torch.randn((2, 64, c["hw"], c["hw"]), generator=torch.Generator().manual_seed(c["xseed"]))
    if name in O.GATED_VARIANTS:
        assert O.gated_backend(name, int(E)) == c["backend"]
    extra = () if name == "UltraOptimizedMoE" else (c["split"],)             # (in, out, num_experts, top_k[, split_ratio])
    y, w, idx, _ = O._LAYER_FN[name](sd, "m", x, 64, 64, int(E), 2, *extra, return_route=True)
    assert torch.equal(idx, c["route_idx"])                      # router choices (taken before the complexity gate in the fixture)
    torch.testing.assert_close(y, c["y"], atol=2e-4, rtol=2e-4)
