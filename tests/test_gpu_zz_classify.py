"""GPU parity of the Classify head (`ym_classify_head`) and the v0_1 classification model (`ClassificationModel`: BatchNorm eps 1e-5,
unlike detection models) against the reference golden and the CPU oracle.  The kernel body also runs under g++ and the whole model on CPU
emulation in the CPU suite; on the B200 since round 2 (profiles/r02_gpu_suite.txt)."""
import os

import pytest
import torch

from _util import GOLD, assert_within_noise, synth_sd_from_keys, yaml_of
from oracle import yolo_master_oracle as O
from yolo_master_b200 import ops
from yolo_master_b200.nn.tasks import ClassificationModel
from yolo_master_b200.utils.synth import synth_images

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAME, CFG = "yolo-master-cls-n-v0_1", "master/v0_1/cls/yolo-master-cls-n.yaml"


def test_classify_head_kernel():
    g = torch.Generator().manual_seed(2)
    v = torch.randn((5, 1, 1, 1280), generator=g).half()
    w, b = torch.randn((1000, 1280), generator=g) * 0.05, torch.randn((1000,), generator=g)
    probs, logits = ops.classify_head(v.to(DEV), w.to(DEV), b.to(DEV))
    want = v.view(5, 1280).float() @ w.t() + b
    torch.testing.assert_close(logits.cpu(), want, atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(probs.cpu(), torch.softmax(want, 1), atol=1e-6, rtol=1e-3)


def test_classification_model_matches_reference_golden():
    m = ClassificationModel(CFG)
    sd = synth_sd_from_keys(0, NAME)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    c = torch.load(os.path.join(GOLD, f"{NAME}.golden.pt"))["cases"]["b3_64"]
    x = synth_images(c["B"], c["H"], c["W"], c["seed"]).half()
    with torch.no_grad():
        y, logits = m(x.to(DEV))
    spec = O.parse_spec(yaml_of(CFG))
    _, rys = O.forward(spec, sd, x.float(), return_layers=True)
    with O.fp16_storage(), O.fp16_weights():
        _, sim = O.forward(spec, sd, x.float(), return_layers=True)
    assert_within_noise(logits, rys["logits"], sim["logits"], what="cls logits")
    assert_within_noise(logits, c["logits"], sim["logits"], what="cls logits vs reference golden")
    torch.testing.assert_close(y.sum(1).cpu(), torch.ones(3), atol=1e-5, rtol=0)
