"""GPU parity of the v0_1 zoo model (yolo-master-n: `ModularRouterExpertMoE` = `OptimizedMOEImproved` as a top-level layer that
owns its residual, A2C2f area attention, DFL Detect) against the reference golden and the CPU oracle, through the public API.

Every kernel on this path is verified on hardware by the other suites (the MoE-FFN chain inside A2C2fMoE, test_gpu_ops.py); this
is the model-level composition (on the B200 since round 2, profiles/r02_gpu_suite.txt)."""
import os

import pytest
import torch

from _util import GOLD, assert_within_noise, synth_sd_from_keys, yaml_of
from oracle import yolo_master_oracle as O
from yolo_master_b200.nn.tasks import DetectionModel
from yolo_master_b200.utils.synth import synth_images

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAME, CFG = "yolo-master-n-v0_1", "master/v0_1/det/yolo-master-n.yaml"


@pytest.fixture(scope="module")
def model():
    m = DetectionModel(CFG)
    sd = synth_sd_from_keys(0, NAME)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd, O.parse_spec(yaml_of(CFG))


def test_v0_1_model_matches_reference_golden(model):
    m, sd, spec = model
    c = torch.load(os.path.join(GOLD, f"{NAME}.golden.pt"))["cases"]["b2_128"]
    x = synth_images(c["B"], c["H"], c["W"], c["seed"]).half()
    feats = {}
    hooks = [mod.register_forward_hook(lambda mod, i, o, k=k: feats.__setitem__(k, o)) for k, mod in enumerate(m.model)]
    with torch.no_grad():
        y = m(x.to(DEV))[0].float().cpu()
    for h in hooks:
        h.remove()
    ref, _ = O.forward(spec, sd, x.float(), return_layers=True)
    with O.fp16_storage(), O.fp16_weights():
        ysim, sim = O.forward(spec, sd, x.float(), return_layers=True)
    for i, g in c["layers"].items():
        assert_within_noise(feats[i], g, sim[i], what=f"{NAME} layer {i} vs reference golden")
    assert_within_noise(y[:, :4], ref[:, :4], ysim[:, :4], what=f"{NAME} boxes")
    assert_within_noise(y[:, 4:], ref[:, 4:], ysim[:, 4:], what=f"{NAME} scores")


def test_v0_1_per_image_independence_and_graph(model):
    m, _, _ = model
    x = synth_images(3, 256, 256, 5).half().to(DEV)
    with torch.no_grad():
        eager = m(x)[0].clone()
        alone = m(x[1:2])[0]
    assert torch.equal(eager[1], alone[0])                       # image-level routing: a batch is a set of independent images
    out = m.graphed(3, 256, 256)(x).clone()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
