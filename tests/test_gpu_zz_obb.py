"""GPU parity of the OBB head: angle towers + `ym_obb_finish` and the v0_1 obb model against the reference
golden and the CPU oracle.  The per-anchor function is also checked under g++ and the whole model on CPU
emulation in the CPU suite; on the B200 since round 2 (profiles/r02_gpu_suite.txt)."""
import os

import pytest
import torch

from _util import GOLD, assert_within_noise, synth_sd_from_keys, yaml_of
from oracle import yolo_master_oracle as O
from yolo_master_b200.nn.tasks import OBBModel
from yolo_master_b200.utils.synth import synth_images

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAME, CFG = "yolo-master-obb-n-v0_1", "master/v0_1/obb/yolo-master-obb-n.yaml"


def test_obb_model_matches_reference_golden():
    m = OBBModel(CFG)
    sd = synth_sd_from_keys(0, NAME)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    c = torch.load(os.path.join(GOLD, f"{NAME}.golden.pt"))["cases"]["b2_96"]
    x = synth_images(c["B"], c["H"], c["W"], c["seed"]).half()
    with torch.no_grad():
        y = m(x.to(DEV))[0].float().cpu()
    spec = O.parse_spec(yaml_of(CFG))
    ref = O.forward(spec, sd, x.float())
    with O.fp16_storage(), O.fp16_weights():
        sim = O.forward(spec, sd, x.float())
    assert y.shape == ref.shape
    assert_within_noise(y[:, :4], ref[:, :4], sim[:, :4], what="obb boxes")
    assert_within_noise(y[:, 4:-1], ref[:, 4:-1], sim[:, 4:-1], what="obb scores")
    assert_within_noise(y[:, -1:], ref[:, -1:], sim[:, -1:], what="obb angle")
    assert_within_noise(y, c["final"].float(), sim, what="obb vs reference golden")
    out = m.graphed(2, 96, 96)(x.to(DEV)).clone()                  # whole forward as one CUDA graph
    torch.cuda.synchronize()
    assert torch.equal(out.float().cpu(), y)
