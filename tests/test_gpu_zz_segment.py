"""GPU parity of the Segment head: mask-coefficient towers, prototype branch (the 2x2 stride-2 transposed
convolution as a 1x1 convolution + depth-to-space) and the v0_1 seg model against the reference golden and the CPU oracle.
The whole-model wiring is also verified on CPU emulation; on the B200 since round 2 (profiles/r02_gpu_suite.txt)."""
import os

import pytest
import torch
import torch.nn.functional as F

from _util import GOLD, assert_within_noise, synth_sd_from_keys, yaml_of
from oracle import yolo_master_oracle as O
from yolo_master_b200.nn.modules.head import Proto
from yolo_master_b200.nn.tasks import SegmentationModel
from yolo_master_b200.utils.synth import fill_state_dict_, synth_images

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAME, CFG = "yolo-master-seg-n-v0_1", "master/v0_1/seg/yolo-master-seg-n.yaml"


def test_proto_matches_torch():
    m = Proto(64, 32, 16)
    sd = m.state_dict()
    fill_state_dict_(sd, 3)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    x = torch.randn((2, 64, 12, 10), generator=torch.Generator().manual_seed(1)).half()
    with torch.no_grad():
        y = m(x.to(DEV).contiguous(memory_format=torch.channels_last)).float().cpu()
    sdm = {"p." + k: v.float() for k, v in sd.items()}
    fn = lambda: O.conv_block(sdm, "p.cv3", O.conv_block(sdm, "p.cv2", O._st(F.conv_transpose2d(
        O.conv_block(sdm, "p.cv1", x.float()), O._w(sdm["p.upsample.weight"]), sdm["p.upsample.bias"], 2, 0))))
    ref = fn()
    with O.fp16_storage(), O.fp16_weights():
        sim = fn()
    assert y.shape == ref.shape == (2, 16, 24, 20)
    assert_within_noise(y, ref, sim, what="Proto")


def test_segment_model_matches_reference_golden():
    m = SegmentationModel(CFG)
    sd = synth_sd_from_keys(0, NAME)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    c = torch.load(os.path.join(GOLD, f"{NAME}.golden.pt"))["cases"]["b2_96"]
    x = synth_images(c["B"], c["H"], c["W"], c["seed"]).half()
    with torch.no_grad():
        (y, proto), _ = m(x.to(DEV))
    spec = O.parse_spec(yaml_of(CFG))
    ref, ys = O.forward(spec, sd, x.float(), return_layers=True)
    with O.fp16_storage(), O.fp16_weights():
        sim, ss = O.forward(spec, sd, x.float(), return_layers=True)
    y = y.float().cpu()
    assert y.shape == (2, 116, 189) and proto.shape == (2, 32, 24, 24)
    assert_within_noise(y[:, :4], ref[:, :4], sim[:, :4], what="seg boxes")
    assert_within_noise(y[:, 4:84], ref[:, 4:84], sim[:, 4:84], what="seg scores")
    assert_within_noise(y[:, 84:], ref[:, 84:], sim[:, 84:], what="seg mask coefficients")
    assert_within_noise(proto, ys["proto"], ss["proto"], what="seg prototypes")
    assert_within_noise(proto, c["proto"], ss["proto"], what="seg prototypes vs reference golden")
