"""GPU parity of the Pose head: `ym_kpts_decode` and the v0_1 pose model (`PoseModel`, ModularRouterExpertMoE
backbone) against the reference golden and the CPU oracle, plus NMS carrying the keypoint columns.

The kernel body is also checked under g++ in tests/test_preproc_host.py and the whole-model wiring on CPU emulation in
tests/test_host_model_wiring.py; on the B200 since round 2 (profiles/r02_gpu_suite.txt)."""
import os

import pytest
import torch

from _util import GOLD, assert_within_noise, synth_sd_from_keys, yaml_of
from oracle import nms_oracle as N
from oracle import yolo_master_oracle as O
from yolo_master_b200 import ops
from yolo_master_b200.nn.tasks import PoseModel
from yolo_master_b200.utils.nms import non_max_suppression
from yolo_master_b200.utils.synth import synth_images

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAME, CFG = "yolo-master-pose-n-v0_1", "master/v0_1/pose/yolo-master-pose-n.yaml"


def test_kpts_decode_kernel():
    g = torch.Generator().manual_seed(6)
    for ndim in (3, 2):
        nk, B, shapes, strides = 17 * ndim, 3, [(80, 80), (40, 40), (20, 20)], [8.0, 16.0, 32.0]
        levels = [torch.randn((B, h, w, nk), generator=g) for h, w in shapes]
        y = ops.kpts_decode([t.to(DEV) for t in levels], strides, ndim).cpu()
        raw = torch.cat([t.reshape(B, -1, nk).transpose(1, 2) for t in levels], 2)
        anchors, st = O.make_anchors(shapes, strides)
        want = raw.clone()
        if ndim == 3:
            want[:, 2::ndim] = want[:, 2::ndim].sigmoid()
        want[:, 0::ndim] = (raw[:, 0::ndim] * 2.0 + (anchors.t()[0] - 0.5)) * st.t()
        want[:, 1::ndim] = (raw[:, 1::ndim] * 2.0 + (anchors.t()[1] - 0.5)) * st.t()
        torch.testing.assert_close(y, want, atol=1e-5, rtol=1e-6)


@pytest.fixture(scope="module")
def model():
    m = PoseModel(CFG)
    sd = synth_sd_from_keys(0, NAME)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd, O.parse_spec(yaml_of(CFG))


def test_pose_model_matches_reference_golden(model):
    m, sd, spec = model
    assert m.kpt_shape == (17, 3)
    c = torch.load(os.path.join(GOLD, f"{NAME}.golden.pt"))["cases"]["b2_128"]
    x = synth_images(c["B"], c["H"], c["W"], c["seed"]).half()
    with torch.no_grad():
        y = m(x.to(DEV))[0].float().cpu()
    ref = O.forward(spec, sd, x.float())
    with O.fp16_storage(), O.fp16_weights():
        sim = O.forward(spec, sd, x.float())
    assert y.shape == ref.shape == (2, 5 + 51, 336)
    assert_within_noise(y[:, :4], ref[:, :4], sim[:, :4], what="pose boxes")
    assert_within_noise(y[:, 4:5], ref[:, 4:5], sim[:, 4:5], what="pose scores")
    assert_within_noise(y[:, 5:], ref[:, 5:], sim[:, 5:], what="pose keypoints")
    assert_within_noise(y, c["final"].float(), sim, what="pose vs reference golden")


def test_nms_carries_keypoints(model):
    """non_max_suppression(nc=1) on the (B, 4 + 1 + 51, A) prediction: kept anchors equal the oracle's on the box / score rows, and
    every output row is [box, conf, cls, the anchor's 51 keypoint values]."""
    m, _, _ = model
    x = synth_images(3, 256, 256, 11).half().to(DEV)
    with torch.no_grad():
        y = m(x)[0]
    conf = float(y[:, 4].flatten().kthvalue(int(0.9 * y.shape[0] * y.shape[2]))[0])
    out, keep = non_max_suppression(y, conf, 0.6, max_det=100, nc=1, return_idxs=True)
    ro, rk = N.non_max_suppression(y[:, :5].float().cpu(), conf, 0.6, max_det=100)
    for b, (o, k, a, kk) in enumerate(zip(out, keep, ro, rk)):
        assert torch.equal(k.cpu(), kk) and o.shape[1] == 6 + 51
        torch.testing.assert_close(o[:, :6].cpu(), a, atol=1e-4, rtol=1e-5)
        assert torch.equal(o[:, 6:], y[b, 5:, k].t().float())
