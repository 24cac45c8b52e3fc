"""BASELINE.json configs[2] / configs[3] at their FULL sizes on the GPU, checked through size-independent properties (the CPU
oracle needs minutes per batch at these sizes; oracle parity of the same models is in test_gpu_model_{v0,mot}.py at sizes it
finishes in seconds):
  * per-image independence: image i of a batch gives bit-identical output to the same image run alone / in another batch slot
    (routing, GroupNorm, LayerNorm, top-k and NMS are per image - SURVEY.md §8e: this is what makes batch sharding exact);
  * determinism: two runs are bit-identical (no atomics on the data path);
  * CUDA-graph replay == eager; outputs finite; boxes inside a sane range; NMS / CW-NMS post-processing runs on the result.
"""
import pytest
import torch

from _util import synth_sd_from_keys
from yolo_master_b200.nn.tasks import DetectionModel, yaml_model_load
from yolo_master_b200.utils.nms import non_max_suppression
from yolo_master_b200.utils.synth import synth_images

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(name, cfg):
    m = DetectionModel(cfg)
    m.load_state_dict(synth_sd_from_keys(0, name), strict=True)
    return m.to(DEV).eval()


def test_config2_s_mot_moa_bs64_640_cwnms():
    """configs[2]: YOLO-Master-S, MoT + MoA enabled, bs64 @ 640x640, CW-NMS on (non-end2end head)."""
    d = yaml_model_load("yolo26-master-moa-mot-n.yaml")
    d["scales"]["s"] = [0.50, 0.50, 1024]
    d["scale"] = "s"
    m = _model("yolo26-master-moa-mot-s", d)
    x = synth_images(64, 640, 640, 77).half().to(DEV)
    with torch.no_grad():
        y = m(x)[0]
        y2 = m(x)[0]
        alone = m(x[5:6])[0]
        shuffled = m(torch.cat([x[40:], x[:40]]))[0]
    assert y.shape == (64, 300, 6) and torch.isfinite(y).all()
    assert torch.equal(y, y2)                                   # deterministic
    assert torch.equal(y[5], alone[0])                          # per-image independence
    assert torch.equal(y[3], shuffled[27]) and torch.equal(y[63], shuffled[23])
    # "CW-NMS on": one2many head -> dense (B, 84, 8400) -> CW-NMS
    m.end2end = False
    with torch.no_grad():
        dense = m(x[:8])[0]
    assert dense.shape == (8, 84, 8400) and torch.isfinite(dense).all()
    conf = float(dense[:, 4:].amax(1).flatten().kthvalue(int(0.98 * 8 * 8400))[0])
    plain, keep = non_max_suppression(dense, conf, 0.7, max_det=300, return_idxs=True)
    cw, keep_cw = non_max_suppression(dense, conf, 0.7, max_det=300, return_idxs=True, cluster=True, frame_wh=(640, 640))
    assert all(len(k) > 0 for k in keep)
    for a, b, o in zip(keep, keep_cw, cw):
        # same greedy per-class suppression in both modes; they differ only in the class offset (7680 vs 2*max(w,h)+8192,
        # common.cpp:138) and the IoU precision (fp32 vs double), which can flip pairs sitting exactly at the threshold
        sa, sb = set(a.tolist()), set(b.tolist())
        assert len(sa & sb) >= 0.9 * max(len(sa), len(sb))
        assert bool((o[:, 2:4] > 0).all()) and bool((o[:, :2] >= 0).all())
    m.end2end = True


def test_config3_l_1280_shard():
    """configs[3]: YOLO-Master-L @ 1280x1280, one rank's shard of the 8-GPU batch (16 images per GPU; 4 here to bound test time),
    dense DFL output (B, 84, 33600) + NMS."""
    m = _model("yolo-master-l-v0", "master/v0/det/yolo-master-l.yaml")
    x = synth_images(4, 1280, 1280, 78).half().to(DEV)
    with torch.no_grad():
        y = m(x)[0]
        y2 = m(x)[0]
        alone = m(x[2:3])[0]
    assert y.shape == (4, 84, 33600) and torch.isfinite(y).all()
    assert torch.equal(y, y2) and torch.equal(y[2], alone[0])
    assert float(y[:, :2].min()) > -200 and float(y[:, :2].max()) < 1480      # box centres near the 1280 frame
    assert float(y[:, 4:].min()) >= 0 and float(y[:, 4:].max()) <= 1
    g = m.graphed(4, 1280, 1280)
    out = g(x).clone()
    torch.cuda.synchronize()
    assert torch.equal(out, y)
    conf = float(y[:, 4:].amax(1).flatten().kthvalue(int(0.99 * 4 * 33600))[0])
    dets = non_max_suppression(y, conf, 0.7, max_det=300)
    assert all(0 < len(d) <= 300 for d in dets)
