"""GPU parity of the Segment / OBB post-processing: `ym_process_mask` and `ym_nms_rotated` through the host
mirrors `utils.ops.process_mask` / `utils.nms.non_max_suppression(rotated=True)` against the reference goldens
(tests/golden/postproc.golden.pt) and the CPU oracle, plus size-independent properties at full size (640 x 640 masks, 8400 and
33 600 anchors).  Both translation-unit kernels also run under the CUDA-on-host emulation
(tests/test_cuda_host_emu.py); on the B200 since round 2 (profiles/r02_gpu_suite.txt).

Tolerances: masks are bit-exact except pixels whose fp32 field (the value compared with 0) lies within 1e-4 of zero; rotated NMS is
exact (kept anchors, order, rows) on inputs whose closest ProbIoU-to-threshold distance exceeds 1e-5 (asserted on the oracle)."""
import os

import numpy as np
import pytest
import torch

from _util import GOLD
from oracle import postproc_oracle as PP
from yolo_master_b200.utils.nms import non_max_suppression
from yolo_master_b200.utils.ops import process_mask

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = torch.load(os.path.join(GOLD, "postproc.golden.pt"))


def _unpack(case, key):
    shape = tuple(case[key + "_shape"])
    return torch.from_numpy(np.unpackbits(case[key + "_bits"].numpy())[:int(np.prod(shape))].reshape(shape))


def _check_masks(got, want, field, keep, what, tol=1e-4):
    got = got.cpu()
    assert got.shape == want.shape and got.dtype == torch.uint8, what
    bad = got != want
    assert not (bad & ~keep).any(), (what, "mask pixels outside the crop")
    assert not (bad & (field.abs() > tol)).any(), (what, int(bad.sum()))
    assert bad.float().mean().item() < 1e-3, what


def test_process_mask_matches_reference_goldens():
    for i, c in enumerate(G["masks"]):
        n = c["coef"].shape[0]
        mh, mw = c["protos"].shape[1:]
        for key, up in (("up", True), ("native", False)):
            want = _unpack(c, key)
            for dt in (torch.float32, torch.float16):
                got = process_mask(c["protos"].to(DEV, dt), c["coef"].to(DEV), c["boxes"].to(DEV), c["shape"], upsample=up)
                if n == 0:
                    assert tuple(got.shape) == tuple(want.shape)
                    continue
                field = PP.mask_field(c["protos"], c["coef"], c["shape"], up)
                boxes = c["boxes"] if up else c["boxes"] * torch.tensor([[mw / c["shape"][1], mh / c["shape"][0]] * 2])
                _check_masks(got, want, field, PP.crop_keep(boxes, *field.shape[1:]), (i, key, dt))


def test_process_mask_full_size_properties():
    """100 detections x 640 x 640 from 32 x 160 x 160 prototypes: nothing outside the crop, sign symmetry (negating the coefficients
    flips every in-crop pixel whose field is clear of zero), and agreement with the oracle."""
    g = torch.Generator().manual_seed(11)
    protos = torch.nn.functional.avg_pool2d(torch.randn((1, 32, 160, 160), generator=g), 5, 1, 2)[0].half()
    n = 100
    coef = torch.randn((n, 32), generator=g)
    cxy = torch.rand((n, 2), generator=g) * 640
    wh = torch.rand((n, 2), generator=g) * 300 + 4
    boxes = torch.cat([cxy - wh / 2, cxy + wh / 2], 1)
    pos = process_mask(protos.to(DEV), coef.to(DEV), boxes.to(DEV), (640, 640), upsample=True)
    neg = process_mask(protos.to(DEV), (-coef).to(DEV), boxes.to(DEV), (640, 640), upsample=True)
    keep = PP.crop_keep(boxes, 640, 640)
    field = PP.mask_field(protos.float(), coef, (640, 640), True)
    _check_masks(pos, PP.process_mask(protos.float(), coef, boxes, (640, 640), True), field, keep, "full size")
    clear = keep & (field.abs() > 1e-4)
    assert ((pos.cpu() + neg.cpu())[clear] == 1).all()
    assert (pos.cpu()[~keep] == 0).all() and (neg.cpu()[~keep] == 0).all()


def test_rotated_nms_matches_reference_goldens():
    for i, c in enumerate(G["nms"]):
        _, _, margin = PP.non_max_suppression_rotated(c["pred"], c["conf"], c["iou"], c["max_det"], c["max_nms"])
        assert margin > 1e-5, (i, margin)
        out, keep = non_max_suppression(c["pred"].to(DEV), c["conf"], c["iou"], nc=c["nc"], max_det=c["max_det"], max_nms=c["max_nms"],
                                        rotated=True, return_idxs=True)
        for b in range(c["pred"].shape[0]):
            assert torch.equal(keep[b].cpu(), c["keep"][b]), (i, b)
            assert torch.equal(out[b].cpu(), c["out"][b]), (i, b)


@pytest.mark.parametrize("A", [8400, 33600])
def test_rotated_nms_full_size_properties(A):
    """Idempotence (the survivors survive a second pass unchanged), score order, and survivors pairwise below the threshold."""
    g = torch.Generator().manual_seed(A)
    B, nc = 2, 15
    pred = torch.zeros((B, 4 + nc + 1, A))
    pred[:, 0:2] = torch.rand((B, 2, A), generator=g) * 1024
    pred[:, 2:4] = torch.exp(torch.randn((B, 2, A), generator=g) * 0.5 + 3.5)
    pred[:, 4:4 + nc] = torch.rand((B, nc, A), generator=g) ** 4
    pred[:, -1] = (torch.rand((B, A), generator=g) - 0.25) * torch.pi
    out, keep = non_max_suppression(pred.to(DEV), 0.25, 0.45, nc=nc, max_det=300, rotated=True, return_idxs=True)
    for b in range(B):
        o, k = out[b].cpu(), keep[b].cpu()
        assert 0 < len(o) <= 300 and (o[:-1, 4] >= o[1:, 4]).all()
        assert torch.equal(o[:, :4], pred[b, :4, k].t()) and torch.equal(o[:, 6], pred[b, -1, k])
        boxes = torch.cat([o[:, :2] + o[:, 5:6] * 7680, o[:, 2:4], o[:, 6:7]], 1)
        iou = PP.batch_probiou(boxes, boxes).triu_(1)
        assert (iou < 0.45 + 1e-5).all()
        again = torch.zeros((1, 4 + nc + 1, len(o)))
        again[0, :4], again[0, -1] = o[:, :4].t(), o[:, 6]
        again[0, 4:4 + nc].scatter_(0, o[:, 5].long()[None], o[None, :, 4])
        o2, k2 = non_max_suppression(again.to(DEV), 0.25, 0.45, nc=nc, max_det=300, rotated=True, return_idxs=True)
        assert k2[0].cpu().tolist() == list(range(len(o))) and torch.equal(o2[0].cpu(), o)


def test_obb_predictor_on_the_reference_golden():
    """OBBPredictor.postprocess on the device: ym_nms_rotated + ym_scale_boxes(xywh) against the reference's rows + oracle rescale."""
    from oracle import letterbox_oracle as L
    from yolo_master_b200.engine import OBBPredictor

    class Stub(torch.nn.Module):
        def __init__(self, nc):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))
            self.stride, self.names = torch.tensor([8.0, 16.0, 32.0]), {i: str(i) for i in range(nc)}

    c = G["nms"][0]
    B = c["pred"].shape[0]
    pred = OBBPredictor(Stub(c["nc"]).to(DEV), imgsz=640, conf=c["conf"], iou=c["iou"], max_det=c["max_det"])
    frames = [np.zeros((480, 640, 3), np.uint8), np.zeros((360, 500, 3), np.uint8)][:B]
    res = pred.postprocess((c["pred"].to(DEV), {}), torch.zeros((B, 3, 640, 640), device=DEV), frames)
    for b, r in enumerate(res):
        ref = c["out"][b]
        want = torch.cat([torch.from_numpy(L.scale_boxes((640, 640), ref[:, :4].numpy(), frames[b].shape, xywh=True)), ref[:, 6:7], ref[:, 4:6]], 1)
        assert torch.equal(r.obb.data.cpu(), want), b


def test_scale_coords_on_device_bit_exact():
    """ym_scale_coords through utils.ops.scale_coords against the oracle (pinned to ops.scale_coords), (n, 17, 2 | 3) keypoints."""
    from oracle import letterbox_oracle as L
    from yolo_master_b200.utils.ops import scale_coords
    g = torch.Generator().manual_seed(6)
    for shape0 in ((480, 640), (1080, 1920), (100, 37)):
        for last in (2, 3):
            k = torch.rand((300, 17, last), generator=g) * 700 - 30
            for norm in (False, True):
                got = scale_coords((640, 640), k.clone().to(DEV), shape0, normalize=norm).cpu()
                assert np.array_equal(got.numpy(), L.scale_coords((640, 640), k.numpy(), shape0, normalize=norm)), (shape0, last, norm)
