"""GPU parity of the predictor's pre-/post-processing kernels: `ym_letterbox_u8` and
`ym_scale_boxes` through the public mirror (LetterBox, DetectionPredictor, utils.ops.scale_boxes), bit-exact against the oracle
(pinned to the real LetterBox + cv2 and ops.scale_boxes by tests/test_letterbox_oracle.py) and the reference goldens.

Their arithmetic is also verified on the host (tests/test_preproc_host.py compiles the same per-pixel code with g++); the CUDA launches
run on the B200 since round 2 (profiles/r02_gpu_suite.txt)."""
import os
import zlib

import numpy as np
import pytest
import torch

from _util import GOLD, synth_sd_from_keys
from oracle import letterbox_oracle as L
from yolo_master_b200.data.augment import LetterBox
from yolo_master_b200.engine import DetectionPredictor
from yolo_master_b200.nn.tasks import DetectionModel
from yolo_master_b200.utils import ops as box_ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = torch.load(os.path.join(GOLD, "letterbox.golden.pt"))


def _frame(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"{c['h']}x{c['w']}")
def test_letterbox_kernel_matches_reference_golden(case):
    img = _frame(case["seed"], case["h"], case["w"])
    out = LetterBox((640, 640)).apply_batch(torch.from_numpy(img)[None].to(DEV), swap_rb=True, chw=True)[0].cpu().numpy()
    assert np.array_equal(out, L.preprocess_frame(img))
    assert zlib.crc32(out.tobytes()) == case["crc"]


@pytest.mark.parametrize("case", GOLDEN["variants"], ids=lambda c: f"{c['h']}x{c['w']}-{'-'.join(c['kw']) or 'rect'}")
def test_letterbox_class_matches_reference_golden(case):
    """`LetterBox(...)(image=frame)`: HWC BGR output of the transform for auto / scaleup / scale_fill / center / non-square."""
    img = _frame(case["seed"], case["h"], case["w"])
    out = LetterBox(tuple(case["new_shape"]), stride=32, **case["kw"])(image=img).cpu().numpy()
    assert list(out.shape) == case["shape"] and zlib.crc32(out.tobytes()) == case["crc"]


def test_letterbox_kernel_shapes_dtypes_batches():
    """Exact 2x area path, identity, odd widths (scalar store path), degenerate sources; fp16 / fp32 scaling; B > 1."""
    rng = np.random.default_rng(3)
    for (h, w, new_shape) in ((1280, 1280, (640, 640)), (640, 640, (640, 640)), (1, 1, (64, 64)), (20, 900, (96, 128)), (33, 7, (70, 50)),
                              (250, 250, (125, 125)), (251, 250, (125, 125)), (90, 160, (101, 203)), (720, 1280, (736, 1280))):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = L.preprocess_frame(img, new_shape)
        lb = LetterBox(new_shape)
        dev = torch.from_numpy(img)[None].to(DEV)
        assert np.array_equal(lb.apply_batch(dev, swap_rb=True, chw=True)[0].cpu().numpy(), want), (h, w, new_shape)
        assert np.array_equal(lb.apply_batch(dev)[0].cpu().numpy(), L.letterbox_frame(img, new_shape)), (h, w, new_shape)
        u8 = torch.from_numpy(want).to(DEV)
        assert torch.equal(lb.apply_batch(dev, swap_rb=True, chw=True, dtype=torch.float16)[0], u8.half() / 255)
        # fp32: the kernel divides exactly (x / 255 correctly rounded = the reference's CPU `im.float() / 255`); torch's CUDA scalar division
        # multiplies by the reciprocal instead (1 ulp apart on some values), so the exact reference is computed on the host
        assert torch.equal(lb.apply_batch(dev, swap_rb=True, chw=True, dtype=torch.float32)[0].cpu(), u8.cpu().float() / 255)
    frames = np.stack([_frame(20 + i, 360, 500) for i in range(5)])
    out = LetterBox((640, 640)).apply_batch(torch.from_numpy(frames).to(DEV), swap_rb=True, chw=True).cpu().numpy()
    for i in range(5):
        assert np.array_equal(out[i], L.preprocess_frame(frames[i])), i


@pytest.mark.parametrize("case", GOLDEN["scale_boxes"], ids=lambda c: f"{c['img0'][0]}x{c['img0'][1]}-{'xywh' if c['xywh'] else 'xyxy'}")
def test_scale_boxes_kernel_matches_reference_golden(case):
    g = torch.Generator().manual_seed(case["seed"])
    img1 = case["img1"]
    b = torch.rand((64, 6), generator=g) * torch.tensor([img1[1], img1[0], img1[1], img1[0], 1, 80]) * 1.1 - 8.0
    rows = b.to(DEV)
    box_ops.scale_boxes(img1, rows[:, :4], case["img0"], xywh=case["xywh"])           # strided view, in place
    assert torch.equal(rows[:, :4].cpu(), case["out"]) and torch.equal(rows[:, 4:].cpu(), b[:, 4:])


def test_scale_boxes_batch_and_ragged():
    g = torch.Generator().manual_seed(4)
    shapes = [(480, 640), (1080, 1920), (100, 37), (333, 500)] * 40                   # 160 images: two parameter blocks
    b = torch.rand((len(shapes), 300, 6), generator=g) * 700 - 30
    dev = b.clone().to(DEV)
    box_ops.scale_boxes_batch((640, 640), dev, shapes)
    for i, s in enumerate(shapes):
        assert np.array_equal(dev[i, :, :4].cpu().numpy(), L.scale_boxes((640, 640), b[i, :, :4].numpy(), s)), i
    assert torch.equal(dev[:, :, 4:].cpu(), b[:, :, 4:])
    counts = [7, 0, 31, 2]
    flat = torch.rand((sum(counts), 6), generator=g) * 700 - 30
    row_img = torch.repeat_interleave(torch.arange(4, dtype=torch.int32), torch.tensor(counts)).to(DEV)
    dev = flat.clone().to(DEV)
    box_ops.scale_boxes_batch((640, 640), dev, shapes[:4], row_img=row_img)
    lo = 0
    for i, n in enumerate(counts):
        assert np.array_equal(dev[lo:lo + n, :4].cpu().numpy(), L.scale_boxes((640, 640), flat[lo:lo + n, :4].numpy(), shapes[i])), i
        lo += n
    clipped = box_ops.clip_boxes(torch.tensor([[-5.0, 3.0, 700.0, 500.0, 0.5, 1.0]], device=DEV), (480, 640))
    assert clipped.cpu().tolist() == [[0.0, 3.0, 640.0, 480.0, 0.5, 1.0]]


def test_predictor_end_to_end():
    """frames -> DetectionPredictor (device letterbox -> model -> confidence filter -> device rescale) equals the model run on
    the oracle's pre-processed batch followed by the oracle's rescale: same kernels on identical bytes, so exactly equal."""
    m = DetectionModel("yolo26-master-n.yaml")
    m.load_state_dict(synth_sd_from_keys(0), strict=True)
    m = m.to(DEV).eval()
    shapes = [(480, 640), (720, 1280), (480, 640), (333, 500)]
    frames = [_frame(100 + i, h, w) for i, (h, w) in enumerate(shapes)]
    pred = DetectionPredictor(m, imgsz=640, conf=0.0)
    results = pred(frames)
    batch = torch.from_numpy(np.stack([L.preprocess_frame(f) for f in frames])).to(DEV)
    assert torch.equal(pred.preprocess(frames), batch)
    with torch.no_grad():
        y = m(batch)[0].float().cpu()
    for i, (r, f) in enumerate(zip(results, frames)):
        keep = y[i][y[i][:, 4] > 0.0]
        want = keep.clone()
        want[:, :4] = torch.from_numpy(L.scale_boxes((640, 640), keep[:, :4].numpy(), f.shape))
        assert torch.equal(r.boxes.data.cpu(), want), i
        assert r.orig_shape == f.shape[:2]
