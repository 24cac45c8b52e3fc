"""Pre-/post-processing without a GPU.

1. The per-pixel / per-box code the kernels run (yolo-master_b200/csrc/preproc_core.cuh) is compiled for the HOST with g++
   (tests/native/preproc_host.cpp) and compared bit-for-bit with the oracle and the reference goldens: this checks the integer
   arithmetic and the host-built tap tables, everything except the CUDA thread indexing (tests/test_gpu_zz_predictor.py).
2. The host mirror (LetterBox.get_params, DetectionPredictor.preprocess / postprocess, utils.ops.scale_boxes) is run on CPU
   tensors with `ops.letterbox` / `ops.scale_boxes` replaced by that host build, against the oracle."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest
import torch

from _util import GOLD, ROOT
from oracle import letterbox_oracle as L
from yolo_master_b200 import ops
from yolo_master_b200.data.augment import LetterBox, _axis_taps
from yolo_master_b200.engine import DetectionPredictor, Results
from yolo_master_b200.utils import ops as box_ops

GOLDEN = torch.load(os.path.join(GOLD, "letterbox.golden.pt"))
vp, ci, cll = C.c_void_p, C.c_int, C.c_longlong


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("preproc_host") / "libpreproc_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "yolo-master_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "preproc_host.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.host_letterbox_u8.argtypes = [vp, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, ci, ci, ci]
    lib.host_scale_boxes.argtypes = [vp, ci, cll, ci, vp, vp, ci, ci]
    lib.host_kpts_decode_level.argtypes = [vp, ci, ci, C.c_float, ci, ci, ci, ci, ci, vp]
    lib.host_obb_finish_level.argtypes = [vp, ci, ci, C.c_float, ci, ci, ci, ci, vp, vp]
    lib.host_letterbox_u8.restype = lib.host_scale_boxes.restype = lib.host_kpts_decode_level.restype = lib.host_obb_finish_level.restype = None
    return lib


def _emu_letterbox(lib):
    def letterbox(src, xtab, ytab, area2x, nw, nh, top, left, H, W, pad_value=114, swap_rb=True, chw=True, dtype=torch.uint8, out=None):
        B, sh, sw, _ = src.shape
        assert src.dtype == torch.uint8 and src.is_contiguous()
        res = torch.empty((B, 3, H, W) if chw else (B, H, W, 3), dtype=torch.uint8)
        for b in range(B):
            lib.host_letterbox_u8(src[b].data_ptr(), sh, sw, 3 * sw, None if area2x else xtab.data_ptr(),
                                  None if area2x else ytab.data_ptr(), int(area2x), nw, nh, top, left, pad_value, int(swap_rb),
                                  res[b].data_ptr(), int(chw), H, W)
        if dtype != torch.uint8:
            res = res.to(dtype) / 255       # im.half() / 255 resp. im.float() / 255, predictor.py:173-175
        if out is not None:
            out.copy_(res)
            return out
        return res
    return letterbox


def _emu_scale_boxes(lib):
    def scale_boxes(boxes, params, rows_per_img=0, row_img=None, padding=True, xywh=False):
        assert boxes.dtype == torch.float32 and params.dtype == torch.float32 and params.shape[1] == 5
        if boxes.dim() == 2 and boxes.stride(1) == 1:
            ld = boxes.stride(0) if boxes.shape[0] > 1 else boxes.shape[1]
        else:
            assert boxes.is_contiguous()
            ld = boxes.shape[-1]
        n = boxes.numel() // boxes.shape[-1]
        assert row_img is not None or n <= rows_per_img * params.shape[0]
        lib.host_scale_boxes(boxes.data_ptr(), ld, n, rows_per_img, None if row_img is None else row_img.data_ptr(),
                             params.data_ptr(), int(padding), int(xywh))
        return boxes
    return scale_boxes


@pytest.fixture()
def emu(host, monkeypatch):
    monkeypatch.setattr(ops, "letterbox", _emu_letterbox(host))
    monkeypatch.setattr(ops, "scale_boxes", _emu_scale_boxes(host))
    return host


def _frame(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


# ------------------------------------------------------------------------------------------------ 1. kernel arithmetic on the host
def test_axis_taps_pack_the_oracle_coefficients():
    for dn, sn in ((640, 64), (237, 37), (426, 333), (640, 480), (360, 1080), (9, 2), (5, 1), (640, 640)):
        for clamp in (True, False):
            i0, i1, a0, a1 = L._coeffs(dn, sn, 1.0 / (dn / sn), clamp)
            t = _axis_taps(dn, sn, clamp).astype(np.int64)
            assert np.array_equal(t[:, 0] & 0xFFFF, i0) and np.array_equal(t[:, 0] >> 16, i1), (dn, sn, clamp)
            assert np.array_equal(t[:, 1] & 0xFFFF, a0) and np.array_equal(t[:, 1] >> 16, a1), (dn, sn, clamp)
            assert np.all((t[:, 1] & 0xFFFF) + (t[:, 1] >> 16) == 2048)


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"{c['h']}x{c['w']}")
def test_kernel_arithmetic_matches_reference_golden(emu, case):
    """LetterBox.plan tables + the kernel's per-pixel code (host build) against the REAL LetterBox + cv2 pipeline (CRC)."""
    img = _frame(case["seed"], case["h"], case["w"])
    out = LetterBox((640, 640)).apply_batch(torch.from_numpy(img)[None], swap_rb=True, chw=True)[0].numpy()
    assert list(out.shape) == case["shape"]
    assert np.array_equal(out, L.preprocess_frame(img))
    assert zlib.crc32(out.tobytes()) == case["crc"]


@pytest.mark.parametrize("case", GOLDEN["variants"], ids=lambda c: f"{c['h']}x{c['w']}-{'-'.join(c['kw']) or 'rect'}")
def test_letterbox_class_variants_match_reference_golden(emu, case):
    """auto / scaleup=False / scale_fill / center=False / non-square: get_params and the HWC output of `LetterBox(image=...)`."""
    img = _frame(case["seed"], case["h"], case["w"])
    lb = LetterBox(tuple(case["new_shape"]), stride=32, **case["kw"])
    prm = lb.get_params({"img": img})
    assert [list(prm["new_unpad"]), prm["top"], prm["bottom"], prm["left"], prm["right"]] == case["params"]
    out = lb.apply_batch(torch.from_numpy(img)[None])[0].numpy()
    assert list(out.shape) == case["shape"] and zlib.crc32(out.tobytes()) == case["crc"]


def test_kernel_arithmetic_random_shapes(emu):
    """Up / down / mixed / exact-2x / identity / degenerate sources, odd target widths, against the oracle."""
    rng = np.random.default_rng(3)
    for (h, w, new_shape) in ((1280, 1280, (640, 640)), (640, 640, (640, 640)), (1, 1, (64, 64)), (20, 900, (96, 128)), (33, 7, (70, 50)),
                              (250, 250, (125, 125)), (251, 250, (125, 125)), (90, 160, (101, 203)), (720, 1280, (736, 1280))):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out = LetterBox(new_shape).apply_batch(torch.from_numpy(img)[None], swap_rb=True, chw=True)[0].numpy()
        assert np.array_equal(out, L.preprocess_frame(img, new_shape)), (h, w, new_shape)


@pytest.mark.parametrize("case", GOLDEN["scale_boxes"], ids=lambda c: f"{c['img0'][0]}x{c['img0'][1]}-{'xywh' if c['xywh'] else 'xyxy'}")
def test_scale_boxes_matches_reference_golden(emu, case):
    g = torch.Generator().manual_seed(case["seed"])
    img1 = case["img1"]
    b = torch.rand((64, 6), generator=g) * torch.tensor([img1[1], img1[0], img1[1], img1[0], 1, 80]) * 1.1 - 8.0
    rows = b.clone()
    out = box_ops.scale_boxes(img1, rows[:, :4], case["img0"], xywh=case["xywh"])     # the strided [:, :4] view, in place
    assert out.data_ptr() == rows.data_ptr() and torch.equal(rows[:, :4], case["out"]) and torch.equal(rows[:, 4:], b[:, 4:])


def test_clip_boxes_and_empty(emu):
    b = torch.tensor([[-5.0, 3.0, 700.0, 500.0, 0.5, 1.0], [10.0, -1.0, 20.0, 479.5, 0.25, 3.0]])
    box_ops.clip_boxes(b, (480, 640))
    assert b.tolist() == [[0.0, 3.0, 640.0, 480.0, 0.5, 1.0], [10.0, 0.0, 20.0, 479.5, 0.25, 3.0]]
    e = torch.zeros((0, 6))
    assert box_ops.scale_boxes((640, 640), e, (480, 640)) is e


# ------------------------------------------------------------------------------------------------ 2. host mirror of the predictor
class _StubModel(torch.nn.Module):
    """Records what the predictor feeds it and returns a fixed end2end-style (B, K, 6) prediction."""

    def __init__(self, preds, end2end=True):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.preds, self.end2end, self.stride, self.names, self.seen = preds, end2end, torch.tensor([8.0, 16.0, 32.0]), {0: "a"}, None

    def forward(self, x):
        self.seen = x
        return self.preds.clone(), {}


def _preds(B, K, seed, wh=640):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand((B, K, 2), generator=g) * wh
    p = torch.cat([xy, xy + torch.rand((B, K, 2), generator=g) * 200, torch.rand((B, K, 1), generator=g),
                   torch.randint(0, 80, (B, K, 1), generator=g).float()], 2)
    return p


def test_predictor_preprocess_mixed_shapes(emu):
    """Frames of three shapes, interleaved (groups are not contiguous runs): every slot equals the reference pipeline's frame."""
    shapes = [(480, 640), (720, 1280), (480, 640), (100, 37), (720, 1280)]
    frames = [_frame(40 + i, h, w) for i, (h, w) in enumerate(shapes)]
    pred = DetectionPredictor(_StubModel(_preds(5, 8, 0)), imgsz=640, device="cpu")
    im = pred.preprocess(frames)
    assert im.dtype == torch.uint8 and tuple(im.shape) == (5, 3, 640, 640)
    for i, f in enumerate(frames):
        assert np.array_equal(im[i].numpy(), L.preprocess_frame(f)), i
    half = DetectionPredictor(_StubModel(_preds(5, 8, 0)), imgsz=640, half=True, device="cpu").preprocess(frames)
    ref = torch.from_numpy(np.stack([L.preprocess_frame(f) for f in frames])).half()
    ref /= 255                                                                     # predictor.py:173-175
    assert half.dtype == torch.float16 and torch.equal(half, ref)


def test_predictor_rect_and_pre_transform(emu):
    """rect=True with same-sized frames -> minimum rectangle (auto), as predictor.py:195-203; pre_transform returns BGR HWC."""
    frames = [_frame(60 + i, 300, 400) for i in range(2)]
    pred = DetectionPredictor(_StubModel(_preds(2, 4, 1)), imgsz=640, rect=True, device="cpu")
    im = pred.preprocess(frames)
    assert tuple(im.shape) == (2, 3, 480, 640)
    for i, f in enumerate(frames):
        assert np.array_equal(im[i].numpy(), L.preprocess_frame(f, (640, 640), auto=True))
    monkey_cuda = torch.Tensor.cuda
    try:
        torch.Tensor.cuda = lambda self, *a, **k: self          # LetterBox.__call__ moves the frame to the GPU
        hwc = pred.pre_transform(frames)
    finally:
        torch.Tensor.cuda = monkey_cuda
    assert np.array_equal(hwc[0].numpy(), L.letterbox_frame(frames[0], (640, 640), auto=True))
    with pytest.raises(ValueError):
        DetectionPredictor(_StubModel(_preds(2, 4, 1)), rect=True, device="cpu").preprocess([_frame(1, 300, 400), _frame(2, 64, 64, )][:1] + [np.zeros((4, 4))])


def test_predictor_postprocess_end2end(emu):
    """(B, 300, 6) end2end predictions: confidence filter, then every image's boxes rescaled to ITS original frame."""
    shapes = [(480, 640), (1080, 1920), (100, 37)]
    frames = [_frame(80 + i, h, w) for i, (h, w) in enumerate(shapes)]
    preds = _preds(3, 300, 5)
    preds[1, :, 4] = 0.0                                       # image 1: nothing above the threshold
    model = _StubModel(preds)
    pred = DetectionPredictor(model, imgsz=640, conf=0.25, device="cpu")
    results = pred(frames, paths=["a.jpg", "b.jpg", "c.jpg"])
    assert np.array_equal(model.seen[2].numpy(), L.preprocess_frame(frames[2]))
    assert [type(r) for r in results] == [Results] * 3 and [r.path for r in results] == ["a.jpg", "b.jpg", "c.jpg"]
    assert len(results[1]) == 0 and results[1].boxes.data.shape == (0, 6)
    for b, (r, f) in enumerate(zip(results, frames)):
        keep = preds[b][preds[b, :, 4] > 0.25]
        want = keep.clone()
        want[:, :4] = torch.from_numpy(L.scale_boxes((640, 640), keep[:, :4].numpy(), f.shape))
        assert torch.equal(r.boxes.data, want), b
        assert r.orig_shape == f.shape[:2] and r.orig_img is f
        if len(r):
            assert float(r.boxes.xyxy[:, [0, 2]].max()) <= f.shape[1] and float(r.boxes.xyxy[:, [1, 3]].max()) <= f.shape[0]
            assert r.summary()[0]["class"] == int(r.boxes.cls[0])
    single = pred.construct_result(preds[0][:5].clone(), model.seen, frames[0], "x.jpg")
    want = torch.from_numpy(L.scale_boxes((640, 640), preds[0][:5, :4].numpy(), frames[0].shape))
    assert torch.equal(single.boxes.xyxy, want)


def test_results_boxes_accessors():
    data = torch.tensor([[10.0, 20.0, 110.0, 220.0, 0.9, 3.0], [0.0, 0.0, 50.0, 40.0, 0.4, 1.0]])
    r = Results(np.zeros((400, 200, 3), np.uint8), path="p", names={1: "one", 3: "three"}, boxes=data)
    assert r.boxes.xywh.tolist()[0] == [60.0, 120.0, 100.0, 200.0]
    assert r.boxes.xyxyn.tolist()[0] == pytest.approx([0.05, 0.05, 0.55, 0.55])
    assert r.boxes.conf.tolist() == pytest.approx([0.9, 0.4]) and r.boxes.cls.tolist() == [3.0, 1.0]
    assert len(r[0]) == 1 and r.numpy().boxes.data.shape == (2, 6)
    assert r.summary()[0]["name"] == "three" and r.summary(normalize=True)[1]["box"]["x2"] == 0.25 and r.summary()[1]["confidence"] == 0.4


def test_letterbox_rejects_unsupported():
    with pytest.raises(NotImplementedError):
        LetterBox(interpolation=3)
    with pytest.raises(NotImplementedError):
        LetterBox()(labels={"img": np.zeros((4, 4, 3), np.uint8), "instances": object()})
    with pytest.raises(ValueError):
        LetterBox()(image=np.zeros((4, 4), np.uint8))
    with pytest.raises(ValueError):
        LetterBox((96, 128)).plan((2, 900), "cpu")           # would resize to zero rows (cv2.resize raises there too)


def test_kpts_decode_body_matches_reference_formula(host):
    """`ym_kpts_decode`'s per-element function (g++ build) against Pose.kpts_decode head.py:644-664 written with torch, for
    (x, y, visibility) and (x, y) keypoints over three levels."""
    from oracle import yolo_master_oracle as O
    g = torch.Generator().manual_seed(6)
    for ndim in (3, 2):
        nk, B, shapes, strides = 17 * ndim, 2, [(8, 6), (4, 3), (2, 2)], [8.0, 16.0, 32.0]
        A = sum(h * w for h, w in shapes)
        levels = [torch.randn((B, h, w, nk), generator=g) for h, w in shapes]
        y = torch.empty((B, nk, A))
        a0 = 0
        for t, (h, w), s in zip(levels, shapes, strides):
            host.host_kpts_decode_level(t.data_ptr(), h, w, s, a0, B, nk, ndim, A, y.data_ptr())
            a0 += h * w
        raw = torch.cat([t.reshape(B, -1, nk).transpose(1, 2) for t in levels], 2)
        anchors, st = O.make_anchors(shapes, strides)
        want = raw.clone()
        if ndim == 3:
            want[:, 2::ndim] = want[:, 2::ndim].sigmoid()
        want[:, 0::ndim] = (raw[:, 0::ndim] * 2.0 + (anchors.t()[0] - 0.5)) * st.t()
        want[:, 1::ndim] = (raw[:, 1::ndim] * 2.0 + (anchors.t()[1] - 0.5)) * st.t()
        torch.testing.assert_close(y, want, atol=1e-6, rtol=1e-6)


def test_obb_finish_body_matches_dist2rbox(host):
    """`ym_obb_finish`'s per-anchor function (g++ build) applied to an axis-aligned xywh decode against dist2rbox (utils/tal.py:447-453)
    evaluated directly on the distances: rotated centres agree to fp32 rounding, w / h / scores are copied, the angle row is appended."""
    import math
    from oracle import yolo_master_oracle as O
    g = torch.Generator().manual_seed(8)
    B, nc, shapes, strides = 2, 3, [(6, 5), (3, 3)], [8.0, 16.0]
    A = sum(h * w for h, w in shapes)
    dist = torch.rand((B, 4, A), generator=g) * 3                                   # l, t, r, b in grid units
    scores = torch.rand((B, nc, A), generator=g)
    raw = [torch.randn((B, h, w, 1), generator=g) for h, w in shapes]
    anchors, st = O.make_anchors(shapes, strides)
    anchors, st = anchors.t().unsqueeze(0), st.t()
    lt, rb = dist.split(2, 1)
    yin = torch.cat([((anchors - lt) + (anchors + rb)) / 2 * st, (lt + rb) * st, scores], 1).contiguous()   # what ym_detect_dense emits
    yout = torch.empty((B, 4 + nc + 1, A))
    a0 = 0
    for t, (h, w), s in zip(raw, shapes, strides):
        host.host_obb_finish_level(t.data_ptr(), h, w, s, a0, B, nc, A, yin.data_ptr(), yout.data_ptr())
        a0 += h * w
    angle = (torch.cat([t.reshape(B, -1, 1).transpose(1, 2) for t in raw], 2).sigmoid() - 0.25) * math.pi
    xf, yf = ((rb - lt) / 2).split(1, 1)
    xy = (torch.cat([xf * angle.cos() - yf * angle.sin(), xf * angle.sin() + yf * angle.cos()], 1) + anchors) * st
    torch.testing.assert_close(yout[:, :2], xy, atol=2e-4, rtol=1e-5)
    assert torch.equal(yout[:, 2:4 + nc], yin[:, 2:])
    torch.testing.assert_close(yout[:, -1:], angle, atol=1e-6, rtol=1e-6)
