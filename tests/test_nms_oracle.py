"""NMS oracle pinned to the real reference function; CW-NMS oracle pinned to the reference's own C++ (`nms_and_cap`, common.cpp:127-205,
compiled as it lies by oracle/Makefile -> tests/golden/cwnms.golden.pt) and checked on hand-derived answers."""
import importlib.util
import math
import os

import numpy as np
import torch

from _util import GOLD
from oracle import nms_oracle as N

_spec = importlib.util.spec_from_file_location("make_golden_helpers", os.path.join(GOLD, "make_golden.py"))


def synth_predictions(*a, **k):
    # the generator lives in the golden script; import it without executing its reference-dependent imports
    src = open(os.path.join(GOLD, "make_golden.py")).read()
    start = src.index("def synth_predictions")
    end = src.index("def nms_golden")
    ns = {"torch": torch}
    exec(src[start:end], ns)
    return ns["synth_predictions"](*a, **k)


def test_nms_oracle_matches_reference():
    g = torch.load(os.path.join(GOLD, "nms.golden.pt"))
    for c in g["cases"]:
        pred = synth_predictions(c["B"], c["nc"], c["A"], c["seed"])
        out, keep = N.non_max_suppression(pred, c["conf"], c["iou"], max_det=c["max_det"])
        for o, k, ro, rk in zip(out, keep, c["out"], c["keep"]):
            assert torch.equal(k, rk)                       # kept anchor indices bit-exact, in order
            torch.testing.assert_close(o, ro, atol=0, rtol=0)


def test_cw_nms_known_answers():
    # two overlapping same-class boxes + one of another class at the same place + one far away
    boxes = np.array([[100, 100, 50, 50], [104, 102, 50, 50], [100, 100, 50, 50], [300, 300, 40, 40]], dtype=np.float32)
    scores = np.array([0.9, 0.6, 0.8, 0.5], dtype=np.float32)
    cls = np.array([0, 0, 1, 0])
    dets, kept = N.cw_nms(boxes, scores, cls, conf=0.25, iou_thr=0.5, sigma=0.1, max_det=300, frame_w=640, frame_h=640)
    assert kept == [0, 2, 3]                                # box 1 suppressed by box 0; class 1 untouched
    # hand computation for survivor 0: cluster = {0 (IoU 1), 1}
    iw, ih = 50 - 4, 50 - 2
    iou01 = iw * ih / (2500 + 2500 - iw * ih)
    w0, w1 = 0.9, float(np.float32(0.6)) * math.exp(-((1 - iou01) ** 2) / 0.1)
    x = (w0 * 100 + w1 * 104) / (w0 + w1)
    y = (w0 * 100 + w1 * 102) / (w0 + w1)
    np.testing.assert_allclose(dets[0, :4], [x, y, 50, 50], rtol=1e-6)
    np.testing.assert_allclose(dets[1, :4], [100, 100, 50, 50], rtol=1e-7)   # other class: its own cluster only
    np.testing.assert_allclose(dets[2, :4], [300, 300, 40, 40], rtol=1e-7)
    assert dets[:, 5].tolist() == [0, 1, 0]
    # without refinement the survivor set/order/scores are identical (common.cpp:148-149)
    d2, k2 = N.cw_nms(boxes, scores, cls, 0.25, 0.5, 0.1, 300, 640, 640, cluster=False)
    assert k2 == kept and np.array_equal(d2[:, 4:], dets[:, 4:])
    np.testing.assert_allclose(d2[0, :4], [100, 100, 50, 50])


def test_cw_nms_clip_cap_and_threshold_edges():
    boxes = np.array([[-10, -10, 30, 30], [630, 630, 30, 30], [700, 700, 10, 10], [50, 50, 10, 10]], dtype=np.float32)
    scores = np.array([0.5, 0.25, 0.9, 0.2499], dtype=np.float32)
    cls = np.zeros(4, dtype=int)
    dets, kept = N.cw_nms(boxes, scores, cls, conf=0.25, iou_thr=0.5, sigma=0.1, max_det=300, frame_w=640, frame_h=640)
    assert kept == [0, 1]                    # score == conf is kept (>=), 0.2499 is not, fully-outside box is dropped
    np.testing.assert_allclose(dets[0, :4], [0, 0, 20, 20])
    np.testing.assert_allclose(dets[1, :4], [630, 630, 10, 10])
    dets, kept = N.cw_nms(boxes, scores, cls, 0.25, 0.5, 0.1, 1, 640, 640)
    assert kept == [0]                       # max_det cap counts emitted detections


def _cw_oracle_image(pred_img, c):
    """oracle.cw_nms on one raw (4 + nc, A) prediction, candidates decoded like decode_candidates (common.cpp:93-125) with an
    identity letterbox: first-max class, top-left xywh in fp32."""
    p = pred_img
    sc, cl = p[4:].max(0)
    cx, cy, w, h = p[0], p[1], p[2], p[3]
    boxes = torch.stack([cx - 0.5 * w, cy - 0.5 * h, w, h], 1).numpy()
    return N.cw_nms(boxes, sc.numpy(), cl.numpy(), c["conf"], c["iou"], c["sigma"], c["max_det"], c["frame_w"], c["frame_h"],
                    cluster=bool(c["cluster"]))


def test_cw_nms_oracle_matches_reference_cpp_golden():
    """The restatement against the reference's own compiled `decode` = decode_candidates + nms_and_cap: same survivors in the same
    order (score / class columns exact), boxes equal to the last fp32 bit or one ulp of it (libm exp vs numpy exp in float64)."""
    g = torch.load(os.path.join(GOLD, "cwnms.golden.pt"))
    for c in g["cases"]:
        pred = synth_predictions(c["B"], c["nc"], c["A"], c["seed"], c["dense"])
        for b in range(c["B"]):
            dets, _ = _cw_oracle_image(pred[b], c)
            ref = c["dets"][b].numpy()
            assert dets.shape == ref.shape, (c["seed"], b, dets.shape, ref.shape)
            assert np.array_equal(dets[:, 4:], ref[:, 4:]), "survivor set / order differs from the reference C++"
            np.testing.assert_allclose(dets[:, :4], ref[:, :4], rtol=3e-7, atol=2e-5)


def test_cw_nms_reference_library_live():
    """When oracle/_ref/libcwnms_ref.so is present (build container, or shipped to the GPU box), the golden is what it produces NOW."""
    import pytest
    so = os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "libcwnms_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libcwnms_ref.so not built (make -C oracle needs /root/reference)")
    spec = importlib.util.spec_from_file_location("make_cwnms_golden", os.path.join(GOLD, "make_cwnms_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    f = mk.load_ref()
    g = torch.load(os.path.join(GOLD, "cwnms.golden.pt"))
    for c in g["cases"][:3]:
        pred = synth_predictions(c["B"], c["nc"], c["A"], c["seed"], c["dense"])
        for b in range(c["B"]):
            d = mk.reference_dets(f, pred[b], c["nc"], c["conf"], c["iou"], c["max_det"], c["cluster"], c["sigma"], c["frame_w"], c["frame_h"])
            assert torch.equal(d, c["dets"][b])
