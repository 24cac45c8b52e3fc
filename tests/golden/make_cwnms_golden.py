"""Generate the CW-NMS golden from the reference's OWN C++ (run in the build container; needs oracle/_ref/libcwnms_ref.so, i.e.
`make -C oracle` where /root/reference exists).

    python tests/golden/make_cwnms_golden.py        -> tests/golden/cwnms.golden.pt

The library is the reference's examples/YOLO-Master-Cross-Platform-Edge-Deployment/cpp/src/common.cpp compiled as it lies
(oracle/Makefile) - `decode` = decode_candidates (:93-125) + nms_and_cap (:127-205) - behind oracle/cwnms_ref_wrap.cpp.
Each case stores the generator arguments of `synth_predictions` (make_golden.py) and the reference's detections
(x, y, w, h, conf, cls) per image, so the fixture is a few hundred KB.
"""
import ctypes
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_nms_oracle import synth_predictions  # noqa: E402  (the seeded generator shared with nms.golden.pt)

# (B, nc, A, seed, dense, conf, iou, sigma, cluster, max_det, frame_w, frame_h)
CASES = [
    (2, 80, 2100, 11, True, 0.25, 0.50, 0.1, 1, 300, 640, 640),
    (1, 20, 8400, 12, True, 0.30, 0.60, 0.5, 1, 300, 640, 640),
    (2, 1, 6000, 13, True, 0.01, 0.45, 0.1, 1, 300, 640, 640),       # one class, > 3000 candidates: the top-3000 pool cap (:153-158)
    (1, 3, 12000, 14, True, 0.02, 0.50, 0.05, 1, 100, 640, 640),     # dense clusters, small sigma, max_det cap
    (2, 80, 2100, 15, False, 0.25, 0.50, 0.1, 0, 300, 640, 640),     # Standard mode: survivors only
    (2, 10, 3000, 16, True, 0.20, 0.40, 0.2, 1, 300, 480, 320),      # frame smaller than the boxes' range: clip / drop (:187-197)
    (1, 80, 64, 17, True, 0.999, 0.5, 0.1, 1, 300, 640, 640),        # nothing above conf
]


def load_ref():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcwnms_ref.so"))
    f = lib.cwnms_ref_decode_nms
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                  ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    return f


def reference_dets(f, pred_img, nc, conf, iou, max_det, cluster, sigma, fw, fh):
    p = np.ascontiguousarray(pred_img.numpy(), dtype=np.float32)          # (4 + nc, A)
    out = np.zeros((max(max_det, 1), 6), dtype=np.float32)
    n = f(p.ctypes.data, nc, p.shape[1], conf, iou, max_det, cluster, sigma, fw, fh, out.ctypes.data, out.shape[0])
    return torch.from_numpy(out[:n].copy())


def main():
    f = load_ref()
    cases = []
    for (B, nc, A, seed, dense, conf, iou, sigma, cluster, max_det, fw, fh) in CASES:
        pred = synth_predictions(B, nc, A, seed, dense)
        dets = [reference_dets(f, pred[b], nc, conf, iou, max_det, cluster, sigma, fw, fh) for b in range(B)]
        print("cwnms case", seed, [len(d) for d in dets])
        cases.append({"B": B, "nc": nc, "A": A, "seed": seed, "dense": dense, "conf": conf, "iou": iou, "sigma": sigma,
                      "cluster": cluster, "max_det": max_det, "frame_w": fw, "frame_h": fh, "dets": dets})
    torch.save({"cases": cases}, os.path.join(HERE, "cwnms.golden.pt"))


if __name__ == "__main__":
    main()
