"""Large-candidate NMS path (validation at conf 0.001, up to max_nms = 30000 candidates per image) without a
GPU: the per-image algorithm of yolo-master_b200/csrc/nms_large_core.cuh is compiled for the HOST with g++ (tests/native/
nms_large_host.cpp: the executor runs every step for all 1024 thread ids, which is what the CTA does between barriers) and compared
bit-for-bit with the NMS oracle (pinned to the reference's non_max_suppression by tests/test_nms_oracle.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from _util import ROOT
from oracle import nms_oracle as N

vp, ci, cf = C.c_void_p, C.c_int, C.c_float


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("nms_large_host") / "libnms_large_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "yolo-master_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "nms_large_host.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.host_nms_batched_large.argtypes = [vp, ci, ci, ci, cf, cf, ci, ci, cf, vp, vp, vp]
    return lib


def _scene(B, nc, A, seed, frame=1280.0):
    """Seeded dense scene: clustered boxes (so suppression happens), Beta-like scores with exact ties thrown in."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.rand((B, 2, A // 20 + 1), generator=g) * frame
    cxy = centres.repeat_interleave(20, 2)[:, :, :A] + torch.randn((B, 2, A), generator=g) * 6
    wh = torch.exp(torch.randn((B, 2, A), generator=g) * 0.5 + 3.5)
    scores = torch.rand((B, nc, A), generator=g) ** 6
    m = scores[:, :, 1::97].shape[2]
    scores[:, :, ::97][:, :, :m] = scores[:, :, 1::97]                                   # exact score ties across anchors
    return torch.cat([cxy, wh, scores], 1).contiguous()


def _run(host, pred, conf, iou, max_det=300, max_nms=30000, max_wh=7680.0):
    B, no, A = pred.shape
    out = torch.empty((B, max_det, 6))
    cnt = torch.empty((B,), dtype=torch.int32)
    idx = torch.empty((B, max_det), dtype=torch.int32)
    host.host_nms_batched_large(pred.data_ptr(), B, no - 4, A, conf, iou, max_det, max_nms, max_wh, out.data_ptr(), cnt.data_ptr(), idx.data_ptr())
    return out, cnt, idx


@pytest.mark.parametrize("B,nc,A,conf,iou,max_nms", [
    (2, 3, 33600, 0.001, 0.7, 30000),      # validation setting at 1280^2: > 30000 candidates -> the max_nms cut is exercised
    (1, 80, 8400, 0.001, 0.6, 30000),      # 640^2, all 8400 anchors are candidates
    (2, 2, 20000, 0.3, 0.5, 5000),         # a few thousand candidates, small max_nms
    (1, 1, 70, 0.999999, 0.5, 30000),      # nothing above the threshold
    (2, 4, 777, 0.05, 0.45, 30000),        # A not a power of two, fewer candidates than max_det
])
def test_large_path_matches_oracle(host, B, nc, A, conf, iou, max_nms):
    pred = _scene(B, nc, A, 17 + A)
    out, cnt, idx = _run(host, pred, conf, iou, max_nms=max_nms)
    ro, rk = N.non_max_suppression(pred, conf, iou, max_det=300, max_nms=max_nms)
    ncand = [int((pred[b, 4:].amax(0) > conf).sum()) for b in range(B)]
    if A == 33600:
        assert max(ncand) > 30000                                   # the case is past the shared-memory path's 16384 AND max_nms
    for b in range(B):
        n = int(cnt[b])
        assert n == len(rk[b]), (b, n, len(rk[b]), ncand[b])
        assert torch.equal(idx[b, :n].long(), rk[b])                # kept anchors, in order: bit-exact decisions
        assert torch.equal(out[b, :n], ro[b])                       # boxes (xyxy), confidence, class: bit-exact
        assert bool((idx[b, n:] == -1).all()) and bool((out[b, n:] == 0).all())


def test_large_path_max_det_and_ties(host):
    """max_det smaller than the survivor count; identical boxes with identical scores keep the lower anchor."""
    pred = torch.zeros((1, 5, 6))
    pred[0, :4] = torch.tensor([[50.0, 50.0, 50.0, 200.0, 200.0, 400.0], [50.0] * 6, [20.0] * 6, [20.0] * 6])
    pred[0, 4] = torch.tensor([0.9, 0.9, 0.9, 0.8, 0.8, 0.7])
    out, cnt, idx = _run(host, pred, 0.1, 0.5, max_det=2)
    assert int(cnt[0]) == 2 and idx[0].tolist() == [0, 3]
    ro, rk = N.non_max_suppression(pred, 0.1, 0.5, max_det=2)
    assert rk[0].tolist() == [0, 3] and torch.equal(out[0, :2], ro[0])
