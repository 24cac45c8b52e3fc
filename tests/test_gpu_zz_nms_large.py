"""GPU parity of the large-candidate NMS path (`ym_nms_batched_large`, SURVEY.md 8(f) rank 3: validation at conf 0.001) against the
NMS oracle, against the shared-memory kernel where both apply, and through `non_max_suppression`'s automatic switch.

The algorithm is also verified on the host (tests/test_nms_large_host.py runs the same per-image function under g++, bit-exact
against the oracle); the CUDA launch runs on the B200 since round 2 (profiles/r02_gpu_suite.txt)."""
import pytest
import torch

from oracle import nms_oracle as N
from yolo_master_b200 import ops
from yolo_master_b200.utils.nms import non_max_suppression

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scene(B, nc, A, seed, frame=1280.0):
    g = torch.Generator().manual_seed(seed)
    centres = torch.rand((B, 2, A // 20 + 1), generator=g) * frame
    cxy = centres.repeat_interleave(20, 2)[:, :, :A] + torch.randn((B, 2, A), generator=g) * 6
    wh = torch.exp(torch.randn((B, 2, A), generator=g) * 0.5 + 3.5)
    scores = torch.rand((B, nc, A), generator=g) ** 6
    m = scores[:, :, 1::97].shape[2]
    scores[:, :, ::97][:, :, :m] = scores[:, :, 1::97]
    return torch.cat([cxy, wh, scores], 1).contiguous()


@pytest.mark.parametrize("B,nc,A,conf,iou,max_nms", [(2, 3, 33600, 0.001, 0.7, 30000), (3, 80, 8400, 0.001, 0.6, 30000),
                                                     (2, 2, 20000, 0.3, 0.5, 5000), (1, 1, 70, 0.999999, 0.5, 30000), (2, 4, 777, 0.05, 0.45, 30000)])
def test_large_kernel_matches_oracle(B, nc, A, conf, iou, max_nms):
    pred = _scene(B, nc, A, 17 + A)
    out, cnt, idx = ops.nms_batched_large(pred.to(DEV), conf, iou, 300, max_nms)
    ro, rk = N.non_max_suppression(pred, conf, iou, max_det=300, max_nms=max_nms)
    for b in range(B):
        n = int(cnt[b])
        assert n == len(rk[b])
        assert torch.equal(idx[b, :n].long().cpu(), rk[b]) and torch.equal(out[b, :n].cpu(), ro[b])


def test_large_and_shared_memory_kernels_agree():
    pred = _scene(4, 80, 8400, 5).to(DEV)
    a, ca, ia, _ = ops.nms_batched(pred, 0.25, 0.7, 300)
    b, cb, ib = ops.nms_batched_large(pred, 0.25, 0.7, 300)
    assert torch.equal(ca, cb) and torch.equal(ia, ib) and torch.equal(a, b)


def test_non_max_suppression_switches_paths():
    """conf 0.001 on a 33600-anchor prediction overflows the 16384-candidate kernel: the public function falls through to the
    global-memory path and returns the oracle's rows."""
    pred = _scene(2, 3, 33600, 23)
    out, keep = non_max_suppression(pred.to(DEV), 0.001, 0.7, max_det=300, return_idxs=True)
    ro, rk = N.non_max_suppression(pred, 0.001, 0.7, max_det=300)
    for o, k, a, b in zip(out, keep, ro, rk):
        assert torch.equal(k.cpu(), b) and torch.equal(o.cpu(), a)
