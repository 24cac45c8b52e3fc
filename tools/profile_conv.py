"""Profiles representative conv launches (tc_conv v2) inside a profiler window: L2.cv2 (1x1 48->64 @160^2), an attention
proj with residual (1x1 64->64 @80^2), L3 (3x3 s2 64->64) and a neck 3x3 (32->32 @40^2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yolo_master_b200 import ops  # noqa: E402
from yolo_master_b200.nn import modules as M  # noqa: E402

torch.manual_seed(0)
cases = [(48, 64, 1, 1, 160, False), (64, 64, 1, 1, 80, True), (64, 64, 3, 2, 160, False), (32, 32, 3, 1, 40, False),
         (64, 192, 1, 1, 80, False)]
mods = []
for c1, c2, k, s, hw, res in cases:
    m = M.Conv(c1, c2, k, s).cuda().eval()
    x = torch.randn(32, hw, hw, c1, device="cuda").half()
    r = torch.randn(32, hw // s, hw // s, c2, device="cuda").half() if res else None
    mods.append((m, x, r))
for impl in ("tc", "legacy"):
    ops.CONV_IMPL = impl
    for m, x, r in mods:
        for _ in range(2):
            m.fwd_nhwc(x, res=r)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for impl in ("tc", "legacy"):
    ops.CONV_IMPL = impl
    for (m, x, r), c in zip(mods, cases):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.fwd_nhwc(x, res=r)
        e1.record()
        torch.cuda.synchronize()
        print(impl, c, "ms", round(e0.elapsed_time(e1), 4))
torch.cuda.profiler.stop()
