"""One launch of each depthwise 7x7 kernel at the P3 shape (bs32, 80x80, 64 channels) for ncu:
    ncu --set full --import-source on -k regex:dwconv -o out python tools/profile_dwconv.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yolo_master_b200 import _lib  # noqa: E402
from yolo_master_b200.nn import modules as M  # noqa: E402

L = _lib.load()
conv = M.Conv(64, 64, 7, 1, None, g=64, act=False).cuda().eval()
x = torch.randn((32, 64, 80, 80), device="cuda").half().contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    for impl in (1, 0):
        L.ym_set_dwconv_tc(impl)
        for _ in range(2):
            conv(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        conv(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
L.ym_set_dwconv_tc(1)
