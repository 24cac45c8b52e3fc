"""Single kernels of the forward timed alone at the bs32 / 640x640 shapes (CUDA events, 20 launches after 3 warm-ups).
    python tools/op_bench.py [out.json]
stem: both implementations (ym_set_stem_impl 0 = FFMA, 1 = mma.sync); concat2 at the P3 neck shape.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yolo_master_b200 import _lib, ops  # noqa: E402
from yolo_master_b200.nn import modules as M  # noqa: E402

L = _lib.load()
res = {}


def timed(fn, reps=20):
    """us per call: `reps` calls captured in one CUDA graph (the Python launch path costs ~45 us per call, more than most of these
    kernels), the graph replayed 5 times after a warm-up replay."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


x = torch.rand((32, 3, 640, 640), device="cuda").half()
stem = M.Conv(3, 16, 3, 2).cuda().eval()
for impl in (0, 1):
    prev = L.ym_set_stem_impl(impl)
    with torch.no_grad():
        us = timed(lambda: stem(x))
    L.ym_set_stem_impl(prev)
    mb = (x.numel() * 2 + 32 * 320 * 320 * 16 * 2) / 1e6
    res[f"stem_impl{impl}"] = {"us": us, "GBps": mb / us * 1e3}
    print("stem impl", impl, res[f"stem_impl{impl}"], flush=True)

a = torch.randn((32, 40, 40, 128), device="cuda").half()
b = torch.randn((32, 80, 80, 64), device="cuda").half()
out = torch.empty((32, 80, 80, 192), device="cuda", dtype=torch.float16)
us = timed(lambda: ops.concat2(a, b, up=2, out=out)) if hasattr(ops, "concat2") else None
if us:
    mb = (a.numel() + b.numel() + out.numel()) * 2 / 1e6
    res["concat2_p3"] = {"us": us, "GBps": mb / us * 1e3}
    print("concat2 P3", res["concat2_p3"], flush=True)
# depthwise 7x7 (the pe convs of the area-attention blocks): Toeplitz-GEMM kernel (1) against the FFMA kernel (0)
for (c, hw) in [(64, 80), (128, 40), (256, 20)]:
    conv = M.Conv(c, c, 7, 1, None, g=c, act=False).cuda().eval()
    xin = torch.randn((32, c, hw, hw), device="cuda").half().contiguous(memory_format=torch.channels_last)
    for impl in (1, 0):
        prev = L.ym_set_dwconv_tc(impl)
        with torch.no_grad():
            us = timed(lambda: conv(xin))
        L.ym_set_dwconv_tc(prev)
        mb = 2 * xin.numel() * 2 / 1e6
        res[f"dwconv7_{c}_{hw}_impl{impl}"] = {"us": us, "GBps": mb / us * 1e3}
        print("dwconv7", c, hw, "impl", impl, res[f"dwconv7_{c}_{hw}_impl{impl}"], flush=True)
# 3x3 layers over 8 / 16 input channels: patch-staged kernel (1) against the implicit GEMM (0)
for (c1, c2, st, hw) in [(16, 32, 2, 320), (16, 8, 1, 160), (8, 16, 1, 160), (16, 16, 1, 80), (16, 16, 1, 160)]:
    conv = M.Conv(c1, c2, 3, st).cuda().eval()
    xin = torch.randn((32, c1, hw, hw), device="cuda").half().contiguous(memory_format=torch.channels_last)
    for impl in (1, 0):
        prev = L.ym_set_small_conv_impl(impl)
        with torch.no_grad():
            us = timed(lambda: conv(xin))
        L.ym_set_small_conv_impl(prev)
        mb = (xin.numel() + 32 * c2 * (hw // st) ** 2) * 2 / 1e6
        res[f"conv3_{c1}_{c2}_s{st}_{hw}_impl{impl}"] = {"us": us, "GBps": mb / us * 1e3}
        print("conv3", c1, c2, st, hw, "impl", impl, res[f"conv3_{c1}_{c2}_s{st}_{hw}_impl{impl}"], flush=True)
# persistent tcgen05 conv kernel: two epilogue groups on alternate tiles (2) against eight warps per tile (1), layer shapes of yolo26-master-n at bs32
for (c1, c2, k, st, hw, act) in [(64, 192, 1, 1, 80, False), (48, 64, 1, 1, 160, True), (192, 64, 1, 1, 80, True), (128, 128, 1, 1, 40, True),
                                 (32, 64, 3, 2, 160, True), (64, 64, 3, 1, 80, True), (256, 256, 1, 1, 20, True)]:
    conv = M.Conv(c1, c2, k, st, act=act).cuda().eval()
    xin = torch.randn((32, c1, hw, hw), device="cuda").half().contiguous(memory_format=torch.channels_last)
    outs = {}
    for mode in (2, 1):
        prev = L.ym_set_conv2_epi_groups(mode)
        with torch.no_grad():
            us = timed(lambda: conv(xin))
            outs[mode] = conv(xin).clone()
        L.ym_set_conv2_epi_groups(prev)
        mb = (xin.numel() + 32 * c2 * (hw // st) ** 2) * 2 / 1e6
        res[f"conv_{c1}_{c2}_k{k}_s{st}_{hw}_groups{mode}"] = {"us": us, "GBps": mb / us * 1e3}
        print("conv", c1, c2, k, st, hw, "epilogue groups", mode, res[f"conv_{c1}_{c2}_k{k}_s{st}_{hw}_groups{mode}"], flush=True)
    assert torch.equal(outs[1], outs[2]), "epilogue organisations must agree bit for bit"
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
