"""Runs the P3 attention (N=6400, 2 heads x 32, bs32) once per implementation inside a profiler window."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yolo_master_b200 import _lib, ops  # noqa: E402

B, N, heads, hd = 32, 6400, 2, 32
qkv = torch.randn((B, 80, 80, 3 * heads * hd), device="cuda").half()
out = ops.new_act(B, 80, 80, heads * hd, "cuda")
L = _lib.load()
for impl in (1, 2):
    L.ym_set_attention_impl(impl)
    for _ in range(2):
        ops.attention(qkv, B, N, heads, 3 * hd, 0, hd, 2 * hd, hd, hd, hd ** -0.5, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for impl in (1, 2):
    L.ym_set_attention_impl(impl)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.attention(qkv, B, N, heads, 3 * hd, 0, hd, 2 * hd, hd, hd, hd ** -0.5, out=out)
    e1.record()
    torch.cuda.synchronize()
    print("impl", impl, "ms", e0.elapsed_time(e1))
torch.cuda.profiler.stop()
