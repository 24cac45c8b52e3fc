"""Sweep ym_set_attention_poly on the P3 area-attention shape (bs32, N=6400, 2 heads x d32): ms per launch and the deviation of
the output from the all-MUFU kernel.  Usage: python tools/sweep_attention_poly.py [batch]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_master_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N, heads, hd = 6400, 2, 32
dev = torch.device("cuda")
torch.manual_seed(0)
qkv = torch.randn((B, 80, 80, 3 * heads * hd), device=dev).half()
L = ops.lib()
res = {}
ref = None
for poly in (0, "mode1", "mode2", "mode3", 0, "mode2"):
    if isinstance(poly, str):
        L.ym_set_attention_poly(0)
        L.ym_set_attention_chunked(int(poly[-1]))
    else:
        L.ym_set_attention_chunked(0)
        L.ym_set_attention_poly(poly)
    out = ops.new_act(B, 80, 80, heads * hd, dev)
    for _ in range(3):
        ops.attention(qkv, B, N, heads, 3 * hd, 0, hd, 2 * hd, hd, hd, hd ** -0.5, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ops.attention(qkv, B, N, heads, 3 * hd, 0, hd, 2 * hd, hd, hd, hd ** -0.5, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if ref is None:
        ref = out.float().clone()
    err = float((out.float() - ref).abs().max())
    res[f"poly{poly}" + ("_again" if f"poly{poly}" in res else "")] = {"ms": ms, "max_abs_dev_vs_mufu": err}
L.ym_set_attention_poly(0)
L.ym_set_attention_chunked(0)
print(json.dumps(res))
