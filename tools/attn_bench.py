"""Area-attention kernels behind ym_attention_fwd timed alone (CUDA events, 20 launches each after 3 warm-ups) on the shapes of
yolo26-master-n at 640x640 bs32: P3 N=6400, P4 N=1600, P5 N=400 (2 heads x d32) and C2PSA's N=400, d_v=64.
    python tools/attn_bench.py [out.json]            impl 2 = tc_attention2 (default), 1 = tc_attention, 0 = mma.sync
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yolo_master_b200 import _lib, ops  # noqa: E402

L = _lib.load()
res = {}
for name, (B, N, heads, hd, dv) in {"P3": (32, 6400, 2, 32, 32), "P4": (32, 1600, 2, 32, 32), "P5": (32, 400, 2, 32, 32),
                                    "PSA": (32, 400, 2, 32, 64)}.items():
    hs = 2 * hd + dv
    qkv = torch.randn((B, N, 1, heads * hs), device="cuda").half()
    out = torch.empty((B, N, 1, heads * dv), device="cuda", dtype=torch.float16)
    ref = None
    for impl in (2, 1, 0):
        prev = L.ym_set_attention_impl(impl)
        for _ in range(3):
            ops.attention(qkv, B, N, heads, hs, 0, hd, 2 * hd, hd, dv, hd ** -0.5, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            ops.attention(qkv, B, N, heads, hs, 0, hd, 2 * hd, hd, dv, hd ** -0.5, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        L.ym_set_attention_impl(prev)
        if ref is None:
            ref = out.clone()
        dev = float((out.float() - ref.float()).abs().max())
        scores = float(N) * N * heads * B
        res[f"{name}_impl{impl}"] = {"ms": ms, "Tscores_per_s": scores / (ms * 1e-3) / 1e12, "mufu_frac": scores / (ms * 1e-3) / (148 * 16 * 1.965e9),
                                    "tflops": 4 * scores * hd / (ms * 1e-3) / 1e12 if dv == hd else None, "max_abs_dev_vs_impl2": dev}
        print(name, "impl", impl, res[f"{name}_impl{impl}"], flush=True)
# query tiles per CTA forced to 1 / 2 (ym_set_attention2_qtiles; 0 = the launcher's wave-fit rule) at the small pyramid levels
for name, (B, N, heads, hd, dv) in {"P3": (32, 6400, 2, 32, 32), "P4": (32, 1600, 2, 32, 32), "P5": (32, 400, 2, 32, 32), "PSA": (32, 400, 2, 32, 64)}.items():
    hs = 2 * hd + dv
    qkv = torch.randn((B, N, 1, heads * hs), device="cuda").half()
    out = torch.empty((B, N, 1, heads * dv), device="cuda", dtype=torch.float16)
    for qt in (0, 1, 2):
        old = L.ym_set_attention2_qtiles(qt)
        reps = 20 if name == "P3" else 50
        for _ in range(3):
            ops.attention(qkv, B, N, heads, hs, 0, hd, 2 * hd, hd, dv, hd ** -0.5, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ops.attention(qkv, B, N, heads, hs, 0, hd, 2 * hd, hd, dv, hd ** -0.5, out=out)
        e1.record()
        torch.cuda.synchronize()
        L.ym_set_attention2_qtiles(old)
        ms = e0.elapsed_time(e1) / reps
        res[f"{name}_qtiles{qt}"] = {"ms": ms, "mufu_ceiling_frac": float(N) * N * heads * B / (ms * 1e-3) / (148 * 16 * 1.965e9)}
        print(name, "q_tiles", qt, res[f"{name}_qtiles{qt}"], flush=True)
L.ym_set_attention2_poly(0)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
