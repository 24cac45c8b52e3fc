"""clock64 timeline of CTA 0 (leader) and CTA 1 (peer) of tc_dispatch2_kernel on BASELINE configs[4] (ym_set_dispatch_trace)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yolo_master_b200 import ops  # noqa: E402

B, C, H, W, E, K = 64, 256, 32, 32, 8, 2
g = torch.Generator().manual_seed(0)
xs = [torch.randn((B, H, W, C), generator=g).half().cuda() for _ in range(6)]
outs = [ops.new_act(B, H, W, C, "cuda") for _ in range(6)]
Wt = (torch.randn((E, C, C), generator=g) / C ** 0.5).half().cuda()
idx = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(B)]).int().cuda()
w = torch.rand((B, K), generator=g)
w = (w / w.sum(1, keepdim=True)).cuda()
lib = ops.lib()
lib.ym_set_dispatch_debug(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
tr = torch.zeros((2, 4, 256), dtype=torch.int64, device="cuda")
for i in range(6):
    ops.moe_dispatch(xs[i], Wt, idx, w, out=outs[i])
torch.cuda.synchronize()
lib.ym_set_dispatch_trace(tr.data_ptr())
ops.moe_dispatch(xs[0], Wt, idx, w, out=outs[0])
torch.cuda.synchronize()
lib.ym_set_dispatch_trace(None)
lib.ym_set_dispatch_debug(0)
t = tr.cpu()
for cta in range(2):
    t0 = int(t[cta, 3, 0])
    us = lambda v: (int(v) - t0) / 1965.0
    print(f"== CTA {cta}: prologue done {us(t[cta,3,1]):.2f} us, loops done {us(t[cta,3,2]):.2f}, exit {us(t[cta,3,3]):.2f}")
    for role, name in ((0, "producer (x load / W stage issue)"), (1, "mma (per unit: a_full, then per expert: t_empty + 4 b_full)"), (2, "epilogue (per expert: t_full, slot released; per unit: rows stored)")):
        v = [us(x) for x in t[cta, role] if int(x) != 0]
        print(f"  {name}: {len(v)} events")
        for k in range(0, len(v), 12):
            print("    " + " ".join(f"{x:6.2f}" for x in v[k:k + 12]))
