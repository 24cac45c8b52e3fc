"""ES-MoE dispatch (BASELINE configs[4]: 65536 tokens x d=256, 8 experts top-2) inside a profiler window, plus event timings
of single launches (eager, includes host encode time) for v1 and v2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yolo_master_b200 import ops  # noqa: E402

B, C, H, W, E, K = 64, 256, 32, 32, 8, 2
g = torch.Generator().manual_seed(0)
nrot = 6
xs = [torch.randn((B, H, W, C), generator=g).half().cuda() for _ in range(nrot)]
outs = [ops.new_act(B, H, W, C, "cuda") for _ in range(nrot)]
Wt = (torch.randn((E, C, C), generator=g) / C ** 0.5).half().cuda()
idx = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(B)]).int().cuda()
w = torch.rand((B, K), generator=g)
w = (w / w.sum(1, keepdim=True)).cuda()
for impl in ("v2", "v1"):
    ops.DISPATCH_IMPL = impl
    for i in range(nrot):
        ops.moe_dispatch(xs[i], Wt, idx, w, out=outs[i])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for impl in ("v2", "v1"):
    ops.DISPATCH_IMPL = impl
    for i in range(2):
        ops.moe_dispatch(xs[i], Wt, idx, w, out=outs[i])
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
