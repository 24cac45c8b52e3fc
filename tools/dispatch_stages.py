"""Stage-ablation timing of tc_dispatch_kernel (BASELINE configs[4]): ym_set_dispatch_debug masks switch off weight loads (1),
x loads (2), output stores (4), epilogue math (8) and MMAs (16); each variant is timed in a CUDA graph over 6 rotating buffers."""
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
ops.DISPATCH_IMPL = sys.argv[1] if len(sys.argv) > 1 else "v3"
print("impl", ops.DISPATCH_IMPL)
lib = ops.lib()
names = {0: "full", 1: "-W loads", 2: "-x loads", 4: "-stores", 8: "-epilogue math", 16: "-MMA", 3: "-W -x loads", 12: "-epilogue -stores",
         28: "loads only", 15: "MMA only", 19: "epilogue+stores only", 31: "empty pipeline"}
for mask, nm in names.items():
    lib.ym_set_dispatch_debug(mask)
    for i in range(nrot):
        ops.moe_dispatch(xs[i], Wt, idx, w, out=outs[i])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for i in range(nrot):
            ops.moe_dispatch(xs[i], Wt, idx, w, out=outs[i])
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"mask {mask:2d} {nm:24s} {e0.elapsed_time(e1) / (5 * nrot) * 1e3:7.2f} us")
lib.ym_set_dispatch_debug(0)
