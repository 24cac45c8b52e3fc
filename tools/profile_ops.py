"""Per-op achieved bandwidth of one eager bs32 640x640 forward: every yolo_master_b200.ops call is timed with CUDA events
(synchronised, so launch gaps are excluded) and its tensor arguments + results are counted as algorithmic bytes.
Output: one line per op call, sorted summary by op kind.   python tools/profile_ops.py [batch]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from _util import synth_sd_from_keys  # noqa: E402
from yolo_master_b200 import ops  # noqa: E402
from yolo_master_b200.nn.tasks import DetectionModel  # noqa: E402
from yolo_master_b200.utils.synth import synth_images  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
LOG = []


def tensors(obj, acc):
    if isinstance(obj, torch.Tensor):
        acc[obj.data_ptr()] = max(acc.get(obj.data_ptr(), 0), obj.numel() * obj.element_size())
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            tensors(o, acc)
    elif isinstance(obj, dict):
        for o in obj.values():
            tensors(o, acc)


def wrap(name, fn):
    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        r = fn(*a, **k)
        e1.record()
        torch.cuda.synchronize()
        acc = {}
        tensors(a, acc); tensors(k, acc); tensors(r, acc)
        shp = [tuple(t.shape) for t in list(a) + list(k.values()) if isinstance(t, torch.Tensor)][:2]
        LOG.append((name, sum(acc.values()), e0.elapsed_time(e1) * 1e3, shp, [x for x in a if isinstance(x, (int, bool))][:6]))
        return r
    return inner


NAMES = ["conv2d", "stem_conv", "dwconv", "sppf_pool", "concat2", "attention", "router_topk", "moe_expert_gemm", "gn_finalize",
         "moe_combine", "detect_topk", "detect_dense"]
m = DetectionModel("yolo26-master-n.yaml")
m.load_state_dict(synth_sd_from_keys(0))
m.to("cuda").eval()
x = synth_images(B, 640, 640, 4).half().cuda()
with torch.no_grad():
    m(x); m(x)
    for n in NAMES:
        setattr(ops, n, wrap(n, getattr(ops, n)))
    m(x)
tot = sum(t for _, _, t, _, _ in LOG)
print(f"# {len(LOG)} op calls, {tot:.0f} us (synchronised eager), batch {B}")
for i, (n, by, t, shp, ints) in enumerate(LOG):
    print(f"{i:4d} {n:16s} {t:8.1f} us {by / 1e6:8.1f} MB {by / t / 1e3:8.1f} GB/s  {shp} {ints}")
agg = collections.defaultdict(lambda: [0, 0.0, 0])
for n, by, t, _, _ in LOG:
    agg[n][0] += by; agg[n][1] += t; agg[n][2] += 1
print("# by op")
for n, (by, t, c) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:16s} {c:4d}x {t:8.1f} us {100 * t / tot:5.1f}%  {by / 1e6:9.1f} MB  {by / t / 1e3:8.1f} GB/s")
