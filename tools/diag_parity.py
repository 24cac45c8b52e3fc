"""Diagnostic (not a test): per-layer error of the CUDA path vs the CPU oracle at 640x640, and eager timing."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from _util import close_stats, synth_sd_from_keys, yaml_n  # noqa: E402
from oracle import yolo_master_oracle as O  # noqa: E402
from yolo_master_b200.nn.tasks import DetectionModel  # noqa: E402
from yolo_master_b200.utils.synth import synth_images  # noqa: E402

sd = synth_sd_from_keys(0)
m = DetectionModel("yolo26-master-n.yaml")
m.load_state_dict(sd)
m.to("cuda").eval()
x = synth_images(2, 640, 640, 3)
feats = {}
hooks = [mod.register_forward_hook(lambda mod, i, o, k=k: feats.__setitem__(k, o)) for k, mod in enumerate(m.model)]
with torch.no_grad():
    y = m(x.half().cuda())[0]
torch.cuda.synchronize()
for h in hooks:
    h.remove()
ref, ys = O.forward(O.parse_spec(yaml_n()), sd, x.half().float(), return_layers=True)
for i in range(23):
    if feats.get(i) is None:
        continue
    a, b = feats[i].float().cpu(), ys[i]
    mx, bad = close_stats(a, b)
    print(f"layer {i:2d} {m.model[i].type:10s} shape {tuple(b.shape)} rms {float(b.pow(2).mean().sqrt()):.3f} "
          f"max_err {mx:.3e} mean_err {float((a - b).abs().mean()):.3e} frac_outside_tol {bad:.2e}")
yy = y.float().cpu()
print("final score diff (sorted)", float((yy[..., 4].sort(dim=1, descending=True)[0] - ref[..., 4].sort(dim=1, descending=True)[0]).abs().max()))
print("top5 ours", yy[0, :5].tolist())
print("top5 ref ", ref[0, :5].tolist())
xb = synth_images(32, 640, 640, 4).half().cuda()
with torch.no_grad():
    for _ in range(2):
        m(xb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        m(xb)
    torch.cuda.synchronize()
    print("eager bs32 ms/step", (time.perf_counter() - t0) / 5 * 1e3)
