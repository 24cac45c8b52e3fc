"""Diagnostic (not a test): teacher-forced per-layer error of the v0_1 classification model on the GPU vs the CPU oracle at 64x64
(the golden's shape: maps shrink to 2x2) - each layer is fed the ORACLE's input, so the first layer that deviates is the culprit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from _util import synth_sd_from_keys, yaml_of  # noqa: E402
from oracle import yolo_master_oracle as O  # noqa: E402
from yolo_master_b200.nn.tasks import ClassificationModel  # noqa: E402
from yolo_master_b200.utils.synth import synth_images  # noqa: E402

NAME, CFG = "yolo-master-cls-n-v0_1", "master/v0_1/cls/yolo-master-cls-n.yaml"
sd = synth_sd_from_keys(0, NAME)
m = ClassificationModel(CFG)
m.load_state_dict(sd, strict=True)
m = m.to("cuda").eval()
for size in (64, 128):
    x = synth_images(3, size, size, 5).half()
    ref, ys = O.forward(O.parse_spec(yaml_of(CFG)), sd, x.float(), return_layers=True)
    print("== input", size)
    for i, layer in enumerate(m.model):
        if i not in ys or not torch.is_tensor(ys[i]):
            continue
        f = layer.f
        src = x.float() if (i == 0) else (ys[f if f != -1 else i - 1] if isinstance(f, int) else [ys[j if j != -1 else i - 1] for j in f])
        if isinstance(src, list) or i == len(m.model) - 1:
            continue
        with torch.no_grad():
            out = layer(src.half().cuda().contiguous(memory_format=torch.channels_last))
        out = out[0] if isinstance(out, tuple) else out
        a, b = out.float().cpu(), ys[i]
        print(f"layer {i:2d} {type(layer).__name__:24s} out {tuple(b.shape)} rms {float(b.pow(2).mean().sqrt()):.3f} "
              f"max_err {float((a - b).abs().max()):.3e} mean_err {float((a - b).abs().mean()):.3e}", flush=True)
