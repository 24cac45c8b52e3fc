"""One eager forward of a model variant inside a cudaProfilerStart/Stop window (for `ncu --profile-from-start off`).
    python tools/profile_model.py {n|v0n|v0l|moamot|moamots} [batch] [imgsz]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from _util import synth_sd_from_keys  # noqa: E402
from yolo_master_b200.nn.tasks import DetectionModel, yaml_model_load  # noqa: E402
from yolo_master_b200.utils.synth import synth_images  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "n"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 640
if which in ("moamot", "moamots"):
    cfg = yaml_model_load("yolo26-master-moa-mot-n.yaml")
    if which == "moamots":
        cfg["scales"]["s"] = [0.50, 0.50, 1024]
        cfg["scale"] = "s"
    keys = "yolo26-master-moa-mot-s" if which == "moamots" else "yolo26-master-moa-mot-n"
else:
    cfg, keys = {"n": ("yolo26-master-n.yaml", "yolo26-master-n"), "v0n": ("master/v0/det/yolo-master-n.yaml", "yolo-master-n-v0"),
                 "v0l": ("master/v0/det/yolo-master-l.yaml", "yolo-master-l-v0")}[which]
m = DetectionModel(cfg)
m.load_state_dict(synth_sd_from_keys(0, keys))
m.to("cuda").eval()
x = synth_images(B, S, S, 4).half().cuda()
with torch.no_grad():
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
