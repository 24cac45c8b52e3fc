"""One eager bs32 640x640 forward inside a cudaProfilerStart/Stop window (for `ncu --profile-from-start off`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from _util import synth_sd_from_keys  # noqa: E402
from yolo_master_b200.nn.tasks import DetectionModel  # noqa: E402
from yolo_master_b200.utils.synth import synth_images  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = DetectionModel("yolo26-master-n.yaml")
m.load_state_dict(synth_sd_from_keys(0))
m.to("cuda").eval()
x = synth_images(B, 640, 640, 4).half().cuda()
with torch.no_grad():
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
