"""Device-resident CUDA-graph throughput of every supported model family (not the bench.py line, which is yolo26-master-n):
    python tools/bench_models.py [batch] > gpurun_out/models.json
Variants: yolo26-master-n (A2C2fMoE), yolo-master-{n,s,l} v0 (ES_MOE + A2C2f + DFL), yolo26-master-moa-mot-{n,s} (C2fMoT/C2fMoA),
plus configs[2] (s MoT+MoA bs64) and configs[3] (l @1280, 16 images = one rank's shard), the v0_1 zoo (ModularRouterExpertMoE) and the
v0_10 zoo (VisualEnhancedAdaptiveGateMoE)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import synth_sd_from_keys  # noqa: E402
from yolo_master_b200.nn.tasks import DetectionModel, yaml_model_load  # noqa: E402
from yolo_master_b200.utils.synth import synth_images  # noqa: E402

B0 = int(sys.argv[1]) if len(sys.argv) > 1 else 32


def moamot(scale):
    d = yaml_model_load("yolo26-master-moa-mot-n.yaml")
    if scale:
        d["scales"]["s"] = [0.50, 0.50, 1024]
        d["scale"] = "s"
    return d


CASES = [("yolo26-master-n", "yolo26-master-n.yaml", "yolo26-master-n", B0, 640),
         ("yolo-master-n-v0", "master/v0/det/yolo-master-n.yaml", "yolo-master-n-v0", B0, 640),
         ("yolo-master-l-v0 @1280 (configs[3] shard)", "master/v0/det/yolo-master-l.yaml", "yolo-master-l-v0", 16, 1280),
         ("yolo26-master-moa-mot-n", moamot(False), "yolo26-master-moa-mot-n", B0, 640),
         ("yolo26-master-moa-mot-s bs64 (configs[2])", moamot(True), "yolo26-master-moa-mot-s", 64, 640),
         # families added after round 1's GPU budget (first hardware numbers pending): v0_1 ModularRouterExpertMoE, v0_10 gated MoE
         ("yolo-master-n-v0_1", "master/v0_1/det/yolo-master-n.yaml", "yolo-master-n-v0_1", B0, 640),
         ("yolo-master-n-v0_10", "master/v0_10/det/yolo-master-n.yaml", "yolo-master-n-v0_10", B0, 640)]
res = {}
for tag, cfg, keys, B, S in CASES:
    try:
        m = DetectionModel(cfg)
        m.load_state_dict(synth_sd_from_keys(0, keys))
        m.to("cuda").eval()
        xs = [synth_images(B, S, S, 500 + i).half().cuda() for i in range(3)]
        g = m.graphed(B, S, S)
        for i in range(3):
            g(xs[i % 3])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        n = 10
        e0.record()
        for i in range(n):
            g(xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[tag] = {"batch": B, "imgsz": S, "ms_per_step": ms, "images_per_s": B / (ms * 1e-3), "kernels_per_step": g.kernels_per_replay}
        del m, g, xs
        torch.cuda.empty_cache()
    except Exception as e:
        res[tag] = {"error": str(e)[:300]}
print(json.dumps(res, indent=1))
