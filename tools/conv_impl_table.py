"""Times every distinct conv2d call of a model forward under both implementations (tc: TMA + tcgen05 persistent kernel;
legacy: mma.sync implicit GEMM) at the bench shape, CUDA events over 20 warm launches each.
    python tools/conv_impl_table.py [batch] > gpurun_out/conv_table.json"""
import json
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
m = DetectionModel("yolo26-master-n.yaml")
m.load_state_dict(synth_sd_from_keys(0))
m.to("cuda").eval()
calls = {}
orig = ops.conv2d


def spy(x, w_packed, bias, Cout, KH, KW, stride, pad, act, out=None, res=None, out_f32=False):
    key = (tuple(x.shape), ops.pitch(x), Cout, KH, stride, pad, bool(act), res is not None, bool(out_f32),
           None if out is None else ops.pitch(out, torch.float32 if out_f32 else torch.float16))
    if key not in calls:
        calls[key] = [0, (x, w_packed, bias, Cout, KH, KW, stride, pad, act, out, res, out_f32)]
    calls[key][0] += 1
    return orig(x, w_packed, bias, Cout, KH, KW, stride, pad, act, out=out, res=res, out_f32=out_f32)


ops.conv2d = spy          # the modules call `ops.conv2d(...)` through the module attribute, so patching it is enough
with torch.no_grad():
    m(synth_images(B, 640, 640, 1).half().cuda())
ops.conv2d = orig
rows = []
for key, (n, a) in calls.items():
    x, w, b, Cout, KH, KW, s, p, act, out, res, f32 = a
    t = {}
    for impl in ("tc", "legacy"):
        ops.CONV_IMPL = impl
        for _ in range(3):
            orig(x, w, b, Cout, KH, KW, s, p, act, out=out, res=res, out_f32=f32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            orig(x, w, b, Cout, KH, KW, s, p, act, out=out, res=res, out_f32=f32)
        e1.record()
        torch.cuda.synchronize()
        t[impl] = e0.elapsed_time(e1) / 20 * 1e3
    ops.CONV_IMPL = "tc"
    Bx, H, W, Cin = x.shape
    rows.append({"calls": n, "in": [Bx, H, W, Cin], "ldx": key[1], "Cout": Cout, "k": KH, "s": s, "res": key[7], "f32": f32,
                 "tc_us": round(t["tc"], 1), "legacy_us": round(t["legacy"], 1)})
rows.sort(key=lambda r: -r["calls"] * min(r["tc_us"], r["legacy_us"]))
tot_tc = sum(r["calls"] * r["tc_us"] for r in rows)
tot_best = sum(r["calls"] * min(r["tc_us"], r["legacy_us"]) for r in rows)
print(json.dumps({"total_tc_us": tot_tc, "total_best_us": tot_best, "rows": rows}, indent=0))
