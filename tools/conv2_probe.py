"""Where does the persistent tcgen05 conv kernel spend its time?  Parts of it are switched off one at a time (ym_set_conv2_debug; outputs
are invalid, only the time is read) on three layer shapes of yolo26-master-n at bs32, graph-timed (20 launches per graph, 5 replays).
    python tools/conv2_probe.py [out.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from yolo_master_b200 import _lib  # noqa: E402
from yolo_master_b200.nn import modules as M  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tools"))
L = _lib.load()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


FLAGS = [("full kernel", 0), ("no TMA store", 1), ("no store, no proxy fence", 1 | 16), ("no store, no fence, no barrier", 1 | 16 | 32),
         ("no store / fence / barrier / staging writes (arithmetic only)", 1 | 16 | 32 | 64), ("no epilogue arithmetic / staging / store", 3),
         ("no tensor-memory read either", 7), ("weights once per CTA", 8), ("ring 2", 2 << 8), ("ring 4 (one CTA per SM)", 4 << 8)]
res = {}
for (c1, c2, k, st, hw, act) in [(64, 192, 1, 1, 80, False), (48, 64, 1, 1, 160, True), (192, 64, 1, 1, 80, True), (32, 64, 3, 2, 160, True)]:
    conv = M.Conv(c1, c2, k, st, act=act).cuda().eval()
    xin = torch.randn((32, c1, hw, hw), device="cuda").half().contiguous(memory_format=torch.channels_last)
    for name, fl in FLAGS:
        prev = L.ym_set_conv2_debug(fl)
        try:
            with torch.no_grad():
                us = timed(lambda: conv(xin))
        except Exception as e:
            us = f"{type(e).__name__}: {str(e)[:80]}"
        L.ym_set_conv2_debug(prev)
        res[f"{c1}_{c2}_k{k}_s{st}_{hw}:{name}"] = us
        print(f"conv {c1}->{c2} k{k} s{st} @{hw}: {name}: {us}", flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
