"""YAML -> model graph and the layer loop, mirroring `ultralytics/nn/tasks.py`.

`parse_model` follows tasks.py:2022-2275 (+ mixture_registry.py:84-156 for the registered MoE modules) for the modules on
the hot path; `DetectionModel` follows tasks.py:530-577 / `_predict_once` :182-218.  Differences, all deliberate:
  * strides are derived analytically from the layer table (the reference runs a CPU forward, tasks.py:555-559);
  * `nn.Upsample` + `Concat` pairs are executed as one kernel;
  * the whole forward is host-sync free, so `graphed()` captures it into a CUDA graph per input shape.
"""
from __future__ import annotations

import ast
import contextlib
import math
import os
from copy import deepcopy

import torch
import torch.nn as nn

from . import modules as M

_CFG_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cfg", "models")

MODULES = {name: getattr(M, name) for name in M.__all__ if isinstance(getattr(M, name), type)}
MIXTURE_MODULES = {"A2C2fMoE": M.A2C2fMoE, "ES_MOE": M.ES_MOE, "C2fMoT": M.C2fMoT, "C2fMoA": M.C2fMoA}
MIXTURE_MODULES.update({"ModularRouterExpertMoE": M.OptimizedMOEImproved, "OptimizedMOEImproved": M.OptimizedMOEImproved})   # modules.py:1745
MIXTURE_MODULES["UltraOptimizedMoE"] = M.UltraOptimizedMoE                                                                       # v0_1 uomoe / v0_2 zoos
MIXTURE_MODULES.update({n: getattr(M, n) for n in (          # the AdaptiveGateMoE line, v0_4 ... v0_10 zoos (nn/modules/gated.py)
    "AdaptiveGateMoE", "FusedAdaptiveGateMoE", "HybridAdaptiveGateMoE", "LowRankHybridAdaptiveGateMoE",
    "RefinedLowRankHybridAdaptiveGateMoE", "DetailAwareLowRankHybridAdaptiveGateMoE", "ContextRefinedLowRankHybridAdaptiveGateMoE",
    "VisualEnhancedAdaptiveGateMoE", "UltimateOptimizedMoE", "HybridAdaptiveGateMoEv2", "OptimalHybridGateMoE", "MultiHeadRouterMoE",
    "GatedFusionMoE", "SharedExpertMoE", "DiversifiedExpertMoE")})                                      # + the v0_3 zoo block
BASE_MODULES = frozenset({M.Conv, M.DWConv, M.Bottleneck, M.SPPF, M.C2PSA, M.C2f, M.C3k2, M.C3, M.A2C2f, M.Classify})
REPEAT_MODULES = frozenset({M.C2f, M.C3k2, M.C3, M.C2PSA, M.A2C2f})
MIXTURE_BASE_MODULES = frozenset(MIXTURE_MODULES.values())
MIXTURE_REPEAT_MODULES = frozenset({M.A2C2fMoE, M.C2fMoA, M.C2fMoT})


def make_divisible(x, divisor):
    return int(math.ceil(x / divisor) * divisor)


def yaml_model_load(path):
    """Load a model YAML; bare names resolve against this package's cfg/models tree (e.g. 'yolo26-master-n.yaml')."""
    import yaml

    if not os.path.exists(path):
        if os.path.exists(os.path.join(_CFG_ROOT, path)):          # 'master/v0_10/det/yolo-master-n.yaml'
            path = os.path.join(_CFG_ROOT, path)
        else:                                                     # bare name: must be unique in the tree (the reference's check_yaml raises too)
            hits = []
            for root, dirs, files in os.walk(_CFG_ROOT):
                dirs.sort()
                if os.path.basename(path) in files:
                    hits.append(os.path.join(root, os.path.basename(path)))
            if not hits:
                raise FileNotFoundError(path)
            if len(hits) > 1:
                rel = [os.path.relpath(h, _CFG_ROOT) for h in hits]
                raise FileNotFoundError(f"Multiple files match '{path}', specify the path below cfg/models: {rel}")
            path = hits[0]
    with open(path) as f:
        d = yaml.safe_load(f)
    d["yaml_file"] = path
    return d


def _resolve(name):
    if name == "nn.Upsample":
        return M.Upsample
    if name.startswith("nn."):
        return getattr(nn, name[3:])
    if name in MODULES:
        return MODULES[name]
    if name in MIXTURE_MODULES:
        return MIXTURE_MODULES[name]
    raise KeyError(f"unknown model module {name!r} (not on the B200 hot path)")


def parse_model(d, ch, verbose=False):
    """Parse a YOLO model.yaml dictionary into (nn.Sequential, save list, per-layer output strides)."""
    legacy = True
    max_channels = float("inf")
    nc, scales, end2end = (d.get(x) for x in ("nc", "scales", "end2end"))
    reg_max = d.get("reg_max", 16)
    depth, width = d.get("depth_multiple", 1.0), d.get("width_multiple", 1.0)
    scale = d.get("scale")
    if scales:
        if not scale:
            scale = next(iter(scales.keys()))
        depth, width, max_channels = scales[scale]
    if d.get("activation"):
        raise NotImplementedError("custom default activations are not on the B200 path (SiLU only)")
    ch = [ch]
    strides = []
    layers, save, c2 = [], [], ch[-1]
    for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
        m = _resolve(m)
        args = list(args)
        for j, a in enumerate(args):
            if isinstance(a, str):
                with contextlib.suppress(ValueError, SyntaxError):
                    named = {"nc": nc, "reg_max": reg_max, "end2end": end2end, "kpt_shape": d.get("kpt_shape")}
                    args[j] = named[a] if a in named else ast.literal_eval(a)
        n = n_ = max(round(n * depth), 1) if n > 1 else n
        s_in = (1 if i == 0 else strides[f]) if isinstance(f, int) else strides[f[0]]
        s_out = s_in
        if m in BASE_MODULES:
            c1, c2 = ch[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [c1, c2, *args[1:]]
            if m in REPEAT_MODULES:
                args.insert(2, n)
                n = 1
            if m is M.C3k2:
                legacy = False
                if scale and scale in "mlx":
                    args[3] = True
            if m is M.A2C2f:
                legacy = False
                if scale and scale in "lx":
                    args.extend((True, 1.2))
            if m in (M.Conv, M.DWConv) and len(args) > 3:
                s_out = s_in * args[3]
        elif m is M.LatentMixture:                 # multi-input mixture module (mixture_registry.py:96-141): [in_channels list, c2, *rest]
            c2 = args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [[ch[x] for x in f], c2, *args[1:]]
        elif m in MIXTURE_BASE_MODULES:
            c1, c2 = ch[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [c1, c2, *args[1:]]
            if m in MIXTURE_REPEAT_MODULES:
                args.insert(2, n)
                n = 1
            if m is M.A2C2fMoE:
                legacy = False
        elif m is M.Concat:
            c2 = sum(ch[x] for x in f)
        elif m in (M.Detect, M.Pose, M.Segment, M.OBB):
            args.extend([reg_max, end2end, [ch[x] for x in f]])
            if m is M.Segment:
                args[2] = make_divisible(min(args[2], max_channels) * width, 8)      # npr scales with the width (tasks.py:2224-2225)
            m.legacy = legacy
        elif m is M.Upsample:
            c2 = ch[f]
            s_out = s_in / float(args[1])
        else:
            c2 = ch[f]
        m_ = nn.Sequential(*(m(*args) for _ in range(n))) if n > 1 else m(*args)
        m_.np = sum(x.numel() for x in m_.parameters())
        m_.i, m_.f, m_.type = i, f, m.__name__
        if verbose:
            print(f"{i:>3}{f!s:>20}{n_:>3}{m_.np:10.0f}  {m_.type:<20}{args!s:<30}")
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(m_)
        if i == 0:
            ch = []
        ch.append(c2)
        strides.append(s_out)
    return nn.Sequential(*layers), sorted(save), strides


class DetectionModel(nn.Module):
    """`DetectionModel(cfg='yolo26-master-n.yaml', ch=3, nc=None, verbose=False)` — inference-only B200 model."""

    def __init__(self, cfg="yolo26-master-n.yaml", ch=3, nc=None, verbose=False):
        super().__init__()
        self.yaml = cfg if isinstance(cfg, dict) else yaml_model_load(cfg)
        self.yaml["channels"] = ch
        if nc and nc != self.yaml["nc"]:
            self.yaml["nc"] = nc
        self.model, self.save, layer_strides = parse_model(deepcopy(self.yaml), ch=ch, verbose=verbose)
        self.names = {i: f"{i}" for i in range(self.yaml["nc"])}
        self.inplace = True
        head = self.model[-1]
        if isinstance(head, M.Detect):
            head.stride = torch.tensor([float(layer_strides[j]) for j in head.f])
            self.stride = head.stride
        else:
            self.stride = torch.tensor([32.0])
        # Upsample immediately followed by Concat([-1, j]) and consumed by nothing else: executed as one kernel
        self._fused_up = {}
        for i, m in enumerate(self.model[:-1]):
            nxt = self.model[i + 1]
            if isinstance(m, M.Upsample) and isinstance(nxt, M.Concat) and isinstance(nxt.f, list) and nxt.f[0] == -1 \
                    and i not in self.save and m.mode == "nearest" and float(m.scale_factor).is_integer():
                self._fused_up[i] = int(m.scale_factor)
        self._graphs = {}
        self.eval()

    @property
    def end2end(self):
        return getattr(self.model[-1], "end2end", False)

    @end2end.setter
    def end2end(self, value):
        self.model[-1].end2end = value

    def fuse(self, verbose=False):
        """No-op kept for API parity: BatchNorm folding happens in each module's weight pack (tasks.py:285-320)."""
        return self

    def train(self, mode=True):
        if mode:
            raise RuntimeError("yolo_master_b200.DetectionModel is inference-only (training is out of scope)")
        return super().train(False)

    # ------------------------------------------------------------------------------------------
    def _predict_once(self, x):
        y = []
        pending_up = 1
        head = self.model[-1]
        # Detect: level l's towers start (on side streams, under graph capture only) the moment layer head.f[l] has produced its map
        early = {f: l for l, f in enumerate(head.f)} if type(head) is M.Detect and isinstance(head.f, list) else {}
        head.__dict__.pop("_early", None)
        for i, m in enumerate(self.model):
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            if i in self._fused_up:
                pending_up = self._fused_up[i]      # defer: the next Concat reads the low-res tensor directly
                y.append(None)
                continue
            if pending_up > 1:
                x = m(x, up_first=pending_up)
                pending_up = 1
            else:
                x = m(x)
            y.append(x if m.i in self.save else None)
            if i in early:
                head.start_level(early[i], x)
        return x

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("yolo_master_b200.DetectionModel.forward needs CUDA tensors (no CPU fallback)")
        if x.dim() != 4:
            raise ValueError(f"expected a (B, C, H, W) image batch, got shape {tuple(x.shape)}")
        if x.shape[0] == 0:                       # empty batch: nothing to launch, shapes as the reference would return them
            return self._empty_result(x)
        return self._predict_once(x)

    def _empty_result(self, x):
        head = self.model[-1]
        H, W = x.shape[2:]
        A = sum(math.ceil(H / float(s)) * math.ceil(W / float(s)) for s in self.stride.tolist())
        if getattr(head, "end2end", False):
            y = torch.zeros((0, min(head.max_det, A), 6), dtype=torch.float32, device=x.device)
        else:
            y = torch.zeros((0, 4 + self.yaml["nc"], A), dtype=torch.float32, device=x.device)
        return y, {"boxes": [], "scores": [], "feats": []}

    predict = forward

    # ------------------------------------------------------------------------------------------
    def graphed(self, batch, height, width, dtype=torch.float16, warmup=3):
        """CUDA-graph runner for a fixed input shape: returns `GraphedForward` (call with a device tensor)."""
        key = (batch, height, width, dtype)
        g = self._graphs.get(key)
        if g is None:
            g = GraphedForward(self, batch, height, width, dtype, warmup)
            self._graphs[key] = g
        return g

    def pipelined(self, batch, height, width, dtype=torch.float16, depth=2, warmup=3):
        """`PipelinedForward`: `depth` graph instances on `depth` streams (throughput mode; results identical to `graphed`)."""
        key = (batch, height, width, dtype, "pipe", depth)
        g = self._graphs.get(key)
        if g is None:
            g = PipelinedForward(self, batch, height, width, dtype, depth, warmup)
            self._graphs[key] = g
        return g


class PipelinedForward:
    """`depth` CUDA-graph instances of the forward (own activation pools, shared weights) on `depth` streams; consecutive batches go to
    consecutive instances.  A single forward leaves most of the GPU idle most of the time - ~160 dependent launches, the long ones bound
    by ONE unit each (area attention by the MUFU, the convolutions by launch latency / HBM) - so two batches in flight overlap the
    exp-bound kernels of one with the memory-bound kernels of the other.  Per-batch arithmetic is untouched (same graph, same kernels):
    results are bit-identical to `GraphedForward`'s; only throughput changes.

        pipe = model.pipelined(32, 640, 640, torch.uint8, depth=2)
        for out in pipe.stream_host(pinned_uint8_batches):      # (B, 300, 6) pinned host tensors, in order
            ...
    """

    def __init__(self, model, batch, height, width, dtype, depth=2, warmup=3):
        dev = next(model.parameters()).device
        self.depth = int(depth)
        self.graphs = [GraphedForward(model, batch, height, width, dtype, warmup) for _ in range(self.depth)]
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(self.depth)]
        self.kernels_per_replay = self.graphs[0].kernels_per_replay
        self.dev = dev
        self._pipe = None

    # ---- device-resident batches -----------------------------------------------------------------
    def run_device(self, batches):
        """Enqueue one forward per device batch (copied into the instance's static input); returns the list of static outputs the
        LAST `depth` batches landed in.  The caller's current stream is joined before returning (not synchronised)."""
        cur = torch.cuda.current_stream(self.dev)
        for s in self.streams:
            s.wait_stream(cur)
        n = 0
        for i, x in enumerate(batches):
            k = i % self.depth
            with torch.cuda.stream(self.streams[k]):
                if x is not None:
                    self.graphs[k].static_in.copy_(x, non_blocking=True)
                self.graphs[k].graph.replay()
            n += 1
        for s in self.streams:
            cur.wait_stream(s)
        return [g.static_out for g in self.graphs[:min(n, self.depth)]]

    # ---- host buffers: H2D -> forward -> D2H per batch, `depth` batches in flight ------------------------------
    def stream_host(self, host_batches):
        """For each pinned host batch (shape / dtype of the captured input) yields the pinned host result, in order.  Batch i runs on
        instance i % depth: its H2D copy goes straight into that instance's static input (on the copy stream, once the instance's previous
        forward has consumed it), the forward on the instance's stream, the D2H on the read-back stream.  A yielded tensor is overwritten
        `depth` batches later."""
        if self._pipe is None:
            ev = lambda: [torch.cuda.Event() for _ in range(self.depth)]
            lead = [g._lead for g in self.graphs]
            self._pipe = {"h2d": torch.cuda.Stream(device=self.dev), "d2h": torch.cuda.Stream(device=self.dev),
                          "host_out": [torch.empty(l.shape, dtype=l.dtype, pin_memory=True) for l in lead],
                          "h2d_done": ev(), "fwd_done": ev(), "d2h_done": ev()}
        p = self._pipe
        pending = []
        for i, hb in enumerate(host_batches):
            k = i % self.depth
            g, cs = self.graphs[k], self.streams[k]
            if i >= self.depth:
                p["h2d"].wait_event(p["fwd_done"][k])          # the instance's previous forward has read its static input
            with torch.cuda.stream(p["h2d"]):
                g.static_in.copy_(hb, non_blocking=True)
                p["h2d_done"][k].record(p["h2d"])
            cs.wait_event(p["h2d_done"][k])
            if i >= self.depth:
                cs.wait_event(p["d2h_done"][k])                # the instance's previous result has been read back
            with torch.cuda.stream(cs):
                g.graph.replay()
                p["fwd_done"][k].record(cs)
            p["d2h"].wait_event(p["fwd_done"][k])
            with torch.cuda.stream(p["d2h"]):
                p["host_out"][k].copy_(g._lead, non_blocking=True)
                p["d2h_done"][k].record(p["d2h"])
            pending.append(k)
            if len(pending) == self.depth:                      # oldest batch in flight: hand it out before its slot is reused
                j = pending.pop(0)
                p["d2h_done"][j].synchronize()
                yield p["host_out"][j]
        for j in pending:
            p["d2h_done"][j].synchronize()
            yield p["host_out"][j]


class PoseModel(DetectionModel):
    """`PoseModel(cfg, ch=3, nc=None, data_kpt_shape=(None, None))` (tasks.py:801-846): a DetectionModel whose head is `Pose`."""

    def __init__(self, cfg="yolo-master-pose-n.yaml", ch=3, nc=None, data_kpt_shape=(None, None), verbose=False):
        cfg = cfg if isinstance(cfg, dict) else yaml_model_load(cfg)
        if any(data_kpt_shape) and list(data_kpt_shape) != list(cfg["kpt_shape"]):
            cfg["kpt_shape"] = list(data_kpt_shape)
        super().__init__(cfg, ch=ch, nc=nc, verbose=verbose)
        self.kpt_shape = tuple(self.yaml["kpt_shape"])


class SegmentationModel(DetectionModel):
    """`SegmentationModel(cfg, ch=3, nc=None)` (tasks.py:775-798): a DetectionModel whose head is `Segment`; the eval forward returns
    ((y, proto), aux)."""


class OBBModel(DetectionModel):
    """`OBBModel(cfg, ch=3, nc=None)` (tasks.py:749-772): a DetectionModel whose head is `OBB`."""


class ClassificationModel(DetectionModel):
    """`ClassificationModel(cfg, ch=3, nc=None)` (tasks.py:913-990): backbone + `Classify`; the eval forward returns (probs, logits).

    The reference's DetectionModel rewrites every BatchNorm2d to eps = 1e-3 (`initialize_weights`, tasks.py:565); its
    ClassificationModel does not, so BatchNorms keep torch's default eps of 1e-5 - mirrored here (the weight packs read `bn.eps`)."""

    def __init__(self, cfg="yolo-master-cls-n.yaml", ch=3, nc=None, verbose=False):
        super().__init__(cfg, ch=ch, nc=nc, verbose=verbose)
        for mod in self.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.eps = 1e-5


class GraphedForward:
    """Whole-forward CUDA graph with static input/output buffers (+ pinned host staging for the host-buffer API)."""

    def __init__(self, model, batch, height, width, dtype, warmup=3):
        dev = next(model.parameters()).device
        self.model = model
        self.static_in = torch.zeros((batch, model.yaml.get("channels", 3), height, width), dtype=dtype, device=dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream), torch.no_grad():
            for _ in range(warmup):
                out = model(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        from .. import ops
        k0 = ops.KERNELS
        with torch.cuda.graph(self.graph, stream=self.stream), torch.no_grad():
            out = model(self.static_in)
        self.kernels_per_replay = ops.KERNELS - k0
        self.static_out = out[0] if isinstance(out, tuple) else out                  # Segment: (prediction, prototypes)
        self._lead = self.static_out[0] if isinstance(self.static_out, tuple) else self.static_out
        self.host_out = torch.empty(self._lead.shape, dtype=self._lead.dtype, pin_memory=True)

    def __call__(self, x=None):
        """x: device tensor of the captured shape (copied into the static input) or None (reuse the static input)."""
        if x is not None:
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out

    def run_host(self, host_in):
        """Host-buffer call: pinned host images -> H2D -> forward -> D2H into `self.host_out`; returns it (synchronised)."""
        self.static_in.copy_(host_in, non_blocking=True)
        self.graph.replay()
        self.host_out.copy_(self._lead, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.host_out

    # ------------------------------------------------------------------------------------------
    def _pipeline(self):
        p = self.__dict__.get("_pipe")
        if p is None:
            dev = self.static_in.device
            ev = lambda: [torch.cuda.Event() for _ in range(2)]
            p = {"h2d": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev),
                 "stage": [torch.empty_like(self.static_in) for _ in range(2)],
                 "dev_out": [torch.empty_like(self.static_out) for _ in range(2)],
                 "host_out": [torch.empty(self.static_out.shape, dtype=self.static_out.dtype, pin_memory=True) for _ in range(2)],
                 "h2d_done": ev(), "consumed": ev(), "fwd_done": ev(), "d2h_done": ev()}
            self.__dict__["_pipe"] = p
        return p

    def stream_host(self, host_batches):
        """Pipelined host-buffer API: for each pinned host batch (shape/dtype of the captured input) yields the pinned host
        result `(B, n, 6)` in order.  The H2D copy of batch i+1 and the D2H copy of batch i-1 run on their own streams while
        batch i computes (double-buffered staging), so a stream of batches runs at max(copy, compute) instead of their sum.
        A yielded tensor is overwritten two batches later: consume or clone it before advancing twice."""
        p = self._pipeline()
        k = torch.cuda.current_stream(self.static_in.device)
        pending = None
        for i, hb in enumerate(host_batches):
            s = i & 1
            if i >= 2:
                p["h2d"].wait_event(p["consumed"][s])          # staging slot s was drained by batch i-2
            with torch.cuda.stream(p["h2d"]):
                p["stage"][s].copy_(hb, non_blocking=True)
                p["h2d_done"][s].record(p["h2d"])
            k.wait_event(p["h2d_done"][s])
            self.static_in.copy_(p["stage"][s], non_blocking=True)
            p["consumed"][s].record(k)
            self.graph.replay()
            if i >= 2:
                k.wait_event(p["d2h_done"][s])                 # dev_out[s] was read back by batch i-2
            p["dev_out"][s].copy_(self.static_out, non_blocking=True)
            p["fwd_done"][s].record(k)
            p["d2h"].wait_event(p["fwd_done"][s])
            with torch.cuda.stream(p["d2h"]):
                p["host_out"][s].copy_(p["dev_out"][s], non_blocking=True)
                p["d2h_done"][s].record(p["d2h"])
            if pending is not None:
                p["d2h_done"][pending].synchronize()
                yield p["host_out"][pending]
            pending = s
        if pending is not None:
            p["d2h_done"][pending].synchronize()
            yield p["host_out"][pending]
