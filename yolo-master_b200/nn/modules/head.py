"""Detect head with the reference's name, signature and state_dict keys (`ultralytics/nn/modules/head.py:37-262`).

Towers run on the conv kernels (last 1x1 writes fp32 logits / distances), then ONE kernel per batch does
anchors + dist2bbox + sigmoid + the end2end two-stage top-k (head.py:173-258), or the dense decode for the NMS path.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from ... import ops
from ._base import PackCache, cached_f32, fwd_child, pack_gemm_weight, plain_conv_fwd, require_eval, run_branches, to_nchw, to_nhwc
from .conv import Conv, DWConv, PlainConv2d

__all__ = ("Detect", "DFL", "Pose", "Proto", "Segment", "OBB", "Classify")


_EARLY_STREAMS: dict = {}


class DFL(nn.Module):
    """`DFL(c1=16)` (block.py:63-85): frozen 1x1 conv with weights 0..c1-1 = expectation over the softmaxed bins.
    Only holds the state_dict key (`dfl.conv.weight`); the arithmetic is fused into `ym_detect_dense`."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def check_frozen(self):
        w = self.conv.weight.detach().float().view(-1).cpu()
        if not torch.equal(w, torch.arange(self.c1, dtype=torch.float)):
            raise NotImplementedError("DFL: conv.weight differs from arange(reg_max); only the reference's frozen DFL is supported")


class Detect(nn.Module):
    """`Detect(nc=80, reg_max=16, end2end=False, ch=())`."""

    dynamic = False
    export = False
    format = None
    max_det = 300
    agnostic_nms = False
    shape = None
    legacy = False
    xyxy = False

    def __init__(self, nc=80, reg_max=16, end2end=False, ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = reg_max
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), PlainConv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = (
            nn.ModuleList(nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), PlainConv2d(c3, self.nc, 1)) for x in ch)
            if self.legacy
            else nn.ModuleList(
                nn.Sequential(
                    nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                    nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)),
                    PlainConv2d(c3, self.nc, 1),
                )
                for x in ch
            )
        )
        self.dfl = DFL(self.reg_max) if self.reg_max > 1 else nn.Identity()
        self._dfl_checked = None
        if end2end:
            self.one2one_cv2 = copy.deepcopy(self.cv2)
            self.one2one_cv3 = copy.deepcopy(self.cv3)

    @property
    def end2end(self):
        return getattr(self, "_end2end", True) and hasattr(self, "one2one_cv2")

    @end2end.setter
    def end2end(self, value):
        self._end2end = value

    def fuse(self):
        """Reference Detect.fuse() drops the one2many towers (head.py:260-262); they are simply not executed here."""

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _tower(seq, x, out_f32=True):
        mods = list(seq)
        for m in mods[:-1]:
            x = fwd_child(m, x)
        return plain_conv_fwd(mods[-1], x, out_f32=out_f32)    # PlainConv2d here, a bare nn.Conv2d on reference-built instances

    def head_raw(self, feats_nhwc):
        """Per level (boxes fp32 (B,h,w,4*reg_max), logits fp32 (B,h,w,nc)) from the active towers."""
        e2e = self.end2end
        box_head = self.one2one_cv2 if e2e else self.cv2
        cls_head = self.one2one_cv3 if e2e else self.cv3
        n = len(feats_nhwc)
        early = self.__dict__.pop("_early", None)
        if early is not None and len(early["out"]) == n and all(early["src"][i] == feats_nhwc[i].data_ptr() for i in range(n)):
            # towers were started level by level while the neck was still running (start_level): join their streams
            cur = torch.cuda.current_stream()
            for st in early["streams"]:
                cur.wait_stream(st)
            return [early["out"][i][0] for i in range(n)], [early["out"][i][1] for i in range(n)]
        # 2 x nl independent towers: parallel graph branches under capture (largest maps first), serial otherwise
        jobs = [(lambda i=i: self._tower(cls_head[i], feats_nhwc[i])) for i in range(n)] + \
               [(lambda i=i: self._tower(box_head[i], feats_nhwc[i])) for i in range(n)]
        res = run_branches(jobs)
        return res[n:], res[:n]

    def start_level(self, level: int, feat_nchw: torch.Tensor) -> bool:
        """Under CUDA-graph capture: start the box / class towers of pyramid level `level` on two side streams as soon as its feature map
        exists, so that they overlap the rest of the neck (the P3 towers are the longest chains of the head and their input is ready
        six layers before the P5 map).  `head_raw` joins the streams.  Returns False (and does nothing) outside a capture."""
        if not (feat_nchw.is_cuda and torch.cuda.is_current_stream_capturing()) or self.training:
            return False
        e2e = self.end2end
        box_head = self.one2one_cv2 if e2e else self.cv2
        cls_head = self.one2one_cv3 if e2e else self.cv3
        f = to_nhwc(feat_nchw)
        cur = torch.cuda.current_stream()
        key = ("detect_early", cur.device.index)
        pool = _EARLY_STREAMS.setdefault(key, [torch.cuda.Stream(device=cur.device) for _ in range(2 * self.nl)])
        early = self.__dict__.setdefault("_early", {"out": {}, "src": {}, "streams": []})
        outs = []
        for j, tower in enumerate((box_head[level], cls_head[level])):
            st = pool[2 * level + j]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs.append(self._tower(tower, f))
            early["streams"].append(st)
        early["out"][level], early["src"][level] = (outs[0], outs[1]), f.data_ptr()
        return True

    def forward(self, x):
        require_eval(self)
        if self.reg_max > 1:
            if self.end2end:
                raise NotImplementedError("Detect: end2end top-k with DFL (reg_max > 1) is not on the B200 path")
            ver = self.dfl.conv.weight._version
            if getattr(self, "_dfl_checked", None) != ver:          # one host read per weight version, never inside a captured forward
                self.dfl.check_frozen()
                self._dfl_checked = ver
        if self.agnostic_nms:
            raise NotImplementedError("Detect: agnostic_nms top-k is not on the B200 path")
        feats = [to_nhwc(f) for f in x]
        boxes, logits = self.head_raw(feats)
        strides = [float(s) for s in self.stride.tolist()]
        if any(s <= 0 for s in strides):
            raise RuntimeError("Detect.stride is not initialised (build the model through DetectionModel)")
        if self.end2end:
            y = ops.detect_topk(boxes, logits, strides, self.nc, self.max_det)
        else:
            y = ops.detect_dense(boxes, logits, strides, self.nc, xyxy=self.xyxy, reg_max=self.reg_max)
        if self.export:
            return y
        return y, {"boxes": boxes, "scores": logits, "feats": x}


class Pose(Detect):
    """`Pose(nc=80, kpt_shape=(17, 3), reg_max=16, end2end=False, ch=())` (head.py:558-664): Detect plus a keypoint tower per level;
    the dense output gains `nk` rows of decoded keypoints (`ym_kpts_decode`): (B, 4 + nc + nk, A)."""

    def __init__(self, nc=80, kpt_shape=(17, 3), reg_max=16, end2end=False, ch=()):
        super().__init__(nc, reg_max, end2end, ch)
        self.kpt_shape = tuple(kpt_shape)
        self.nk = self.kpt_shape[0] * self.kpt_shape[1]
        c4 = max(ch[0] // 4, self.nk)
        self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), PlainConv2d(c4, self.nk, 1)) for x in ch)
        if end2end:
            self.one2one_cv4 = copy.deepcopy(self.cv4)

    def forward(self, x):
        if self.end2end:
            raise NotImplementedError("Pose: the end2end (one2one) head is not on the B200 path")
        y, aux = super().forward(x)
        feats = [to_nhwc(f) for f in x]
        kpts = [self._tower(self.cv4[i], f) for i, f in enumerate(feats)]
        ky = ops.kpts_decode(kpts, [float(s) for s in self.stride.tolist()], self.kpt_shape[1])
        aux["kpts"] = kpts
        return torch.cat([y, ky], 1), aux


class Proto(nn.Module, PackCache):
    """`Proto(c1, c_=256, c2=32)` (block.py:88-107): Conv 3x3 -> ConvTranspose2d(2, stride 2) -> Conv 3x3 -> Conv 1x1.
    A 2x2 stride-2 transposed convolution writes each input pixel to its own 2x2 output block, so it runs as a 1x1 convolution
    to 4*c_ channels (one group per (dy, dx)) followed by a depth-to-space rearrangement."""

    def __init__(self, c1, c_=256, c2=32):
        super().__init__()
        self.cv1 = Conv(c1, c_, k=3)
        self.upsample = nn.ConvTranspose2d(c_, c_, 2, 2, 0, bias=True)
        self.cv2 = Conv(c_, c_, k=3)
        self.cv3 = Conv(c_, c2)

    def _pack_sources(self):
        return [self.upsample.weight, self.upsample.bias]

    def _build_pack(self):
        w = self.upsample.weight.detach().float()                          # [ci][co][dy][dx]
        c = w.shape[1]
        w4 = w.permute(2, 3, 1, 0).reshape(4 * c, w.shape[0], 1, 1)        # row (dy*2+dx)*c + co
        return {"w": pack_gemm_weight(w4), "b": self.upsample.bias.detach().float().repeat(4).contiguous(), "c": c}

    def fwd_nhwc(self, x):
        pk = self.get_pack()
        t = self.cv1.fwd_nhwc(x)
        B, H, W, _ = t.shape
        c = pk["c"]
        t = ops.conv2d(t, pk["w"], pk["b"], 4 * c, 1, 1, 1, 0, False)      # (B,H,W,[dy][dx][c])
        t = t.view(B, H, W, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, c)   # depth-to-space (a copy)
        return self.cv3.fwd_nhwc(self.cv2.fwd_nhwc(t))

    def forward(self, x):
        require_eval(self)
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))


class Segment(Detect):
    """`Segment(nc=80, nm=32, npr=256, reg_max=16, end2end=False, ch=())` (head.py:265-346): Detect plus a mask-coefficient tower per
    level and a prototype branch on the first level.  Eval output: ((y (B, 4 + nc + nm, A), proto (B, nm, 2h, 2w)), aux)."""

    def __init__(self, nc=80, nm=32, npr=256, reg_max=16, end2end=False, ch=()):
        super().__init__(nc, reg_max, end2end, ch)
        self.nm, self.npr = nm, npr
        self.proto = Proto(ch[0], self.npr, self.nm)
        c4 = max(ch[0] // 4, self.nm)
        self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), PlainConv2d(c4, self.nm, 1)) for x in ch)
        if end2end:
            self.one2one_cv4 = copy.deepcopy(self.cv4)

    def forward(self, x):
        if self.end2end:
            raise NotImplementedError("Segment: the end2end (one2one) head is not on the B200 path")
        y, aux = super().forward(x)
        feats = [to_nhwc(f) for f in x]
        mc = ops.kpts_decode([self._tower(self.cv4[i], f) for i, f in enumerate(feats)], [float(s) for s in self.stride.tolist()], 1)
        proto = to_nchw(self.proto.fwd_nhwc(feats[0]))
        aux["mask_coefficient"], aux["proto"] = mc, proto
        return (torch.cat([y, mc], 1), proto), aux


class OBB(Detect):
    """`OBB(nc=80, ne=1, reg_max=16, end2end=False, ch=())` (head.py:428-520): Detect plus an angle tower per level; boxes are decoded
    as rotated (cx, cy, w, h) and the angle is appended: (B, 4 + nc + 1, A).  Rotated NMS is left to the caller."""

    def __init__(self, nc=80, ne=1, reg_max=16, end2end=False, ch=()):
        super().__init__(nc, reg_max, end2end, ch)
        if ne != 1:
            raise NotImplementedError("OBB: only one extra parameter (the angle) is on the B200 path")
        self.ne = ne
        c4 = max(ch[0] // 4, self.ne)
        self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), PlainConv2d(c4, self.ne, 1)) for x in ch)
        if end2end:
            self.one2one_cv4 = copy.deepcopy(self.cv4)

    def forward(self, x):
        if self.end2end:
            raise NotImplementedError("OBB: the end2end (one2one) head is not on the B200 path")
        y, aux = super().forward(x)
        feats = [to_nhwc(f) for f in x]
        angles = [self._tower(self.cv4[i], f) for i, f in enumerate(feats)]
        aux["angle"] = angles
        return ops.obb_finish(y, angles, [float(s) for s in self.stride.tolist()], self.nc), aux


class Classify(nn.Module):
    """`Classify(c1, c2, k=1, s=1, p=None, g=1)` (head.py:791-832): Conv to 1280 channels -> global average pool -> Linear -> softmax.
    Eval output: (probabilities (B, c2), logits (B, c2))."""

    export = False

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1):
        super().__init__()
        c_ = 1280
        self.conv = Conv(c1, c_, k, s, p, g)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.drop = nn.Dropout(p=0.0, inplace=True)
        self.linear = nn.Linear(c_, c2)

    def forward(self, x):
        require_eval(self)
        if isinstance(x, list):
            x = torch.cat(x, 1)
        t = self.conv.fwd_nhwc(to_nhwc(x))
        v = ops.gap(t)
        w = cached_f32(self, "w", self.linear.weight)
        b = cached_f32(self, "b", self.linear.bias) if self.linear.bias is not None else None
        y, logits = ops.classify_head(v, w, b)
        return y if self.export else (y, logits)
