"""Conv / DWConv / Concat / Upsample with the reference's names, signatures and state_dict keys.

Mirrors `ultralytics/nn/modules/conv.py` (Conv :39-89, DWConv :185-199, Concat :616-640, autopad :30-36).  Parameters
live in the same `conv` / `bn` children so released checkpoints load; the forward runs hand-written kernels
(`ops.conv2d`, `ops.dwconv`, `ops.stem_conv`) on NHWC fp16 with BatchNorm folded at weight-prep time.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops
from ._base import PackCache, fold_bn, pack_gemm_weight, pad_channels8, plain_conv_run, require_eval, to_nchw, to_nhwc

__all__ = ("Conv", "DWConv", "Concat", "Upsample", "autopad", "PlainConv2d")


def autopad(k, p=None, d=1):
    """Pad to 'same' shape outputs (reference conv.py:30-36)."""
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (x - 1) + 1 for x in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


class Conv(nn.Module, PackCache):
    """Convolution + BatchNorm + activation: `Conv(c1, c2, k=1, s=1, p=None, g=1, d=1, act=True)`."""

    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    # ---- weight preparation -------------------------------------------------------------------
    def _kind(self):
        cv = self.conv
        k = cv.kernel_size[0]
        if cv.kernel_size[0] != cv.kernel_size[1] or cv.dilation != (1, 1) or cv.stride[0] != cv.stride[1]:
            raise NotImplementedError("Conv: only square kernels, equal strides and dilation 1 are on the B200 path")
        if isinstance(self.act, nn.SiLU):
            act = True
        elif isinstance(self.act, nn.Identity):
            act = False
        else:
            raise NotImplementedError(f"Conv: activation {type(self.act).__name__} is not on the B200 path (SiLU / Identity)")
        if cv.groups == 1:
            kind = "stem" if (cv.in_channels <= 4 and k == 3 and cv.stride[0] == 2 and cv.padding[0] == 1) else "gemm"
        elif cv.groups == cv.in_channels == cv.out_channels and cv.stride[0] == 1 and cv.padding[0] == k // 2:
            kind = "dw"
        else:
            raise NotImplementedError(f"Conv: groups={cv.groups} (c1={cv.in_channels}, c2={cv.out_channels}) is not on the B200 path")
        return kind, act

    def _build_pack(self):
        kind, act = self._kind()
        w, b = fold_bn(self.conv.weight, self.conv.bias, getattr(self, "bn", None))
        pk = {"kind": kind, "act": act, "bias": b.contiguous()}
        if kind == "gemm":
            w, pk["bias"], pk["cin_p"], pk["cout_p"] = pad_channels8(w, pk["bias"])
            pk["w"] = pack_gemm_weight(w)
        elif kind == "dw":
            C, _, k, _ = w.shape
            pk["w"] = w.reshape(C, k * k).t().contiguous().half()  # tap-major [k*k][C]
        else:  # stem: fp32 [Cin*9][Cout], k index = (ci*3+ky)*3+kx; HOST copies (they ride in the kernel's parameter block)
            Co = w.shape[0]
            pk["w"] = w.reshape(Co, -1).t().contiguous().cpu()
            pk["bias"] = pk["bias"].cpu()
        return pk

    # ---- forward ------------------------------------------------------------------------------
    def fwd_nhwc(self, x, out=None, res=None):
        """x: (B,H,W,c1) fp16 view.  Optional `out` view to write into and `res` to add after the activation."""
        require_eval(self)
        pk = self.get_pack()
        cv = self.conv
        if pk["kind"] == "gemm":
            if pk["cin_p"] != cv.in_channels or pk["cout_p"] != cv.out_channels:
                # widths that are not multiples of 8 (Pose towers: 51 keypoint channels) travel zero-padded to the next multiple of 8:
                # the producer's padded output IS this conv's input, padded weights / bias keep the pad channels at exactly zero
                if x.shape[-1] != pk["cin_p"] or out is not None or res is not None:
                    raise NotImplementedError(f"Conv({cv.in_channels}->{cv.out_channels}): odd channel widths run zero-padded to a "
                                              "multiple of 8 inside a tower only (no out= / res= views, input from a padded producer)")
            return ops.conv2d(x, pk["w"], pk["bias"], pk["cout_p"], cv.kernel_size[0], cv.kernel_size[1], cv.stride[0],
                              cv.padding[0], pk["act"], out=out, res=res)
        if pk["kind"] == "dw":
            return ops.dwconv(x, pk["w"], pk["bias"], cv.kernel_size[0], pk["act"], cv.out_channels, add=res, out=out)
        raise RuntimeError("Conv: the stem kernel takes the NCHW image; call forward() instead of fwd_nhwc()")

    def forward(self, x):
        require_eval(self)
        pk = self.get_pack()
        if pk["kind"] == "stem":
            if not x.is_cuda:
                raise RuntimeError("yolo_master_b200.Conv runs on CUDA tensors only (no CPU fallback)")
            if not pk["act"]:
                raise NotImplementedError("stem Conv without SiLU is not on the B200 path")
            return to_nchw(ops.stem_conv(x, pk["w"], pk["bias"], self.conv.out_channels))
        y = to_nchw(self.fwd_nhwc(to_nhwc(x)))
        return y if y.shape[1] == self.conv.out_channels else y[:, :self.conv.out_channels]

    forward_fuse = forward


class DWConv(Conv):
    """Depth-wise convolution: `DWConv(c1, c2, k=1, s=1, d=1, act=True)` (reference conv.py:185-199)."""

    def __init__(self, c1, c2, k=1, s=1, d=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), d=d, act=act)


class PlainConv2d(nn.Conv2d, PackCache):
    """`nn.Conv2d` with bias and no norm/activation (Detect's last 1x1, head.py:104-119) on the GEMM kernel.

    Subclasses nn.Conv2d so state_dict keys (`weight`, `bias`) match the reference's bare nn.Conv2d."""

    def _build_pack(self):
        if self.groups != 1 or self.dilation != (1, 1):
            raise NotImplementedError("PlainConv2d: groups/dilation not supported")
        w, b = fold_bn(self.weight, self.bias, None)
        w, b, cin_p, cout_p = pad_channels8(w, b.contiguous(), cout_to=2)
        return {"w": pack_gemm_weight(w), "bias": b, "cin_p": cin_p, "cout_p": cout_p}

    def fwd_nhwc(self, x, out=None, out_f32=False):
        return plain_conv_run(self, self.get_pack(), x, out, out_f32)

    def forward(self, x):
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))


class Concat(nn.Module):
    """Concatenate a list of tensors along dimension 1: `Concat(dimension=1)` (reference conv.py:616-640)."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x, up_first: int = 1):
        if self.d != 1:
            raise NotImplementedError("Concat: only channel concatenation is on the B200 path")
        xs = [to_nhwc(t) for t in x]
        y = ops.concat2(xs[0], xs[1] if len(xs) > 1 else None, up=up_first)
        for t in xs[2:]:
            y = ops.concat2(y, t)
        return to_nchw(y)


class Upsample(nn.Upsample):
    """`nn.Upsample(None, s, 'nearest')` on the copy kernel (fused into the following Concat by DetectionModel)."""

    def forward(self, x):
        s = int(self.scale_factor) if self.scale_factor is not None else 0
        if self.mode != "nearest" or s < 1 or s != self.scale_factor:
            raise NotImplementedError("Upsample: only integer nearest-neighbour scaling is on the B200 path")
        return to_nchw(ops.concat2(to_nhwc(x), None, up=s))
