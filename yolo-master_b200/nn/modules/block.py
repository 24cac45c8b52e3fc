"""CSP / attention blocks with the reference's names, signatures and state_dict keys.

Mirrors `ultralytics/nn/modules/block.py`: Bottleneck :462-486, C2f :293-324, C3 :327-350, C3k :1114-1131,
C3k2 :1074-1108, SPPF :213-242, Attention :1276-1333, PSABlock :1336-1383, C2PSA :1441-1493, AAttn :1646-1732,
ABlock :1735-1797, A2C2f :1800-1879.  Every `torch.cat`/`chunk`/`split` of the reference is a channel slice of one
NHWC buffer here: producers write straight into their slice (`out=` views), so no concat copies are executed.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ._base import cached_f32, cached_pack, fold_bn, fwd_child, pack_gemm_weight, require_eval, to_nchw, to_nhwc
from .conv import Conv

__all__ = ("Bottleneck", "C2f", "C3", "C3k", "C3k2", "SPPF", "Attention", "PSABlock", "C2PSA", "AAttn", "ABlock", "A2C2f")


class _NHWCBlock(nn.Module):
    """Boundary adapter: NCHW-logical in/out, NHWC inside."""

    def forward(self, x):
        require_eval(self)
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))


class Bottleneck(_NHWCBlock):
    """`Bottleneck(c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5)`."""

    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def fwd_nhwc(self, x, out=None):
        # x + cv2(cv1(x)): the shortcut add rides in cv2's epilogue
        return self.cv2.fwd_nhwc(self.cv1.fwd_nhwc(x), out=out, res=x if self.add else None)


class C2f(_NHWCBlock):
    """`C2f(c1, c2, n=1, shortcut=False, g=1, e=0.5)`."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, g, k=((3, 3), (3, 3)), e=1.0) for _ in range(n))

    def fwd_nhwc(self, x, out=None):
        B, H, W, _ = x.shape
        c, n = self.c, len(self.m)
        cat = ops.new_act(B, H, W, (2 + n) * c, x.device)  # [y0 | y1 | m0(y1) | m1(..) ...]
        self.cv1.fwd_nhwc(x, out=cat[..., : 2 * c])
        for j, m in enumerate(self.m):
            fwd_child(m, cat[..., (1 + j) * c:(2 + j) * c], out=cat[..., (2 + j) * c:(3 + j) * c])
        return self.cv2.fwd_nhwc(cat, out=out)


class C3(_NHWCBlock):
    """`C3(c1, c2, n=1, shortcut=True, g=1, e=0.5)`."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=((1, 1), (3, 3)), e=1.0) for _ in range(n)))

    def fwd_nhwc(self, x, out=None):
        B, H, W, _ = x.shape
        c_ = self.cv1.conv.out_channels
        cat = ops.new_act(B, H, W, 2 * c_, x.device)
        t = self.cv1.fwd_nhwc(x)
        for j, m in enumerate(self.m):
            t = m.fwd_nhwc(t, out=cat[..., :c_] if j == len(self.m) - 1 else None)
        if len(self.m) == 0:
            cat[..., :c_].copy_(t)
        self.cv2.fwd_nhwc(x, out=cat[..., c_:])
        return self.cv3.fwd_nhwc(cat, out=out)


class C3k(C3):
    """`C3k(c1, c2, n=1, shortcut=True, g=1, e=0.5, k=3)`."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, k=3):
        super().__init__(c1, c2, n, shortcut, g, e)
        c_ = int(c2 * e)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=(k, k), e=1.0) for _ in range(n)))


class _SeqNHWC(nn.Sequential):
    def fwd_nhwc(self, x, out=None):
        mods = list(self)
        for j, m in enumerate(mods):
            x = m.fwd_nhwc(x, out=out if j == len(mods) - 1 else None)
        return x


class Attention(_NHWCBlock):
    """`Attention(dim, num_heads=8, attn_ratio=0.5)` (PSA attention, key_dim = attn_ratio*head_dim)."""

    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        nh_kd = self.key_dim * num_heads
        h = dim + nh_kd * 2
        self.qkv = Conv(dim, h, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)

    def fwd_nhwc(self, x, out=None, res=None):
        B, H, W, C = x.shape
        kd, hd, nh = self.key_dim, self.head_dim, self.num_heads
        hs = 2 * kd + hd
        qkv = self.qkv.fwd_nhwc(x)  # per head [q kd | k kd | v hd]
        o = ops.attention(qkv, B, H * W, nh, hs, 0, kd, 2 * kd, kd, hd, self.scale)
        pe = self.pe.get_pack()
        ops.dwconv(qkv, pe["w"], pe["bias"], self.pe.conv.kernel_size[0], False, C, grp_w=hd, grp_stride=hs, grp_off=2 * kd,
                   add=o, out=o)  # o += pe(v), v read in place from qkv
        return self.proj.fwd_nhwc(o, out=out, res=res)


class PSABlock(_NHWCBlock):
    """`PSABlock(c, attn_ratio=0.5, num_heads=4, shortcut=True)`."""

    def __init__(self, c, attn_ratio=0.5, num_heads=4, shortcut=True):
        super().__init__()
        self.attn = Attention(c, attn_ratio=attn_ratio, num_heads=num_heads)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))
        self.add = shortcut

    def fwd_nhwc(self, x, out=None):
        x = self.attn.fwd_nhwc(x, res=x if self.add else None)
        return self.ffn[1].fwd_nhwc(self.ffn[0].fwd_nhwc(x), out=out, res=x if self.add else None)


class C3k2(C2f):
    """`C3k2(c1, c2, n=1, c3k=False, e=0.5, attn=False, g=1, shortcut=True)`."""

    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, attn=False, g=1, shortcut=True):
        super().__init__(c1, c2, n, shortcut, g, e)
        self.m = nn.ModuleList(
            _SeqNHWC(Bottleneck(self.c, self.c, shortcut, g), PSABlock(self.c, attn_ratio=0.5, num_heads=max(self.c // 64, 1)))
            if attn
            else C3k(self.c, self.c, 2, shortcut, g)
            if c3k
            else Bottleneck(self.c, self.c, shortcut, g)
            for _ in range(n)
        )


class SPPF(_NHWCBlock):
    """`SPPF(c1, c2, k=5, n=3, shortcut=False)`."""

    def __init__(self, c1, c2, k=5, n=3, shortcut=False):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1, act=False)
        self.cv2 = Conv(c_ * (n + 1), c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)
        self.n = n
        self.add = shortcut and c1 == c2

    def fwd_nhwc(self, x, out=None):
        if self.n != 3:
            raise NotImplementedError("SPPF: the pooling kernel implements n=3 (every master YAML)")
        B, H, W, _ = x.shape
        c_ = self.cv1.conv.out_channels
        cat = ops.new_act(B, H, W, 4 * c_, x.device)
        self.cv1.fwd_nhwc(x, out=cat[..., :c_])
        ops.sppf_pool(cat, c_, self.m.kernel_size)
        return self.cv2.fwd_nhwc(cat, out=out, res=x if self.add else None)


class C2PSA(_NHWCBlock):
    """`C2PSA(c1, c2, n=1, e=0.5)`."""

    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))

    def fwd_nhwc(self, x, out=None):
        c = self.c
        ab = self.cv1.fwd_nhwc(x)  # [a | b]
        b = ab[..., c:]
        mods = list(self.m)
        for j, m in enumerate(mods):
            b = m.fwd_nhwc(b, out=ab[..., c:] if j == len(mods) - 1 else None)  # last block writes b back in place
        return self.cv2.fwd_nhwc(ab, out=out)


class AAttn(_NHWCBlock):
    """`AAttn(dim, num_heads, area=1)`: area attention; q/k/v interleaved per head, pe (dw 7x7) on V."""

    def __init__(self, dim, num_heads, area=1):
        super().__init__()
        self.area = area
        self.num_heads = num_heads
        self.head_dim = head_dim = dim // num_heads
        self.all_head_dim = all_head_dim = head_dim * self.num_heads
        self.qkv = Conv(dim, all_head_dim * 3, 1, act=False)
        self.proj = Conv(all_head_dim, dim, 1, act=False)
        self.pe = Conv(all_head_dim, all_head_dim, 7, 1, 3, g=all_head_dim, act=False)

    def fwd_nhwc(self, x, out=None, res=None):
        B, H, W, _ = x.shape
        hd, nh, C = self.head_dim, self.num_heads, self.all_head_dim
        N = H * W
        if N % self.area != 0:
            raise ValueError(f"AAttn: H*W={N} is not divisible by area={self.area}")
        qkv = self.qkv.fwd_nhwc(x)  # (B,H,W,3C), per head [q hd | k hd | v hd]
        o = ops.attention(qkv, B * self.area, N // self.area, nh, 3 * hd, 0, hd, 2 * hd, hd, hd, hd ** -0.5)
        pe = self.pe.get_pack()
        ops.dwconv(qkv, pe["w"], pe["bias"], self.pe.conv.kernel_size[0], False, C, grp_w=hd, grp_stride=3 * hd, grp_off=2 * hd,
                   add=o, out=o)
        return self.proj.fwd_nhwc(o, out=out, res=res)


class ABlock(_NHWCBlock):
    """`ABlock(dim, num_heads, mlp_ratio=1.2, area=1)`."""

    def __init__(self, dim, num_heads, mlp_ratio=1.2, area=1):
        super().__init__()
        self.attn = AAttn(dim, num_heads=num_heads, area=area)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = nn.Sequential(Conv(dim, mlp_hidden_dim, 1), Conv(mlp_hidden_dim, dim, 1, act=False))
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Conv2d):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def _padded_mlp(self):
        """mlp_ratio 1.2 (l/x scales) gives hidden widths such as int(256*1.2) = 307: zero-pad the hidden channels to a
        multiple of 8 (padded units are SiLU(0) = 0 and meet zero columns of the second conv, so nothing changes)."""
        c0, c1 = self.mlp[0], self.mlp[1]

        def build():
            w0, b0 = fold_bn(c0.conv.weight, None, c0.bn)
            w1, b1 = fold_bn(c1.conv.weight, None, c1.bn)
            hid = w0.shape[0]
            hp = (hid + 7) // 8 * 8
            w0p = torch.zeros((hp, *w0.shape[1:]), device=w0.device)
            w0p[:hid] = w0
            b0p = torch.zeros(hp, device=w0.device)
            b0p[:hid] = b0
            w1p = torch.zeros((w1.shape[0], hp, 1, 1), device=w1.device)
            w1p[:, :hid] = w1
            return {"w0": pack_gemm_weight(w0p), "b0": b0p.contiguous(), "w1": pack_gemm_weight(w1p), "b1": b1.contiguous(), "hp": hp}

        return cached_pack(self, "mlp", [*c0.parameters(), *c0.buffers(), *c1.parameters(), *c1.buffers()], build)

    def fwd_nhwc(self, x, out=None):
        x = self.attn.fwd_nhwc(x, res=x)
        if self.mlp[0].conv.out_channels % 8:
            pk = self._padded_mlp()
            h = ops.conv2d(x, pk["w0"], pk["b0"], pk["hp"], 1, 1, 1, 0, True)
            return ops.conv2d(h, pk["w1"], pk["b1"], self.mlp[1].conv.out_channels, 1, 1, 1, 0, False, out=out, res=x)
        return self.mlp[1].fwd_nhwc(self.mlp[0].fwd_nhwc(x), out=out, res=x)


class A2C2f(_NHWCBlock):
    """`A2C2f(c1, c2, n=1, a2=True, area=1, residual=False, mlp_ratio=2.0, e=0.5, g=1, shortcut=True)`."""

    def __init__(self, c1, c2, n=1, a2=True, area=1, residual=False, mlp_ratio=2.0, e=0.5, g=1, shortcut=True):
        super().__init__()
        c_ = int(c2 * e)
        assert c_ % 32 == 0, "Dimension of ABlock must be a multiple of 32."
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv((1 + n) * c_, c2, 1)
        self.gamma = nn.Parameter(0.01 * torch.ones(c2), requires_grad=True) if a2 and residual else None
        self.m = nn.ModuleList(
            _SeqNHWC(*(ABlock(c_, c_ // 32, mlp_ratio, area) for _ in range(2))) if a2 else C3k(c_, c_, 2, shortcut, g)
            for _ in range(n)
        )

    def fwd_nhwc(self, x, out=None):
        B, H, W, _ = x.shape
        c_, n = self.cv1.conv.out_channels, len(self.m)
        cat = ops.new_act(B, H, W, (1 + n) * c_, x.device)
        self.cv1.fwd_nhwc(x, out=cat[..., :c_])
        for j, m in enumerate(self.m):
            fwd_child(m, cat[..., j * c_:(j + 1) * c_], out=cat[..., (j + 1) * c_:(j + 2) * c_])
        if self.gamma is not None:      # x + gamma * y  (block.py:1877-1879)
            y = self.cv2.fwd_nhwc(cat)
            return ops.ew(ops.EW_SCALE_RES, a=x, b=y, p0=cached_f32(self, 'gamma', self.gamma, (-1,)), out=out)
        return self.cv2.fwd_nhwc(cat, out=out)
