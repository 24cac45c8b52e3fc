"""Mixture-of-Attention blocks with the reference's names, signatures and state_dict keys
(`ultralytics/nn/modules/moa/{heads,router,block,wrappers}.py`), eval forward, dense soft routing (sparse_inference=False).

Per `MoABlock`: `ym_token_router` (soft, fp32) -> local head (depthwise-biased qkv, 7x7-window attention), regional head (queries at
full resolution, keys/values on an adaptive-average-pooled map), global head (exact attention for N <= 448, Performer ReLU-feature
linear attention in fp32 for N > 512, linear blend in between) -> each head's GroupNorm is applied together with its per-token
routing weight and accumulated in one fused kernel -> fusion 1x1 (layer-scale folded, residual in the epilogue) -> FFN.
Head dims that are not multiples of 8 (e.g. 21 at the s scale) are zero-padded to the next multiple of 8 in the packed weights:
padded q/k columns contribute 0 to every score and padded v/proj columns are 0, so results are unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ._base import PackCache, fold_bn, require_eval, to_nchw, to_nhwc
from .conv import Conv
from .moe import get_safe_groups
from .mot import _f32, _lin, _pack_dw, _pack_linear

__all__ = ("C2fMoA", "MoABlock")

LINEAR_ATTN_THRESHOLD, LINEAR_ATTN_BLEND_WINDOW, LINEAR_ATTN_ACTIVATION_LIMIT = 512, 64, 1e4   # moa/_constants.py
DEFAULT_RF_SEED = 0x5F3759DF


def _pad8(n):
    return (n + 7) // 8 * 8


def _pad_heads_out(w, groups, nh, hd, hdp):
    """Rows of w are `groups` blocks of nh heads x hd channels: zero-pad every head to hdp rows."""
    if hd == hdp:
        return w
    tail = w.shape[1:]
    w = w.reshape(groups, nh, hd, *tail)
    z = torch.zeros((groups, nh, hdp - hd, *tail), dtype=w.dtype, device=w.device)
    return torch.cat([w, z], 2).reshape(groups * nh * hdp, *tail)


def _pad_heads_in(w, nh, hd, hdp):
    """Columns of w ([Co, nh*hd]) -> [Co, nh*hdp] with zero columns."""
    if hd == hdp:
        return w
    Co = w.shape[0]
    w = w.reshape(Co, nh, hd)
    return torch.cat([w, torch.zeros((Co, nh, hdp - hd), dtype=w.dtype, device=w.device)], 2).reshape(Co, nh * hdp)


class _Head(nn.Module, PackCache):
    def _norm_pack(self):
        return (_f32(self.norm.weight), _f32(self.norm.bias))

    def forward(self, x):
        """Stand-alone use: un-weighted head output GN(proj(attn))."""
        require_eval(self)
        xh = to_nhwc(x)
        p = self.head_raw(xh)
        return to_nchw(ops.groupnorm(p, self.norm.num_groups, *self.get_pack()["norm"], eps=self.norm.eps))


class _LocalAttnHead(_Head):
    """`_LocalAttnHead(dim, num_heads, head_dim=None, window_size=7)` (moa/heads.py:124-159)."""

    def __init__(self, dim, num_heads, head_dim=None, window_size=7):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = head_dim or max(dim // num_heads, 16)
        self.window_size = max(1, int(window_size))
        inner = self.head_dim * num_heads
        self.qkv_dw = nn.Conv2d(dim, dim, 3, padding=1, groups=dim, bias=False)
        self.qkv_pw = nn.Conv2d(dim, inner * 3, 1, bias=False)
        self.proj = nn.Conv2d(inner, dim, 1, bias=False)
        self.pe = nn.Conv2d(inner, inner, 7, padding=3, groups=inner, bias=False)
        self.norm = nn.GroupNorm(get_safe_groups(dim, 8), dim)
        self.scale = self.head_dim ** -0.5

    def _build_pack(self):
        nh, hd = self.num_heads, self.head_dim
        hdp = _pad8(hd)
        w = self.qkv_pw.weight.detach().float().reshape(3 * nh * hd, -1)
        pe = self.pe.weight.detach().float()
        return {"hdp": hdp, "dw": _pack_dw(self.qkv_dw.weight), "qkv": _pack_linear(_pad_heads_out(w, 3, nh, hd, hdp)),
                "pe": _pack_dw(_pad_heads_out(pe, 1, nh, hd, hdp)),
                "proj": _pack_linear(_pad_heads_in(self.proj.weight.detach().float().reshape(-1, nh * hd), nh, hd, hdp)),
                "norm": self._norm_pack()}

    def head_raw(self, x):
        pk = self.get_pack()
        B, H, W, C = x.shape
        nh, hdp = self.num_heads, pk["hdp"]
        inner = nh * hdp
        if self.window_size > 8:
            raise NotImplementedError("MoA local head: window sizes above 8 are not on the B200 path")
        t = ops.dwconv(x, pk["dw"], None, 3, False, C)
        qkv = _lin(t, pk, "qkv")
        v = qkv[..., 2 * inner:]
        vp = ops.dwconv(v, pk["pe"], None, 7, False, inner, add=v)
        win = max(1, min(self.window_size, H, W))
        o = ops.attn_window(qkv[..., :inner], qkv[..., inner:2 * inner], vp, nh, hdp, win, 0, self.scale)
        return _lin(o, pk, "proj")


class _RegionalAttnHead(_Head):
    """`_RegionalAttnHead(dim, num_heads, head_dim=None, pool_stride=2, max_kv_tokens=4096)` (moa/heads.py:162-246)."""

    def __init__(self, dim, num_heads, head_dim=None, pool_stride=2, max_kv_tokens=4096):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = head_dim or max(dim // num_heads, 16)
        inner = self.head_dim * num_heads
        if pool_stride < 1:
            raise ValueError(f"pool_stride must be ≥ 1, got {pool_stride}")
        if max_kv_tokens is not None and max_kv_tokens < 1:
            raise ValueError(f"max_kv_tokens must be positive or None, got {max_kv_tokens}")
        self.pool_stride = pool_stride
        self.max_kv_tokens = None if max_kv_tokens is None else int(max_kv_tokens)
        self.q_proj = nn.Conv2d(dim, inner, 1, bias=False)
        self.kv_proj = nn.Conv2d(dim, inner * 2, 1, bias=False)
        self.proj = nn.Conv2d(inner, dim, 1, bias=False)
        self.norm = nn.GroupNorm(get_safe_groups(dim, 8), dim)
        self.scale = self.head_dim ** -0.5

    def _build_pack(self):
        nh, hd = self.num_heads, self.head_dim
        hdp = _pad8(hd)
        return {"hdp": hdp,
                "q": _pack_linear(_pad_heads_out(self.q_proj.weight.detach().float().reshape(nh * hd, -1), 1, nh, hd, hdp)),
                "kv": _pack_linear(_pad_heads_out(self.kv_proj.weight.detach().float().reshape(2 * nh * hd, -1), 2, nh, hd, hdp)),
                "proj": _pack_linear(_pad_heads_in(self.proj.weight.detach().float().reshape(-1, nh * hd), nh, hd, hdp)),
                "norm": self._norm_pack()}

    def head_raw(self, x):
        pk = self.get_pack()
        B, H, W, C = x.shape
        nh, hdp = self.num_heads, pk["hdp"]
        inner = nh * hdp
        if min(H, W) <= 1:
            src = x
        else:
            s = self.pool_stride
            if self.max_kv_tokens is not None:
                while max(1, H // s) * max(1, W // s) > self.max_kv_tokens:
                    s *= 2
            src = ops.adaptive_avgpool(x, max(1, H // s), max(1, W // s))
        kv = _lin(src, pk, "kv")
        q = _lin(x, pk, "q")
        o = ops.attn_small(q, kv[..., :inner], kv[..., inner:], nh, hdp, self.scale)
        return _lin(o, pk, "proj")


class _GlobalAttnHead(_Head):
    """`_GlobalAttnHead(dim, num_heads, head_dim=None, nb_features=64, rf_seed=0x5F3759DF)` (moa/heads.py:249-380)."""

    _LINEAR_ATTN_THRESHOLD = LINEAR_ATTN_THRESHOLD
    _BLEND_WINDOW = LINEAR_ATTN_BLEND_WINDOW

    def __init__(self, dim, num_heads, head_dim=None, nb_features=64, rf_seed=DEFAULT_RF_SEED):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = head_dim or max(dim // num_heads, 16)
        inner = self.head_dim * num_heads
        self.nb_features = nb_features
        self.qkv = nn.Conv2d(dim, inner * 3, 1, bias=False)
        self.proj = nn.Conv2d(inner, dim, 1, bias=False)
        self.norm = nn.GroupNorm(get_safe_groups(dim, 8), dim)
        self.scale = self.head_dim ** -0.5
        eff_nb = min(self.nb_features, self.head_dim)
        with torch.no_grad():   # persistent buffer: orthogonal random features, seeded per block (moa/heads.py:298-311)
            gen = torch.Generator().manual_seed(rf_seed)
            rf = torch.randn(self.head_dim, self.head_dim, generator=gen, dtype=torch.float32)
            rf, _ = torch.linalg.qr(rf)
        self.register_buffer("_rf_matrix", rf[:eff_nb].contiguous(), persistent=True)
        self._alpha = {}

    def _build_pack(self):
        nh, hd = self.num_heads, self.head_dim
        hdp = _pad8(hd)
        if hdp > 32:
            raise NotImplementedError(f"MoA global head: head_dim {hd} > 32 is not on the B200 path")
        return {"hdp": hdp,
                "qkv": _pack_linear(_pad_heads_out(self.qkv.weight.detach().float().reshape(3 * nh * hd, -1), 3, nh, hd, hdp)),
                "proj": _pack_linear(_pad_heads_in(self.proj.weight.detach().float().reshape(-1, nh * hd), nh, hd, hdp)),
                "rf": _f32(self._rf_matrix), "norm": self._norm_pack()}

    def head_raw(self, x):
        pk = self.get_pack()
        B, H, W, C = x.shape
        N, nh, hdp, hd = H * W, self.num_heads, pk["hdp"], self.head_dim
        inner = nh * hdp
        qkv = _lin(x, pk, "qkv")
        q, k, v = qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]
        thr, blend = self._LINEAR_ATTN_THRESHOLD, self._BLEND_WINDOW
        if N <= thr:
            o = ops.attn_small(q, k, v, nh, hdp, self.scale)
            if N > thr - blend:   # (1 - a) * exact + a * linear, a = (N - 448) / 64
                a = (N - (thr - blend)) / blend
                key = (N, str(x.device))
                if key not in self._alpha:
                    self._alpha[key] = torch.tensor([a], dtype=torch.float32, device=x.device)
                lin = ops.linear_attn(q, k, v, nh, hdp, hd, pk["rf"], 1e-6, LINEAR_ATTN_ACTIVATION_LIMIT)
                o = ops.ew(ops.EW_LERP, a=lin, b=o, p0=self._alpha[key], out=o)
        else:
            o = ops.linear_attn(q, k, v, nh, hdp, hd, pk["rf"], 1e-6, LINEAR_ATTN_ACTIVATION_LIMIT)
        return _lin(o, pk, "proj")


class _MoARouter(nn.Module, PackCache):
    """`_MoARouter(dim, num_groups, reduction=8, temperature=1.0)` (moa/router.py:28-62): soft per-token routing, fp32."""

    def __init__(self, dim, num_groups, reduction=8, temperature=1.0):
        super().__init__()
        self.temperature = max(temperature, 0.1)
        self.num_groups = num_groups
        hidden = max(dim // reduction, num_groups * 2)
        self.router = nn.Sequential(nn.Conv2d(dim, hidden, 1, bias=False), nn.GroupNorm(get_safe_groups(hidden, 4), hidden),
                                    nn.SiLU(inplace=False), nn.Conv2d(hidden, num_groups, 1, bias=True))
        nn.init.zeros_(self.router[-1].weight)
        nn.init.zeros_(self.router[-1].bias)

    def _build_pack(self):
        r = self.router
        return {"w1": _f32(r[0].weight.reshape(r[0].weight.shape[0], -1)), "gn_w": _f32(r[1].weight), "gn_b": _f32(r[1].bias),
                "G": r[1].num_groups, "w2": _f32(r[3].weight.reshape(self.num_groups, -1)), "b2": _f32(r[3].bias)}

    def route(self, x):
        return ops.token_router(x, self.get_pack(), self.num_groups, temp=self.temperature, want_idx=False)[0]

    def forward(self, x, return_logits=False):
        require_eval(self)
        if return_logits:
            raise NotImplementedError("MoA router: logits are not materialised on the B200 path")
        return self.route(to_nhwc(x)).permute(0, 3, 1, 2).to(x.dtype)


class MoABlock(nn.Module, PackCache):
    """`MoABlock(dim, num_heads=8, mlp_ratio=2.0, temperature=1.0, attn_drop=0.0, shortcut=True, aux_loss_coeff=0.01,
    block_index=0, local_window_size=7, sequential_heads=True, regional_max_kv_tokens=4096, sparse_inference=False,
    sparse_inference_threshold=0.02, inference_sparse_threshold=None)` (moa/block.py:27-278)."""

    NUM_GROUPS = 3

    def __init__(self, dim, num_heads=8, mlp_ratio=2.0, temperature=1.0, attn_drop=0.0, shortcut=True, aux_loss_coeff=0.01,
                 block_index=0, local_window_size=7, sequential_heads=True, regional_max_kv_tokens=4096, sparse_inference=False,
                 sparse_inference_threshold=0.02, inference_sparse_threshold=None):
        super().__init__()
        if inference_sparse_threshold is not None:
            sparse_inference = True
        if sparse_inference:
            raise NotImplementedError("MoABlock: sparse_inference (batch-level head skipping, a host decision) is not on the B200 path")
        if num_heads <= 0 or num_heads % self.NUM_GROUPS != 0:
            raise ValueError(f"num_heads ({num_heads}) must be positive and divisible by NUM_GROUPS ({self.NUM_GROUPS})")
        self.sequential_heads, self.sparse_inference = sequential_heads, False
        self.shortcut, self.aux_loss_coeff = shortcut, aux_loss_coeff
        head_dim = max(dim // num_heads, 16)
        hpg = num_heads // self.NUM_GROUPS
        self.local_head = _LocalAttnHead(dim, hpg, head_dim, window_size=local_window_size)
        self.region_head = _RegionalAttnHead(dim, hpg, head_dim, max_kv_tokens=regional_max_kv_tokens)
        self.global_head = _GlobalAttnHead(dim, hpg, head_dim, rf_seed=block_index * 7919 + 2 * 65537)
        self.router = _MoARouter(dim, self.NUM_GROUPS, temperature=temperature)
        self.fusion = Conv(dim, dim, 1, act=False)
        self.attn_drop = nn.Dropout2d(attn_drop) if attn_drop > 0 else nn.Identity()
        ls = torch.ones(dim, 1, 1) * (0.1 if shortcut else 1.0)
        self.ls_attn = nn.Parameter(ls.clone())
        hidden = int(dim * mlp_ratio)
        self.ffn = nn.Sequential(Conv(dim, hidden, 1), Conv(hidden, dim, 1, act=False))
        self.ls_ffn = nn.Parameter(ls.clone())
        self.last_routing_snapshot = {}

    @property
    def num_experts(self):
        return self.NUM_GROUPS

    @property
    def top_k(self):
        return self.NUM_GROUPS

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.ls_attn.device)

    last_aux_loss = aux_loss

    def _pack_sources(self):
        return [self.ls_attn, self.ls_ffn, *self.fusion.parameters(), *self.fusion.buffers(), *self.ffn[1].parameters(),
                *self.ffn[1].buffers()]

    def _build_pack(self):
        wf, bf = fold_bn(self.fusion.conv.weight, None, self.fusion.bn)
        w1, b1 = fold_bn(self.ffn[1].conv.weight, None, self.ffn[1].bn)
        return {"fusion": _pack_linear(wf, bf, self.ls_attn), "ffn1": _pack_linear(w1, b1, self.ls_ffn)}

    def fwd_nhwc(self, x, out=None):
        require_eval(self)
        pk = self.get_pack()
        wts = self.router.route(x)                                           # (B,H,W,3) fp32
        self.last_routing_snapshot = {"num_experts": self.NUM_GROUPS, "top_k": self.NUM_GROUPS, "weights": wts}
        tok = wts.view(-1)
        mixed = None
        for g, head in enumerate((self.local_head, self.region_head, self.global_head)):
            p = head.head_raw(x)
            mixed = ops.groupnorm(p, head.norm.num_groups, *head.get_pack()["norm"], eps=head.norm.eps, tok=tok,
                                  ldt=self.NUM_GROUPS, toff=g, add=mixed, out=mixed)      # mixed += w_g * GN(proj(attn_g))
        x1 = _lin(mixed, pk, "fusion", res=x if self.shortcut else None)      # x + ls_attn * fusion(mixed)
        h = self.ffn[0].fwd_nhwc(x1)
        return _lin(h, pk, "ffn1", res=x1 if self.shortcut else None, out=out)  # x + ls_ffn * ffn(x)

    def forward(self, x):
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))


class C2fMoA(nn.Module):
    """`C2fMoA(c1, c2, n=1, num_heads=6, mlp_ratio=2.0, temperature=1.0, shortcut=True, e=0.5, aux_loss_coeff=0.01,
    local_window_size=7, sequential_heads=True, regional_max_kv_tokens=4096, sparse_inference=False,
    sparse_inference_threshold=0.02, inference_sparse_threshold=None)` (moa/wrappers.py:28-186)."""

    def __init__(self, c1, c2, n=1, num_heads=6, mlp_ratio=2.0, temperature=1.0, shortcut=True, e=0.5, aux_loss_coeff=0.01,
                 local_window_size=7, sequential_heads=True, regional_max_kv_tokens=4096, sparse_inference=False,
                 sparse_inference_threshold=0.02, inference_sparse_threshold=None):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        heads = num_heads
        while heads % MoABlock.NUM_GROUPS != 0:
            heads += 1
        while self.c // heads < 16 and heads > MoABlock.NUM_GROUPS:
            heads -= MoABlock.NUM_GROUPS
        heads = max(heads, MoABlock.NUM_GROUPS)
        self.m = nn.ModuleList(
            MoABlock(self.c, num_heads=heads, mlp_ratio=mlp_ratio, temperature=temperature, shortcut=shortcut,
                     aux_loss_coeff=aux_loss_coeff, block_index=i, local_window_size=local_window_size,
                     sequential_heads=sequential_heads, regional_max_kv_tokens=regional_max_kv_tokens,
                     sparse_inference=sparse_inference or inference_sparse_threshold is not None,
                     sparse_inference_threshold=sparse_inference_threshold)
            for i in range(n))
        self.last_routing_snapshot = {}

    @property
    def num_experts(self):
        return MoABlock.NUM_GROUPS

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.cv1.conv.weight.device)

    def fwd_nhwc(self, x, out=None):
        require_eval(self)
        B, H, W, _ = x.shape
        c, n = self.c, len(self.m)
        cat = ops.new_act(B, H, W, (2 + n) * c, x.device)
        self.cv1.fwd_nhwc(x, out=cat[..., :2 * c])
        for j, m in enumerate(self.m):
            m.fwd_nhwc(cat[..., (j + 1) * c:(j + 2) * c], out=cat[..., (j + 2) * c:(j + 3) * c])
        return self.cv2.fwd_nhwc(cat, out=out)

    def forward(self, x):
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))
