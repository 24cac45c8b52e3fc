"""Mixture-of-Transformers blocks with the reference's names, signatures and state_dict keys
(`ultralytics/nn/modules/mot/{experts,router,block,wrappers}.py`), eval forward only.

Execution plan per `MoTBlock` (all NHWC fp16, no host synchronisation, CUDA-graph capturable):
  * router: `ym_token_router` (1x1 -> GroupNorm -> SiLU -> 1x1 -> softmax/T -> per-token top-k -> renormalise), fp32;
  * the three experts run on every image and are blended with the router's DENSE per-token weights (zero off the top-k).
    This equals the reference's sample-sparse dispatch (mot/block.py:347-364), where an expert is skipped for an image only
    when none of its tokens selected it, i.e. exactly when all its weights are zero - but it keeps shapes static;
  * layer-scales (ls1/ls2) are folded into the preceding 1x1 projection and the residual add rides in that conv's epilogue;
  * window partition / cyclic shift / padding are index arithmetic inside `ym_attn_window` (no permute/roll/pad copies).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ._base import PackCache, fold_bn, pack_gemm_weight, require_eval, to_nchw, to_nhwc
from .conv import Conv
from .moe import get_safe_groups

__all__ = ("C2fMoT", "MoTBlock")


def _pack_linear(weight, bias=None, scale=None):
    """[Co,Ci(,1,1)] weight (+bias) with an optional per-output-channel scale folded in -> (fp16 [Co,Kpad], fp32 bias)."""
    w = weight.detach().float().reshape(weight.shape[0], -1)
    b = torch.zeros(w.shape[0], device=w.device) if bias is None else bias.detach().float()
    if scale is not None:
        s = scale.detach().float().reshape(-1)
        w, b = w * s[:, None], b * s
    return pack_gemm_weight(w.reshape(w.shape[0], w.shape[1], 1, 1)), b.contiguous()


def _pack_dw(weight):
    """Depthwise [C,1,k,k] -> fp16 tap-major [k*k][C]."""
    C, _, k, _ = weight.shape
    return weight.detach().float().reshape(C, k * k).t().contiguous().half()


def _f32(t):
    return t.detach().float().contiguous()


def _lin(x, pk, name, act=False, res=None, out=None, out_f32=False):
    w, b = pk[name]
    return ops.conv2d(x, w, b, w.shape[0], 1, 1, 1, 0, act, out=out, res=res, out_f32=out_f32)


class _Expert(nn.Module, PackCache):
    def forward(self, x):
        require_eval(self)
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))


class _LocalConvTransformerExpert(_Expert):
    """`_LocalConvTransformerExpert(dim, num_heads, mlp_ratio=2.0, dropout=0.0, local_window_size=0)` (mot/experts.py:72-171)."""

    def __init__(self, dim, num_heads, mlp_ratio=2.0, dropout=0.0, local_window_size=0):
        super().__init__()
        if num_heads <= 0 or dim % num_heads != 0:
            raise ValueError(f"dim ({dim}) must be divisible by positive num_heads ({num_heads})")
        if int(local_window_size) < 0:
            raise ValueError("local_window_size must be non-negative (0 disables window attention)")
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.local_window_size = int(local_window_size)
        self.dw_mix = nn.Conv2d(dim, dim, 3, padding=1, groups=dim, bias=False)
        self.qkv = nn.Conv2d(dim, dim * 3, 1, bias=False)
        self.pe = nn.Conv2d(dim, dim, 7, padding=3, groups=dim, bias=False)
        self.proj = nn.Conv2d(dim, dim, 1, bias=False)
        self.norm1 = nn.GroupNorm(get_safe_groups(dim, 8), dim)
        self.norm2 = nn.GroupNorm(get_safe_groups(dim, 8), dim)
        self.drop = nn.Dropout2d(dropout)
        ffn_hidden = int(dim * mlp_ratio)
        self.ffn_gate = nn.Sequential(Conv(dim, ffn_hidden, 1), nn.Sigmoid())
        self.ffn_val = Conv(dim, ffn_hidden, 1)
        self.ffn_out = Conv(ffn_hidden, dim, 1, act=False)
        self.ls1 = nn.Parameter(torch.ones(dim, 1, 1) * 0.1)
        self.ls2 = nn.Parameter(torch.ones(dim, 1, 1) * 0.1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.trunc_normal_(m.weight, std=0.02)

    def _build_pack(self):
        if self.head_dim % 8:
            raise NotImplementedError(f"MoT LocalConv expert: head_dim {self.head_dim} must be a multiple of 8 on the B200 path")
        wg, bg = fold_bn(self.ffn_gate[0].conv.weight, None, self.ffn_gate[0].bn)
        wv, bv = fold_bn(self.ffn_val.conv.weight, None, self.ffn_val.bn)
        wo, bo = fold_bn(self.ffn_out.conv.weight, None, self.ffn_out.bn)
        return {
            "dw_mix": _pack_dw(self.dw_mix.weight), "pe": _pack_dw(self.pe.weight),
            "qkv": _pack_linear(self.qkv.weight), "proj": _pack_linear(self.proj.weight, None, self.ls1),
            "gv": _pack_linear(torch.cat([wg, wv], 0), torch.cat([bg, bv], 0)),      # gate | value in one GEMM
            "out": _pack_linear(wo, bo, self.ls2),
            "n1": (_f32(self.norm1.weight), _f32(self.norm1.bias)), "n2": (_f32(self.norm2.weight), _f32(self.norm2.bias)),
        }

    def fwd_nhwc(self, x, out=None):
        pk = self.get_pack()
        B, H, W, C = x.shape
        nh, hd = self.num_heads, self.head_dim
        xn = ops.groupnorm(x, self.norm1.num_groups, *pk["n1"], eps=self.norm1.eps)
        t = ops.dwconv(xn, pk["dw_mix"], None, 3, False, C)
        qkv = _lin(t, pk, "qkv")
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        vp = ops.dwconv(v, pk["pe"], None, 7, False, C, add=v)                      # v + pe(v)
        win = self.local_window_size
        if win > 0 and H * W > win * win:
            o = ops.attn_window(q, k, vp, nh, hd, win, 0, self.scale)
        else:
            o = ops.attn_small(q, k, vp, nh, hd, self.scale)
        x1 = _lin(o, pk, "proj", res=x)                                             # x + ls1 * proj(attn)
        xn2 = ops.groupnorm(x1, self.norm2.num_groups, *pk["n2"], eps=self.norm2.eps)
        gv = _lin(xn2, pk, "gv", act=True)
        hid = gv.shape[3] // 2
        ffn = ops.ew(ops.EW_GLU, a=gv[..., :hid], b=gv[..., hid:])                  # sigmoid(gate) * value
        return _lin(ffn, pk, "out", res=x1, out=out)                                # x + ls2 * ffn_out(ffn)


class _TokenFFN:
    """Shared `x + ls2 * Linear(GELU(Linear(LayerNorm(x))))` tail of the window / deformable experts."""

    def _ffn_pack(self):
        return {"ffn0": _pack_linear(self.ffn[0].weight, self.ffn[0].bias),
                "ffn3": _pack_linear(self.ffn[3].weight, self.ffn[3].bias, self.ls2),
                "n2": (_f32(self.norm2.weight), _f32(self.norm2.bias))}

    def _ffn(self, x1, pk, out=None):
        xn2 = ops.layernorm(x1, *pk["n2"], eps=self.norm2.eps)
        h = _lin(xn2, pk, "ffn0")
        g = ops.ew(ops.EW_GELU, a=h, out=h)
        return _lin(g, pk, "ffn3", res=x1, out=out)


def _token_ffn(dim, mlp_ratio, dropout):
    hidden = int(dim * mlp_ratio)
    return nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden, dim))


class _WindowTransformerExpert(_Expert, _TokenFFN):
    """`_WindowTransformerExpert(dim, num_heads, window_size=7, mlp_ratio=2.0, dropout=0.0, shift_size=0)` (mot/experts.py:174-315)."""

    def __init__(self, dim, num_heads, window_size=7, mlp_ratio=2.0, dropout=0.0, shift_size=0):
        super().__init__()
        if num_heads <= 0 or dim % num_heads != 0:
            raise ValueError(f"dim ({dim}) must be divisible by positive num_heads ({num_heads})")
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.win = window_size
        self.shift_size = (window_size // 2) if shift_size else 0
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.proj = nn.Linear(dim, dim, bias=False)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.ffn = _token_ffn(dim, mlp_ratio, dropout)
        self.drop = nn.Dropout(dropout)
        self.ls1 = nn.Parameter(torch.ones(dim) * 0.1)
        self.ls2 = nn.Parameter(torch.ones(dim) * 0.1)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def _build_pack(self):
        if self.head_dim % 8 or not 1 <= self.win <= 8:
            raise NotImplementedError(f"MoT Window expert: head_dim {self.head_dim} % 8 and window {self.win} <= 8 required on the B200 path")
        pk = self._ffn_pack()
        pk["qkv"] = _pack_linear(self.qkv.weight)
        pk["proj"] = _pack_linear(self.proj.weight, None, self.ls1)
        pk["n1"] = (_f32(self.norm1.weight), _f32(self.norm1.bias))
        # tokens added by the pad-to-window are zeros BEFORE LayerNorm (mot/experts.py:270-284): LN(0) = beta, so their
        # q/k/v are the constant vector W_qkv . beta (computed on the fp16-rounded operands the kernels use)
        pad = (self.qkv.weight.detach().half().float() @ self.norm1.bias.detach().half().float()).half().contiguous()
        C = self.qkv.weight.shape[1]
        pk["padk"], pk["padv"] = pad[C:2 * C].contiguous(), pad[2 * C:].contiguous()
        return pk

    def fwd_nhwc(self, x, out=None):
        pk = self.get_pack()
        C = x.shape[3]
        xn = ops.layernorm(x, *pk["n1"], eps=self.norm1.eps)
        qkv = _lin(xn, pk, "qkv")
        o = ops.attn_window(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], self.num_heads, self.head_dim, self.win,
                            self.shift_size, self.scale, padk=pk["padk"], padv=pk["padv"])
        x1 = _lin(o, pk, "proj", res=x)
        return self._ffn(x1, pk, out=out)


class _DeformableTransformerExpert(_Expert, _TokenFFN):
    """`_DeformableTransformerExpert(dim, num_heads, n_points=4, mlp_ratio=2.0, dropout=0.0, align_corners=True)`
    (mot/experts.py:318-496)."""

    def __init__(self, dim, num_heads, n_points=4, mlp_ratio=2.0, dropout=0.0, align_corners=True):
        super().__init__()
        if num_heads <= 0 or dim % num_heads != 0:
            raise ValueError(f"dim ({dim}) must be divisible by positive num_heads ({num_heads})")
        self.num_heads, self.head_dim, self.n_points, self.align_corners = num_heads, dim // num_heads, n_points, align_corners
        self.q_proj = nn.Linear(dim, dim, bias=False)
        self.v_proj = nn.Linear(dim, dim, bias=False)
        self.offset_proj = nn.Linear(dim, num_heads * n_points * 2, bias=True)
        self.attn_proj = nn.Linear(dim, num_heads * n_points, bias=True)
        self.out_proj = nn.Linear(dim, dim, bias=False)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.ffn = _token_ffn(dim, mlp_ratio, dropout)
        self.drop = nn.Dropout(dropout)
        self.ls1 = nn.Parameter(torch.ones(dim) * 0.1)
        self.ls2 = nn.Parameter(torch.ones(dim) * 0.1)
        nn.init.zeros_(self.offset_proj.weight)
        nn.init.zeros_(self.offset_proj.bias)
        nn.init.zeros_(self.attn_proj.weight)
        nn.init.zeros_(self.attn_proj.bias)
        for m in [self.q_proj, self.v_proj, self.out_proj, self.ffn[0], self.ffn[3]]:
            nn.init.trunc_normal_(m.weight, std=0.02)

    def _build_pack(self):
        if self.head_dim % 8:
            raise NotImplementedError(f"MoT Deformable expert: head_dim {self.head_dim} must be a multiple of 8 on the B200 path")
        pk = self._ffn_pack()
        pk["qv"] = _pack_linear(torch.cat([self.q_proj.weight, self.v_proj.weight], 0))                 # q | v in one GEMM
        pk["oa"] = _pack_linear(torch.cat([self.offset_proj.weight, self.attn_proj.weight], 0),
                                torch.cat([self.offset_proj.bias, self.attn_proj.bias], 0))            # offsets | point logits
        pk["out"] = _pack_linear(self.out_proj.weight, None, self.ls1)
        pk["n1"] = (_f32(self.norm1.weight), _f32(self.norm1.bias))
        return pk

    def fwd_nhwc(self, x, out=None):
        pk = self.get_pack()
        C = x.shape[3]
        xn = ops.layernorm(x, *pk["n1"], eps=self.norm1.eps)
        qv = _lin(xn, pk, "qv")
        oa = _lin(qv[..., :C], pk, "oa", out_f32=True)
        o = ops.deform_sample(oa, qv[..., C:], self.num_heads, self.head_dim, self.n_points, self.align_corners)
        x1 = _lin(o, pk, "out", res=x)
        return self._ffn(x1, pk, out=out)


class _MoTRouter(nn.Module, PackCache):
    """`_MoTRouter(dim, num_experts=3, top_k=2, use_spatial=True, temperature=1.0, exploration_eps=0.02, ...)`
    (mot/router.py:63-291); spatial (token-level), non scene-aware routing only."""

    def __init__(self, dim, num_experts=3, top_k=2, use_spatial=True, temperature=1.0, exploration_eps=0.02, scene_aware=False,
                 scene_hidden_dim=None, scene_inference_mode="dynamic"):
        super().__init__()
        if num_experts < 1:
            raise ValueError(f"num_experts must be positive, got {num_experts}")
        if not 1 <= top_k <= num_experts:
            raise ValueError(f"top_k must be in [1, {num_experts}], got {top_k}")
        if not use_spatial or scene_aware:
            raise NotImplementedError("MoT router: image-level (GAP) and scene-aware routing are not on the B200 path")
        self.num_experts, self.top_k, self.use_spatial, self.scene_aware = num_experts, top_k, use_spatial, False
        self.exploration_eps = min(max(exploration_eps, 0.0), 0.2)
        self.register_buffer("temperature", torch.tensor(max(temperature, 0.1)), persistent=True)
        hidden = max(dim // 8, num_experts * 4)
        self.router = nn.Sequential(nn.Conv2d(dim, hidden, 1, bias=False), nn.GroupNorm(get_safe_groups(hidden, 4), hidden),
                                    nn.SiLU(inplace=False), nn.Conv2d(hidden, num_experts, 1, bias=True))

    def _build_pack(self):
        r = self.router
        return {"w1": _f32(r[0].weight.reshape(r[0].weight.shape[0], -1)), "gn_w": _f32(r[1].weight), "gn_b": _f32(r[1].bias),
                "G": r[1].num_groups, "w2": _f32(r[3].weight.reshape(self.num_experts, -1)), "b2": _f32(r[3].bias),
                "temp": _f32(self.temperature.reshape(1))}

    def route(self, x):
        """x: (B,H,W,C) -> dense weights fp32 (B,H,W,E), indices int32 (B,H,W,k)."""
        pk = self.get_pack()
        return ops.token_router(x, pk, self.top_k, temp_dev=pk["temp"])

    def forward(self, x, return_logits=False):
        require_eval(self)
        if return_logits:
            raise NotImplementedError("MoT router: logits are not materialised on the B200 path")
        w, idx = self.route(to_nhwc(x))
        return w.permute(0, 3, 1, 2).to(x.dtype), idx.permute(0, 3, 1, 2).long()


class MoTBlock(nn.Module, PackCache):
    """`MoTBlock(dim, num_heads=8, top_k=2, window_size=7, n_points=4, mlp_ratio=2.0, temperature=1.0, ...)`
    (mot/block.py:21-473)."""

    NUM_EXPERTS = 3

    def __init__(self, dim, num_heads=8, top_k=2, window_size=7, n_points=4, mlp_ratio=2.0, temperature=1.0,
                 use_spatial_router=True, balance_loss_coeff=0.01, router_z_loss_coeff=None, dropout=0.0, exploration_eps=0.02,
                 window_shift=False, grid_align_corners=True, sparse_train=False, scene_aware_router=False, scene_hidden_dim=None,
                 scene_consistency_coeff=0.0, sparse_train_warmup_steps=0, scene_inference_mode="dynamic", local_attn_window=0):
        super().__init__()
        if not 1 <= top_k <= self.NUM_EXPERTS:
            raise ValueError(f"top_k must be in [1, {self.NUM_EXPERTS}], got {top_k}")
        if int(sparse_train_warmup_steps) < 0:
            raise ValueError("sparse_train_warmup_steps must be non-negative")
        self._top_k = int(top_k)
        self.balance_loss_coeff, self.sparse_train = balance_loss_coeff, sparse_train
        self.register_buffer("_sparse_train_step", torch.tensor(0, dtype=torch.long), persistent=True)
        heads = num_heads
        while dim % heads != 0 and heads > 1:
            heads -= 1
        heads = max(1, heads)
        self.experts = nn.ModuleList([
            _LocalConvTransformerExpert(dim, heads, mlp_ratio, dropout, local_window_size=local_attn_window),
            _WindowTransformerExpert(dim, heads, window_size, mlp_ratio, dropout, shift_size=window_size // 2 if window_shift else 0),
            _DeformableTransformerExpert(dim, heads, n_points, mlp_ratio, dropout, align_corners=grid_align_corners),
        ])
        self.router = _MoTRouter(dim, self.NUM_EXPERTS, top_k, use_spatial=use_spatial_router, temperature=temperature,
                                 exploration_eps=exploration_eps, scene_aware=scene_aware_router, scene_hidden_dim=scene_hidden_dim,
                                 scene_inference_mode=scene_inference_mode)
        self.out_norm = nn.GroupNorm(get_safe_groups(dim, 8), dim)
        self.out_proj = nn.Conv2d(dim, dim, 1, bias=False)
        self.last_aux_loss = None
        self.last_routing_snapshot = {}

    @property
    def num_experts(self):
        return self.NUM_EXPERTS

    @property
    def top_k(self):
        return self._top_k

    @top_k.setter
    def top_k(self, value):
        value = int(value)
        if not 1 <= value <= self.NUM_EXPERTS:
            raise ValueError(f"top_k must be in [1, {self.NUM_EXPERTS}], got {value}")
        self._top_k = value
        if hasattr(self, "router"):
            self.router.top_k = value

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.out_proj.weight.device)

    def _pack_sources(self):
        return [self.out_proj.weight, self.out_norm.weight, self.out_norm.bias]

    def _build_pack(self):
        return {"out_proj": _pack_linear(self.out_proj.weight), "on": (_f32(self.out_norm.weight), _f32(self.out_norm.bias))}

    def fwd_nhwc(self, x, out=None):
        require_eval(self)
        pk = self.get_pack()
        wts, idx = self.router.route(x)
        self.last_routing_snapshot = {"num_experts": self.NUM_EXPERTS, "top_k": self.top_k, "weights": wts, "indices": idx}
        acc = None
        for e, expert in enumerate(self.experts):
            y = expert.fwd_nhwc(x)
            acc = ops.ew(ops.EW_TOKEN_ACC, a=acc, b=y, tok=wts.view(-1), ldt=self.NUM_EXPERTS, toff=e, out=acc)
        p = _lin(acc, pk, "out_proj")
        return ops.groupnorm(p, self.out_norm.num_groups, *pk["on"], eps=self.out_norm.eps, add=x, out=out)   # out_norm(.) + x

    def forward(self, x):
        y = to_nchw(self.fwd_nhwc(to_nhwc(x)))
        return y, torch.zeros((), device=y.device, dtype=y.dtype)


class C2fMoT(nn.Module):
    """`C2fMoT(c1, c2, n=1, num_heads=6, top_k=2, window_size=7, n_points=4, mlp_ratio=2.0, temperature=1.0,
    balance_loss_coeff=0.01, e=0.5, sparse_train=False, scene_aware_router=False, scene_hidden_dim=None,
    scene_consistency_coeff=0.0, sparse_train_warmup_steps=0, scene_inference_mode='dynamic', local_attn_window=0)`
    (mot/wrappers.py:19-147)."""

    def __init__(self, c1, c2, n=1, num_heads=6, top_k=2, window_size=7, n_points=4, mlp_ratio=2.0, temperature=1.0,
                 balance_loss_coeff=0.01, e=0.5, sparse_train=False, scene_aware_router=False, scene_hidden_dim=None,
                 scene_consistency_coeff=0.0, sparse_train_warmup_steps=0, scene_inference_mode="dynamic", local_attn_window=0):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        dim, heads = self.c, num_heads
        while heads > 1 and (dim % heads != 0 or dim // heads < 8):
            heads -= 1
        heads = max(1, heads)
        self.m = nn.ModuleList(
            MoTBlock(dim=dim, num_heads=heads, top_k=top_k, window_size=window_size, n_points=n_points, mlp_ratio=mlp_ratio,
                     temperature=temperature, balance_loss_coeff=balance_loss_coeff, window_shift=bool(i % 2),
                     sparse_train=sparse_train, scene_aware_router=scene_aware_router, scene_hidden_dim=scene_hidden_dim,
                     scene_consistency_coeff=scene_consistency_coeff, sparse_train_warmup_steps=sparse_train_warmup_steps,
                     scene_inference_mode=scene_inference_mode, local_attn_window=local_attn_window)
            for i in range(n))
        self.last_routing_snapshot = {}

    @property
    def num_experts(self):
        return MoTBlock.NUM_EXPERTS

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
