"""Gated MoE family with the reference's names, constructor signatures and state_dict keys
(`ultralytics/nn/modules/moe/gated.py`) - eval forward of the whole AdaptiveGateMoE line: `AdaptiveGateMoE`
(v0_4 zoo), `FusedAdaptiveGateMoE` (v0_5), `HybridAdaptiveGateMoE` (v0_6), `LowRankHybridAdaptiveGateMoE` (v0_7),
`RefinedLowRankHybridAdaptiveGateMoE` (v0_8), `DetailAwareLowRankHybridAdaptiveGateMoE` (v0_9),
`ContextRefinedLowRankHybridAdaptiveGateMoE` and `VisualEnhancedAdaptiveGateMoE` (v0_10).  The classes differ in the expert
back-end, the channel shuffle and which of the detail / context / refine stages they carry (gated.py:508-555, :1340-1386,
moe/_gated_visual.py:33-86); one implementation serves them all.

Per block: SE gate (pooled vector -> two-layer MLP, `ym_fc_gate`) scales the channels, which split into a static half
(depthwise 3x3 -> 1x1, BatchNorm folded) and a dynamic half (detail gate -> per-IMAGE top-k routing, `ym_gate_router`, fp32 ->
routed experts); the halves are concatenated, channel-shuffled, mixed with pyramid context, refined, projected and added to the
input.  Routing never touches the host: expert indices and weights stay in device tables.

Expert back-ends, chosen like the reference (gated.py:1318-1331,1498-1507):
  * E <= fused_expert_threshold: `LowRankFusedExpertGroup` - shared 1x1 bottleneck, then ONE 3x3 conv produces every expert (the
    reference's grouped conv, run here as a dense conv with block-diagonal weights) and `ym_gated_select` normalises and sums the
    routed experts' channel slices;
  * E >  fused_expert_threshold: `SharedInvertedExpertGroup` - shared expand / depthwise features, then the routed experts'
    1x1 + GroupNorm projections through the expert-indexed grouped GEMM of the MoE-FFN path (`ym_moe_expert_gemm`).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops
from ._base import BN_EPS, PackCache, fold_bn, pack_gemm_weight, require_eval, to_nchw, to_nhwc
from .moe import get_safe_groups
from .mot import _f32, _pack_dw, _pack_linear

__all__ = ("DualStreamGateRouter", "FusedExpertGroup", "LowRankFusedExpertGroup", "SharedInvertedExpertGroup", "VisualDetailGate",
           "PyramidContextMixer", "AdaptiveGateMoE", "FusedAdaptiveGateMoE", "HybridAdaptiveGateMoE", "LowRankHybridAdaptiveGateMoE",
           "RefinedLowRankHybridAdaptiveGateMoE", "DetailAwareLowRankHybridAdaptiveGateMoE",
           "ContextRefinedLowRankHybridAdaptiveGateMoE", "VisualEnhancedAdaptiveGateMoE", "ZeroCostRouter", "UltimateOptimizedMoE", "DualStreamGateRouterV2", "HybridAdaptiveGateMoEv2", "OptimalHybridGateMoE", "MultiHeadRouterV3", "CrossPathGate",
           "MultiHeadRouterMoE", "GatedFusionMoE", "SharedExpertMoE")


def _gn(channels: int, groups: int = 8) -> nn.GroupNorm:
    return nn.GroupNorm(get_safe_groups(channels, groups), channels)


class DualStreamGateRouter(nn.Module):
    """`DualStreamGateRouter(in_channels, num_experts, top_k, temperature=1.0, local_reduction=16, pool_scale=4)` (gated.py:82-165)."""

    def __init__(self, in_channels, num_experts, top_k, temperature=1.0, local_reduction=16, pool_scale=4):
        super().__init__()
        self.num_experts = num_experts
        self.top_k = top_k
        self.temperature = max(float(temperature), 1e-3)
        self.pool_scale = pool_scale
        self.global_fc = nn.Linear(2 * in_channels, num_experts, bias=False)
        reduced = max(in_channels // local_reduction, 4)
        self.local_conv = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, 3, padding=1, groups=in_channels, bias=False),
            _gn(in_channels, 8),
            nn.SiLU(inplace=False),
            nn.Conv2d(in_channels, reduced, 1, bias=False),
            _gn(reduced, 4),
            nn.SiLU(inplace=False),
            nn.Conv2d(reduced, num_experts, 1, bias=True),
        )
        self.alpha = nn.Parameter(torch.tensor(0.5))

    def pack(self):
        lc = self.local_conv
        C, R, E = lc[0].weight.shape[0], lc[3].weight.shape[0], self.num_experts
        return {
            "E": E, "R": R, "pool": int(self.pool_scale), "G1": lc[1].num_groups, "G2": lc[4].num_groups, "eps": float(lc[1].eps),
            "global_fc": _f32(self.global_fc.weight), "dw": _f32(lc[0].weight).reshape(C, 9).contiguous(),
            "gn1_w": _f32(lc[1].weight), "gn1_b": _f32(lc[1].bias), "pw1": _f32(lc[3].weight).reshape(R, C).contiguous(),
            "gn2_w": _f32(lc[4].weight), "gn2_b": _f32(lc[4].bias), "pw2": _f32(lc[6].weight).reshape(E, R).contiguous(),
            "b2": _f32(lc[6].bias), "alpha": float(torch.sigmoid(self.alpha.detach().float())), "temperature": float(self.temperature),
        }


class DualStreamGateRouterV2(DualStreamGateRouter):
    """`DualStreamGateRouterV2(in_channels, num_experts, top_k, temperature=1.0, local_reduction=16, pool_scale=4, noise_std=0.1)`
    (gated.py:181-260, v0_11 / v0_12 zoos): LayerNorm on the channel statistics + a learnable expert prior on the logits."""

    def __init__(self, in_channels, num_experts, top_k, temperature=1.0, local_reduction=16, pool_scale=4, noise_std=0.1):
        super().__init__(in_channels, num_experts, top_k, temperature, local_reduction, pool_scale)
        self.stat_norm = nn.LayerNorm(2 * in_channels)
        self.expert_prior = nn.Parameter(torch.zeros(num_experts))

    def pack(self):
        pk = super().pack()
        pk["stat_norm"] = (_f32(self.stat_norm.weight), _f32(self.stat_norm.bias), float(self.stat_norm.eps))
        pk["prior"] = _f32(self.expert_prior)
        return pk


class MultiHeadRouterV3(nn.Module):
    """`MultiHeadRouterV3(in_channels, num_experts, top_k, temperature=1.0, num_heads=4, local_reduction=16, pool_scale=4,
    noise_std=0.1, expert_dropout=0.1)` (gated.py:2026-2211, v0_13 zoo).  Its global branch is linear in the normalised statistics
    (a dense projection blended with per-head projections of consecutive chunks), so it packs into ONE effective [E, 2C] matrix
    and runs on the same kernels as `DualStreamGateRouterV2`."""

    def __init__(self, in_channels, num_experts, top_k, temperature=1.0, num_heads=4, local_reduction=16, pool_scale=4, noise_std=0.1,
                 expert_dropout=0.1):
        super().__init__()
        self.num_experts, self.top_k = num_experts, top_k
        self.temperature = max(float(temperature), 1e-3)
        self.pool_scale = pool_scale
        self.num_heads = max(1, min(num_heads, num_experts))
        stat_dim = 2 * in_channels
        self.stat_norm = nn.LayerNorm(stat_dim)
        self._head_dim = max(stat_dim // self.num_heads, 4)
        self.heads = nn.ModuleList([nn.Linear(self._head_dim, num_experts, bias=False) for _ in range(self.num_heads)])
        self.global_proj = nn.Linear(stat_dim, num_experts, bias=False)
        self.head_alpha = nn.Parameter(torch.ones(self.num_heads) / self.num_heads)
        self.global_weight = nn.Parameter(torch.tensor(0.1))
        self.expert_prior = nn.Parameter(torch.zeros(num_experts))
        reduced = max(in_channels // local_reduction, 4)
        self.local_conv = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, 3, padding=1, groups=in_channels, bias=False), _gn(in_channels, 8), nn.SiLU(inplace=False),
            nn.Conv2d(in_channels, reduced, 1, bias=False), _gn(reduced, 4), nn.SiLU(inplace=False),
            nn.Conv2d(reduced, num_experts, 1, bias=True))
        self.alpha = nn.Parameter(torch.tensor(0.5))

    def effective_global_weight(self) -> torch.Tensor:
        """[E, 2C]: gw * global_proj + (1 - gw) * hw_i * heads[i] on the columns of chunk i (gated.py:2123-2140)."""
        gw = torch.sigmoid(self.global_weight.detach().float())
        hw = torch.sigmoid(self.head_alpha.detach().float())
        hw = hw / (hw.sum() + 1e-6)
        w = gw * self.global_proj.weight.detach().float()
        sd, hd = w.shape[1], self._head_dim
        for i, h in enumerate(self.heads):
            lo, hi = i * hd, min((i + 1) * hd, sd)          # columns past 2C are zero padding in the reference
            if lo < hi:
                w[:, lo:hi] += (1 - gw) * hw[i] * h.weight.detach().float()[:, :hi - lo]
        return w.contiguous()

    def pack(self):
        lc = self.local_conv
        C, R, E = lc[0].weight.shape[0], lc[3].weight.shape[0], self.num_experts
        return {
            "E": E, "R": R, "pool": int(self.pool_scale), "G1": lc[1].num_groups, "G2": lc[4].num_groups, "eps": float(lc[1].eps),
            "global_fc": self.effective_global_weight(), "dw": _f32(lc[0].weight).reshape(C, 9).contiguous(),
            "gn1_w": _f32(lc[1].weight), "gn1_b": _f32(lc[1].bias), "pw1": _f32(lc[3].weight).reshape(R, C).contiguous(),
            "gn2_w": _f32(lc[4].weight), "gn2_b": _f32(lc[4].bias), "pw2": _f32(lc[6].weight).reshape(E, R).contiguous(),
            "b2": _f32(lc[6].bias), "alpha": float(torch.sigmoid(self.alpha.detach().float())), "temperature": float(self.temperature),
            "stat_norm": (_f32(self.stat_norm.weight), _f32(self.stat_norm.bias), float(self.stat_norm.eps)),
            "prior": _f32(self.expert_prior),
        }


class CrossPathGate(nn.Module):
    """`CrossPathGate(static_channels, dynamic_channels, out_channels, num_groups=8, drop_prob=0.1)` (gated.py:2347-2412, v0_15)."""

    def __init__(self, static_channels, dynamic_channels, out_channels, num_groups=8, drop_prob=0.1):
        super().__init__()
        self.static_channels, self.dynamic_channels, self.out_channels = static_channels, dynamic_channels, out_channels
        stat_dim = static_channels + dynamic_channels
        hidden = max(stat_dim // 4, 8)
        self.gate_net = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(stat_dim, hidden, bias=False),
                                      nn.SiLU(inplace=False), nn.Linear(hidden, out_channels * 2, bias=True))
        self.gate_scale = nn.Parameter(torch.tensor(0.0))
        self.drop_scale = nn.Parameter(torch.tensor(1.0))


class FusedExpertGroup(nn.Module):
    """`FusedExpertGroup(in_channels, out_channels, num_experts, num_groups=8, top_k=2)` (gated.py:1003-1081)."""

    def __init__(self, in_channels, out_channels, num_experts, num_groups=8, top_k=2):
        super().__init__()
        self.in_channels, self.out_channels, self.num_experts = in_channels, out_channels, num_experts
        self.top_k = min(int(top_k), num_experts)
        fused_out = out_channels * num_experts
        conv_groups = min(get_safe_groups(in_channels, num_groups), fused_out)
        while conv_groups > 1 and (in_channels % conv_groups != 0 or fused_out % conv_groups != 0):
            conv_groups -= 1
        self.num_groups = max(1, conv_groups)
        self.fused_conv = nn.Conv2d(in_channels, fused_out, 3, padding=1, groups=self.num_groups, bias=False)
        self.norm_groups = get_safe_groups(out_channels, num_groups)
        self.expert_norm_weight = nn.Parameter(torch.ones(num_experts, out_channels))
        self.expert_norm_bias = nn.Parameter(torch.zeros(num_experts, out_channels))

    def dense_weight(self) -> torch.Tensor:
        """The grouped conv as a dense [E*oc, Cin, 3, 3] weight: output o of group g reads input slice g, zeros elsewhere."""
        w = self.fused_conv.weight.detach().float()
        Co, cpg = w.shape[0], w.shape[1]
        g = self.num_groups
        dense = torch.zeros((Co, self.in_channels, 3, 3), dtype=torch.float32, device=w.device)
        opg = Co // g
        for k in range(g):
            dense[k * opg:(k + 1) * opg, k * cpg:(k + 1) * cpg] = w[k * opg:(k + 1) * opg]
        return dense


class LowRankFusedExpertGroup(nn.Module):
    """`LowRankFusedExpertGroup(in_channels, out_channels, num_experts, num_groups=8, top_k=2, bottleneck_ratio=0.5,
    min_channels=16)` (gated.py:1101-1147)."""

    def __init__(self, in_channels, out_channels, num_experts, num_groups=8, top_k=2, bottleneck_ratio=0.5, min_channels=16):
        super().__init__()
        self.in_channels, self.out_channels, self.num_experts = in_channels, out_channels, num_experts
        self.top_k = min(int(top_k), num_experts)
        self.bottleneck_channels = min(in_channels, max(min_channels, int(round(in_channels * bottleneck_ratio))))
        self.bottleneck = nn.Sequential(
            nn.Conv2d(in_channels, self.bottleneck_channels, 1, bias=False),
            _gn(self.bottleneck_channels, num_groups),
            nn.SiLU(inplace=False),
        )
        self.fused = FusedExpertGroup(self.bottleneck_channels, out_channels, num_experts, num_groups, top_k=top_k)


class SharedInvertedExpertGroup(nn.Module):
    """`SharedInvertedExpertGroup(in_channels, out_channels, num_experts, expand_ratio=2.0, kernel_size=3, top_k=2,
    weight_threshold=0.0)` (moe/experts.py:176-269)."""

    def __init__(self, in_channels, out_channels, num_experts, expand_ratio=2.0, kernel_size=3, top_k=2, weight_threshold=0.0):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError("SharedInvertedExpertGroup: only kernel_size=3 is on the B200 path")
        self.in_channels, self.out_channels, self.num_experts = in_channels, out_channels, num_experts
        self.top_k, self.weight_threshold = top_k, weight_threshold
        hidden = max(1, int(in_channels * expand_ratio))
        self.shared_feature = nn.Sequential(
            nn.Conv2d(in_channels, hidden, 1, bias=False), _gn(hidden), nn.SiLU(inplace=True),
            nn.Conv2d(hidden, hidden, kernel_size, padding=kernel_size // 2, groups=hidden, bias=False), _gn(hidden), nn.SiLU(inplace=True),
        )
        self.expert_projections = nn.ModuleList(
            nn.Sequential(nn.Conv2d(hidden, out_channels, 1, bias=False), _gn(out_channels)) for _ in range(num_experts))


class DiversifiedExpertGroup(nn.Module):
    """`DiversifiedExpertGroup(in_channels, out_channels, num_experts, expand_ratio=2.0, top_k=2, weight_threshold=0.0, num_groups=8)`
    (gated.py:2214-2333, v0_14): shared expand, then per expert a depthwise 3x3 with dilation 1 + e // 2 and a 1x1 projection.  The
    `dw_dilations` parameters exist for state-dict parity; like the reference's forward, nothing reads them."""

    def __init__(self, in_channels, out_channels, num_experts, expand_ratio=2.0, top_k=2, weight_threshold=0.0, num_groups=8):
        super().__init__()
        self.in_channels, self.out_channels, self.num_experts = in_channels, out_channels, num_experts
        self.top_k, self.weight_threshold = top_k, weight_threshold
        hidden = max(1, int(in_channels * expand_ratio))
        self.shared_expand = nn.Sequential(nn.Conv2d(in_channels, hidden, 1, bias=False), _gn(hidden, num_groups), nn.SiLU(inplace=False))
        self.dw_layers = nn.ModuleList()
        self.dw_dilations = nn.ParameterList()
        for i in range(num_experts):
            d = 1 + (i // 2)
            self.dw_layers.append(nn.Sequential(nn.Conv2d(hidden, hidden, 3, padding=d, dilation=d, groups=hidden, bias=False),
                                                _gn(hidden, num_groups), nn.SiLU(inplace=False)))
            self.dw_dilations.append(nn.Parameter(torch.tensor(float(d))))
        self.expert_projections = nn.ModuleList(
            nn.Sequential(nn.Conv2d(hidden, out_channels, 1, bias=False), _gn(out_channels, num_groups)) for _ in range(num_experts))


class VisualDetailGate(nn.Module):
    """`VisualDetailGate(channels, num_groups=8, reduction=8)` (gated.py:1154-1178)."""

    def __init__(self, channels, num_groups=8, reduction=8):
        super().__init__()
        hidden = max(channels // reduction, 8)
        self.detail_filter = nn.Sequential(
            nn.Conv2d(channels, channels, 3, padding=1, groups=channels, bias=False), _gn(channels, num_groups), nn.SiLU(inplace=False),
            nn.Conv2d(channels, hidden, 1, bias=False), nn.SiLU(inplace=False),
            nn.Conv2d(hidden, channels, 1, bias=True), nn.Sigmoid(),
        )
        self.detail_scale = nn.Parameter(torch.tensor(0.1))


class PyramidContextMixer(nn.Module):
    """`PyramidContextMixer(channels, num_groups=8, pool_scales=(2, 4))` (gated.py:1184-1221)."""

    def __init__(self, channels, num_groups=8, pool_scales=(2, 4)):
        super().__init__()
        self.pool_scales = tuple(pool_scales)
        if len(self.pool_scales) != 2:
            raise NotImplementedError("PyramidContextMixer: exactly two pool scales are on the B200 path")
        self.local_context = nn.Sequential(
            nn.Conv2d(channels, channels, 3, padding=1, groups=channels, bias=False), _gn(channels, num_groups), nn.SiLU(inplace=False))
        self.pool_projections = nn.ModuleList(
            nn.Sequential(nn.Conv2d(channels, channels, 1, bias=False), _gn(channels, num_groups), nn.SiLU(inplace=False))
            for _ in self.pool_scales)
        self.context_gate = nn.Sequential(nn.Conv2d(channels, channels, 1, bias=True), nn.Sigmoid())
        self.context_scale = nn.Parameter(torch.tensor(0.1))


def _gn_args(gn: nn.GroupNorm):
    return gn.num_groups, _f32(gn.weight), _f32(gn.bias), float(gn.eps)


def _norm(x, args, act=False, add=None, out=None):
    G, w, b, eps = args
    return ops.groupnorm(x, G, w, b, eps=eps, act=act, add=add, out=out)


class _GatedMoE(nn.Module, PackCache):
    """Shared body of the AdaptiveGateMoE line.  `backend` in {"shared_inverted", "fused", "low_rank_fused"}; `hooks` is the ordered
    subset of ("detail", "context", "refine") (detail before routing, the others after the concatenation, in this order)."""

    def __init__(self, in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                 backend, shuffle_groups=1, bottleneck_ratio=0.5, hooks=(), refine_reduction=8, detail_reduction=8,
                 fused_expert_threshold=8, router_v2=False, router=None, cross_gate=False):
        super().__init__()
        if in_channels != out_channels:
            raise ValueError(f"{type(self).__name__}: the residual `proj(...) + x` needs in_channels == out_channels")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_experts, self.top_k, self.num_groups = num_experts, top_k, num_groups
        self.initial_temperature, self.final_temperature = initial_temperature, final_temperature
        self.dynamic_channels = int(in_channels * split_ratio)
        self.static_channels = in_channels - self.dynamic_channels
        self.out_dynamic = int(out_channels * split_ratio)
        self.out_static = out_channels - self.out_dynamic
        for n in (self.dynamic_channels, self.static_channels, self.out_dynamic, self.out_static):
            if n % 8:
                raise NotImplementedError(f"{type(self).__name__}: channel halves must be multiples of 8 on the B200 path")
        se_hidden = max(in_channels // 4, 4)
        self.se_gate = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(in_channels, se_hidden, bias=False),
                                     nn.SiLU(inplace=False), nn.Linear(se_hidden, in_channels, bias=True), nn.Sigmoid())
        sc = self.static_channels
        self.static_net = nn.Sequential(
            nn.Conv2d(sc, sc, 3, padding=1, groups=sc, bias=False), nn.BatchNorm2d(sc, eps=BN_EPS, momentum=0.03), nn.SiLU(inplace=False),
            nn.Conv2d(sc, self.out_static, 1, bias=False), nn.BatchNorm2d(self.out_static, eps=BN_EPS, momentum=0.03), nn.SiLU(inplace=False))
        self.routing = router if router is not None else (DualStreamGateRouterV2 if router_v2 else DualStreamGateRouter)(
            self.dynamic_channels, num_experts, top_k, temperature=initial_temperature)
        self.fused_expert_threshold = fused_expert_threshold
        self.shuffle_groups = shuffle_groups if (shuffle_groups and out_channels % shuffle_groups == 0) else 1
        self.expert_backend = backend
        if backend == "low_rank_fused":
            self.fused_experts = LowRankFusedExpertGroup(self.dynamic_channels, self.out_dynamic, num_experts, num_groups,
                                                         top_k=top_k, bottleneck_ratio=bottleneck_ratio)
        elif backend == "fused":
            self.fused_experts = FusedExpertGroup(self.dynamic_channels, self.out_dynamic, num_experts, num_groups, top_k=top_k)
        elif backend == "shared_inverted":
            self.fused_experts = SharedInvertedExpertGroup(self.dynamic_channels, self.out_dynamic, num_experts, top_k=top_k,
                                                           weight_threshold=0.0)
        else:
            raise ValueError(f"unknown expert back-end {backend!r}")
        self.complexity_estimator = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(self.dynamic_channels, 1, 1), nn.Sigmoid())
        self.proj = nn.Conv2d(out_channels, out_channels, 1, bias=False)
        self.bn = _gn(out_channels, num_groups)
        if cross_gate:
            self.cross_gate = CrossPathGate(self.out_static, self.out_dynamic, out_channels, num_groups)
        self.router_hook_names = tuple(hooks)
        oc = out_channels
        if "refine" in hooks:
            hidden = max(oc // refine_reduction, 8)
            self.feature_refiner = nn.Sequential(nn.Conv2d(oc, oc, 3, padding=1, groups=oc, bias=False), _gn(oc, num_groups),
                                                 nn.SiLU(inplace=False))
            self.feature_gate = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(oc, hidden, 1, bias=False), nn.SiLU(inplace=False),
                                              nn.Conv2d(hidden, oc, 1, bias=True), nn.Sigmoid())
            self.refine_scale = nn.Parameter(torch.tensor(0.1))
        if "light_refine" in hooks:   # OptimalHybridGateMoE gated.py:1927-1943: depthwise + GroupNorm (no activation), SE gate
            hidden = max(oc // refine_reduction, 8)
            self.refine_dw = nn.Sequential(nn.Conv2d(oc, oc, 3, padding=1, groups=oc, bias=False), _gn(oc, num_groups))
            self.refine_gate = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(oc, hidden, 1, bias=False), nn.SiLU(inplace=False),
                                             nn.Conv2d(hidden, oc, 1, bias=True), nn.Sigmoid())
            self.refine_scale = nn.Parameter(torch.tensor(0.1))
        if "context" in hooks:
            self.context_mixer = PyramidContextMixer(oc, num_groups)
        if "detail" in hooks:
            self.detail_gate = VisualDetailGate(self.dynamic_channels, num_groups, detail_reduction)
        self.last_routing_snapshot: dict = {}

    # ------------------------------------------------------------------------------------------------------------ weights
    def _build_pack(self):
        dev = self.proj.weight.device
        C, dyn = self.out_channels, self.dynamic_channels
        pk = {}
        pk["se_w1"], pk["se_w2"], pk["se_b2"] = _f32(self.se_gate[2].weight), _f32(self.se_gate[4].weight), _f32(self.se_gate[4].bias)
        if "detail" in self.router_hook_names:
            # x - avg_pool3(x) as ONE depthwise 3x3 (centre 8/9, neighbours -1/9; avg_pool2d counts the padding, so borders agree)
            hp = torch.full((9, dyn), -1.0 / 9.0, dtype=torch.float32, device=dev)
            hp[4] = 8.0 / 9.0
            pk["detail_hp"] = hp.half().contiguous()
            df = self.detail_gate.detail_filter
            pk["df0"], pk["df1"] = _pack_dw(df[0].weight), _gn_args(df[1])
            pk["df3"], pk["df5"] = _pack_linear(df[3].weight), _pack_linear(df[5].weight, df[5].bias)
            pk["detail_t"] = torch.tanh(self.detail_gate.detail_scale.detach().float()).reshape(1).contiguous()
        w, b = fold_bn(self.static_net[0].weight, None, self.static_net[1])
        pk["st_dw"], pk["st_dw_b"] = _pack_dw(w), b.contiguous()
        w, b = fold_bn(self.static_net[3].weight, None, self.static_net[4])
        pk["st_pw"] = (pack_gemm_weight(w), b.contiguous())
        r = self.routing.pack()
        ce = self.complexity_estimator[1]
        r["cx_w"], r["cx_b"] = _f32(ce.weight).reshape(-1).contiguous(), float(ce.bias.detach().float())
        pk["router"] = r
        fe = self.fused_experts
        if isinstance(fe, DiversifiedExpertGroup):
            se = fe.shared_expand
            hid = se[0].weight.shape[0]
            pk["sf0"], pk["sf1"] = _pack_linear(se[0].weight), _gn_args(se[1])
            pk["div_dw"] = torch.stack([_pack_dw(l[0].weight) for l in fe.dw_layers]).contiguous()             # fp16 [E][9][hid]
            pk["div_dil"] = torch.tensor([int(l[0].dilation[0]) for l in fe.dw_layers], dtype=torch.int32, device=dev)
            pk["div_gamma"] = torch.stack([_f32(l[1].weight) for l in fe.dw_layers]).contiguous()
            pk["div_beta"] = torch.stack([_f32(l[1].bias) for l in fe.dw_layers]).contiguous()
            pk["div_G"], pk["div_eps"] = fe.dw_layers[0][1].num_groups, float(fe.dw_layers[0][1].eps)
            pk["div_one"] = torch.ones(hid, dtype=torch.float32, device=dev)
            pk["div_zero"] = torch.zeros(hid, dtype=torch.float32, device=dev)
        elif self.expert_backend == "low_rank_fused":
            pk["bn0"], pk["bn1"] = _pack_linear(fe.bottleneck[0].weight), _gn_args(fe.bottleneck[1])
        if isinstance(fe, DiversifiedExpertGroup):
            pass
        elif self.expert_backend in ("low_rank_fused", "fused"):
            fused = fe.fused if self.expert_backend == "low_rank_fused" else fe
            pk["fused_w"] = pack_gemm_weight(fused.dense_weight())
            pk["fused_gamma"], pk["fused_beta"] = _f32(fused.expert_norm_weight), _f32(fused.expert_norm_bias)
        else:
            sf = fe.shared_feature
            pk["sf0"], pk["sf1"], pk["sf3"], pk["sf4"] = _pack_linear(sf[0].weight), _gn_args(sf[1]), _pack_dw(sf[3].weight), _gn_args(sf[4])
        if hasattr(fe, "expert_projections"):
            pk["proj_w"] = torch.stack([pack_gemm_weight(p[0].weight.detach().float()) for p in fe.expert_projections]).contiguous()
            pk["proj_gamma"] = torch.stack([_f32(p[1].weight) for p in fe.expert_projections]).contiguous()
            pk["proj_beta"] = torch.stack([_f32(p[1].bias) for p in fe.expert_projections]).contiguous()
            pk["proj_G"], pk["proj_eps"] = fe.expert_projections[0][1].num_groups, float(fe.expert_projections[0][1].eps)
            oc_ = pk["proj_gamma"].shape[1]
            pk["proj_one"] = torch.ones(oc_, dtype=torch.float32, device=pk["proj_gamma"].device)
            pk["proj_zero"] = torch.zeros(oc_, dtype=torch.float32, device=pk["proj_gamma"].device)
        if hasattr(self, "cross_gate"):   # only the first C of the gate MLP's 2C outputs are used (gated.py:2404-2408)
            cg = self.cross_gate
            pk["xg_w1"], pk["xg_w2"], pk["xg_b2"] = _f32(cg.gate_net[2].weight), _f32(cg.gate_net[4].weight[:C]), _f32(cg.gate_net[4].bias[:C])
            pk["xg_scale"] = 0.5 * float(torch.tanh(cg.gate_scale.detach().float()))
        sg = self.shuffle_groups
        if sg > 1:   # channel shuffle as a permutation GEMM: new channel i*sg + g <- old channel g*(C/sg) + i  (gated.py:1334-1338)
            perm = torch.zeros((C, C), dtype=torch.float32, device=dev)
            old = torch.arange(C, device=dev)
            perm[(old % (C // sg)) * sg + old // (C // sg), old] = 1.0
            pk["shuffle"] = pack_gemm_weight(perm.reshape(C, C, 1, 1))
        if "context" in self.router_hook_names:
            cm = self.context_mixer
            pk["lc0"], pk["lc1"] = _pack_dw(cm.local_context[0].weight), _gn_args(cm.local_context[1])
            pk["pp"] = [(_pack_linear(p[0].weight), _gn_args(p[1])) for p in cm.pool_projections]
            pk["cg"] = _pack_linear(cm.context_gate[0].weight, cm.context_gate[0].bias)
            pk["ctx_t"] = torch.tanh(cm.context_scale.detach().float()).reshape(1).expand(C).contiguous()
        if "refine" in self.router_hook_names or "light_refine" in self.router_hook_names:
            light = "light_refine" in self.router_hook_names
            dw, fg = (self.refine_dw, self.refine_gate) if light else (self.feature_refiner, self.feature_gate)
            pk["fr0"], pk["fr1"] = _pack_dw(dw[0].weight), _gn_args(dw[1])
            pk["fg_w1"] = _f32(fg[1].weight).reshape(fg[1].weight.shape[0], C).contiguous()
            pk["fg_w2"], pk["fg_b2"] = _f32(fg[3].weight).reshape(C, -1).contiguous(), _f32(fg[3].bias)
            pk["refine_t"] = float(torch.tanh(self.refine_scale.detach().float()))
        pk["proj"], pk["bn"] = _pack_linear(self.proj.weight), _gn_args(self.bn)
        return pk

    # ------------------------------------------------------------------------------------------------------------ forward
    @staticmethod
    def _routed_proj(feat, pk, rj, wj, B, H, W, hid, oc, G):
        """expert_projections[e_b]: grouped 1x1 GEMM of image b with its routed expert + that expert's GroupNorm affine times the
        routing weight, as (o, scale, shift).  GroupNorm statistics ride in the GEMM epilogue when a group is a whole number of its
        8-channel granules; narrower groups (n-scale: 32 or 24 output channels in 8 groups) take a statistics pass over o."""
        HW = H * W
        wide = (oc // G) % 8 == 0
        o, st = ops.moe_expert_gemm(feat, ops.pitch(feat), 1, B, HW, hid, pk["proj_w"], rj, oc, groups=G if wide else 0)
        if wide:
            sc, sh = ops.gn_finalize(st, B, HW, G, oc, HW * (oc // G), pk["proj_eps"], pk["proj_gamma"], pk["proj_beta"], rj, route_w=wj)
        else:
            sc, sh = ops.groupnorm_stats(o.view(B, H, W, oc), G, pk["proj_one"], pk["proj_zero"], pk["proj_eps"])
            ops.route_affine(sc, sh, pk["proj_gamma"], pk["proj_beta"], rj, route_w=wj)
        return o, sc, sh

    def _experts(self, xd, idx, w, pk, out):
        B, H, W, _ = xd.shape
        fe = self.fused_experts
        if isinstance(fe, DiversifiedExpertGroup):
            # shared expand once; per routing rank the ROUTED depthwise (taps and dilation of expert idx[b, j], chosen on the device),
            # its per-expert GroupNorm + SiLU, the grouped projection GEMM and the weighted accumulation (gated.py:2297-2333)
            hid = pk["sf0"][0].shape[0]
            h = _norm(ops.conv2d(xd, *pk["sf0"], hid, 1, 1, 1, 0, False), pk["sf1"], act=True)
            oc, G, HW = self.out_dynamic, pk["proj_G"], H * W
            acc = None
            for j in range(idx.shape[1]):
                rj, wj = idx[:, j].contiguous(), w[:, j].contiguous()
                d = ops.dwconv3_routed(h, pk["div_dw"], rj, pk["div_dil"])
                sc, sh = ops.groupnorm_stats(d, pk["div_G"], pk["div_one"], pk["div_zero"], pk["div_eps"])
                ops.route_affine(sc, sh, pk["div_gamma"], pk["div_beta"], rj)
                feat = ops.ew(ops.EW_AFFINE, a=d, p0=sc, p1=sh, rows_per_img=HW, act=True)
                o, sc2, sh2 = self._routed_proj(feat, pk, rj, wj, B, H, W, hid, oc, G)
                last = j == idx.shape[1] - 1
                acc = ops.ew(ops.EW_AFFINE, a=o.view(B, H, W, oc), b=acc, p0=sc2, p1=sh2, rows_per_img=HW, act=False, out=out if last else None)
            return acc
        if self.expert_backend in ("low_rank_fused", "fused"):
            t, fused = xd, fe
            if self.expert_backend == "low_rank_fused":
                t, fused = _norm(ops.conv2d(xd, *pk["bn0"], pk["bn0"][0].shape[0], 1, 1, 1, 0, False), pk["bn1"], act=True), fe.fused
            E, oc = self.num_experts, self.out_dynamic
            fo = ops.conv2d(t, pk["fused_w"], None, E * oc, 3, 3, 1, 1, False)            # every expert, one dense conv
            return ops.gated_select(fo, idx, w, pk["fused_gamma"], pk["fused_beta"], E, oc, fused.norm_groups, 1e-5, out=out)
        hid = pk["sf0"][0].shape[0]
        h = _norm(ops.conv2d(xd, *pk["sf0"], hid, 1, 1, 1, 0, False), pk["sf1"], act=True)
        feat = _norm(ops.dwconv(h, pk["sf3"], None, 3, False, hid), pk["sf4"], act=True)
        oc, G, HW = self.out_dynamic, pk["proj_G"], H * W
        acc = None
        for j in range(idx.shape[1]):                       # one grouped GEMM per routing rank; a zero weight contributes zero
            rj, wj = idx[:, j].contiguous(), w[:, j].contiguous()
            o, sc, sh = self._routed_proj(feat, pk, rj, wj, B, H, W, hid, oc, G)
            last = j == idx.shape[1] - 1
            acc = ops.ew(ops.EW_AFFINE, a=o.view(B, H, W, oc), b=acc, p0=sc, p1=sh, rows_per_img=HW, act=False, out=out if last else None)
        return acc

    def fwd_nhwc(self, x, out=None):
        require_eval(self)
        B, H, W, C = x.shape
        HW, st_c = H * W, self.static_channels
        pk = self.get_pack()
        zero = torch.zeros((B, C), dtype=torch.float32, device=x.device)
        # SE-gated split (gated.py:325-332, _gated_visual.py:40-45)
        gate = ops.fc_gate(ops.gap(x), pk["se_w1"], pk["se_w2"], pk["se_b2"])
        xg = ops.ew(ops.EW_AFFINE, a=x, p0=gate, p1=zero, rows_per_img=HW)
        xs, xd = xg[..., :st_c], xg[..., st_c:]
        if "detail" in self.router_hook_names:   # detail gate on the dynamic half (gated.py:1174-1178)
            t = ops.dwconv(ops.dwconv(xd, pk["detail_hp"], None, 3, False, xd.shape[3]), pk["df0"], None, 3, False, xd.shape[3])
            t = _norm(t, pk["df1"], act=True)
            t = ops.conv2d(t, *pk["df3"], pk["df3"][0].shape[0], 1, 1, 1, 0, True)
            g = ops.ew(ops.EW_SIGMOID, a=ops.conv2d(t, *pk["df5"], xd.shape[3], 1, 1, 1, 0, False))
            xd = ops.ew(ops.EW_MUL_GATE, a=xd, b=g, p0=pk["detail_t"])
        # static path straight into its half of the concatenation buffer
        cat = ops.new_act(B, H, W, C, x.device)
        ts = ops.dwconv(xs, pk["st_dw"], pk["st_dw_b"], 3, True, st_c)
        ops.conv2d(ts, *pk["st_pw"], self.out_static, 1, 1, 1, 0, True, out=cat[..., :self.out_static])
        # routing (+ batch-level complexity gate) and routed experts into the other half
        idx, w, probs = ops.gate_router(xd, pk["router"], self.top_k)
        self.last_routing_snapshot = {"topk_indices": idx, "topk_weights": w, "router_probs": probs}   # device tensors, lazy
        self._experts(xd, idx, w, pk, cat[..., self.out_static:])
        if "xg_w1" in pk:   # CrossPathGate: cat * (0.5 + 0.5 * tanh(gate_scale) * sigmoid(MLP(GAP(cat))))
            xg = ops.fc_gate(ops.gap(cat), pk["xg_w1"], pk["xg_w2"], pk["xg_b2"], scale=pk["xg_scale"], offset=0.5)
            cat = ops.ew(ops.EW_AFFINE, a=cat, p0=xg, p1=zero, rows_per_img=HW)
        if self.shuffle_groups > 1:
            cat = ops.conv2d(cat, pk["shuffle"], None, C, 1, 1, 1, 0, False)
        for hook in self.router_hook_names:
            if hook == "context":     # pyramid context (gated.py:1210-1221)
                lc = _norm(ops.dwconv(cat, pk["lc0"], None, 3, False, C), pk["lc1"], act=True)
                ctx = [lc]
                for s, (pw, gn) in zip(self.context_mixer.pool_scales, pk["pp"]):
                    h, w_ = max(1, H // s), max(1, W // s)
                    pooled = cat if (h, w_) == (H, W) else ops.adaptive_avgpool(cat, h, w_)
                    ctx.append(_norm(ops.conv2d(pooled, *pw, C, 1, 1, 1, 0, False), gn, act=True))
                c = ops.ctx_mean3(*ctx)
                cgate = ops.ew(ops.EW_SIGMOID, a=ops.conv2d(c, *pk["cg"], C, 1, 1, 1, 0, False))
                cat = ops.ew(ops.EW_SCALE_RES, a=cat, b=ops.ew(ops.EW_MUL, a=c, b=cgate), p0=pk["ctx_t"])
            elif hook in ("refine", "light_refine"):   # cat + tanh(scale) * refiner(cat) * gate(cat)  (moe/hooks.py:50-57; the v0_12
                r = _norm(ops.dwconv(cat, pk["fr0"], None, 3, False, C), pk["fr1"], act=hook == "refine")   # variant has no SiLU)
                fg = ops.fc_gate(ops.gap(cat), pk["fg_w1"], pk["fg_w2"], pk["fg_b2"], scale=pk["refine_t"])
                cat = ops.ew(ops.EW_AFFINE, a=r, b=cat, p0=fg, p1=zero, rows_per_img=HW)
        # projection + GroupNorm + residual
        return _norm(ops.conv2d(cat, *pk["proj"], C, 1, 1, 1, 0, False), pk["bn"], add=x, out=out)

    def forward(self, x):
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.proj.weight.device)


_LOSS_ARGS = "balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01"   # accepted and unused at inference


class AdaptiveGateMoE(_GatedMoE):
    """`AdaptiveGateMoE(in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.0,
    final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01, router_hooks=None,
    detail_reduction=8, refine_reduction=8)` (gated.py:268-640): shared-inverted experts, no shuffle."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.0,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01, router_hooks=None,
                 detail_reduction=8, refine_reduction=8):
        hooks = tuple(router_hooks or ())
        if any(h not in ("detail", "context", "refine") for h in hooks):
            raise NotImplementedError(f"AdaptiveGateMoE: router hooks {hooks} are not on the B200 path (detail / context / refine)")
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         "shared_inverted", 1, hooks=hooks, refine_reduction=refine_reduction, detail_reduction=detail_reduction)


class FusedAdaptiveGateMoE(_GatedMoE):
    """`FusedAdaptiveGateMoE(in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8,
    initial_temperature=1.0, final_temperature=0.5, <loss coefficients>)` (gated.py:1232-1274): always the fused expert group."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.0,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01):
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         "fused", 1)


def _hybrid_backend(num_experts, threshold, low_rank):
    if num_experts > threshold:
        return "shared_inverted"
    return "low_rank_fused" if low_rank else "fused"


class HybridAdaptiveGateMoE(_GatedMoE):
    """`HybridAdaptiveGateMoE(in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8,
    initial_temperature=1.2, final_temperature=0.5, <loss coefficients>, fused_expert_threshold=8, shuffle_groups=2)`
    (gated.py:1277-1386): fused experts up to the threshold, shared-inverted above; channel shuffle."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.2,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
                 fused_expert_threshold=8, shuffle_groups=2):
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         _hybrid_backend(num_experts, fused_expert_threshold, False), shuffle_groups,
                         fused_expert_threshold=fused_expert_threshold)


class LowRankHybridAdaptiveGateMoE(_GatedMoE):
    """`LowRankHybridAdaptiveGateMoE(..., fused_expert_threshold=8, shuffle_groups=2, bottleneck_ratio=0.5)` (gated.py:1455-1508)."""
    HOOKS = ()

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.2,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
                 fused_expert_threshold=8, shuffle_groups=2, bottleneck_ratio=0.5, refine_reduction=8, detail_reduction=8):
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         _hybrid_backend(num_experts, fused_expert_threshold, True), shuffle_groups, bottleneck_ratio, self.HOOKS,
                         refine_reduction, detail_reduction, fused_expert_threshold)


_SHARED_EXPERT_POOLS: dict = {}


class SharedExpertMoE(LowRankHybridAdaptiveGateMoE):
    """`SharedExpertMoE(..., fused_expert_threshold=8, shuffle_groups=2, bottleneck_ratio=0.5, pool_id="shared")`
    (moe/shared_expert_moe.py:27-129): `LowRankHybridAdaptiveGateMoE` blocks with the same `pool_id` share ONE expert group module (the
    first block built owns it), so the state dict carries the same tensors under both blocks' keys, as in the reference."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.2,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
                 fused_expert_threshold=8, shuffle_groups=2, bottleneck_ratio=0.5, pool_id="shared"):
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         balance_loss_coeff, router_z_loss_coeff, entropy_loss_coeff, fused_expert_threshold, shuffle_groups, bottleneck_ratio)
        self.pool_id, self.bottleneck_ratio = pool_id, bottleneck_ratio
        sig = {"in_channels": self.dynamic_channels, "out_channels": self.out_dynamic, "num_experts": num_experts, "top_k": top_k,
               "bottleneck_ratio": bottleneck_ratio}
        pool = _SHARED_EXPERT_POOLS.get(pool_id)
        if pool is None:
            _SHARED_EXPERT_POOLS[pool_id] = {**sig, "fused_experts": self.fused_experts}
            self._is_pool_owner = True
        else:
            for k, v in sig.items():
                if pool[k] != v:
                    raise ValueError(f"SharedExpertMoE pool '{pool_id}' parameter mismatch: {k} expected {pool[k]}, got {v}")
            self.fused_experts = pool["fused_experts"]
            self._is_pool_owner = False

    @classmethod
    def reset_shared_pools(cls):
        """Clear the build-time registry before or after constructing a model (shared_expert_moe.py:114-117)."""
        _SHARED_EXPERT_POOLS.clear()


class RefinedLowRankHybridAdaptiveGateMoE(LowRankHybridAdaptiveGateMoE):
    """`RefinedLowRankHybridAdaptiveGateMoE(..., bottleneck_ratio=0.5, refine_reduction=8)` (gated.py:1511-1585): + refinement."""
    HOOKS = ("refine",)

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.2,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
                 fused_expert_threshold=8, shuffle_groups=2, bottleneck_ratio=0.5, refine_reduction=8):
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         balance_loss_coeff, router_z_loss_coeff, entropy_loss_coeff, fused_expert_threshold, shuffle_groups,
                         bottleneck_ratio, refine_reduction)


class DetailAwareLowRankHybridAdaptiveGateMoE(LowRankHybridAdaptiveGateMoE):
    """`DetailAwareLowRankHybridAdaptiveGateMoE(..., bottleneck_ratio=0.5, detail_reduction=8)` (gated.py:1588-1642): + detail gate."""
    HOOKS = ("detail",)

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.2,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
                 fused_expert_threshold=8, shuffle_groups=2, bottleneck_ratio=0.5, detail_reduction=8):
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         balance_loss_coeff, router_z_loss_coeff, entropy_loss_coeff, fused_expert_threshold, shuffle_groups,
                         bottleneck_ratio, 8, detail_reduction)


class ContextRefinedLowRankHybridAdaptiveGateMoE(RefinedLowRankHybridAdaptiveGateMoE):
    """`ContextRefinedLowRankHybridAdaptiveGateMoE(..., bottleneck_ratio=0.5, refine_reduction=8)` (gated.py:1645-1700): + pyramid context."""
    HOOKS = ("context", "refine")


class VisualEnhancedAdaptiveGateMoE(LowRankHybridAdaptiveGateMoE):
    """`VisualEnhancedAdaptiveGateMoE(in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8,
    initial_temperature=1.2, final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
    fused_expert_threshold=8, shuffle_groups=2, bottleneck_ratio=0.5, refine_reduction=8, detail_reduction=8)` (gated.py:1703-1756)."""
    HOOKS = ("detail", "context", "refine")


class HybridAdaptiveGateMoEv2(_GatedMoE):
    """`HybridAdaptiveGateMoEv2(..., fused_expert_threshold=8, shuffle_groups=2)` (gated.py:1389-1452, v0_11): `HybridAdaptiveGateMoE`
    with `DualStreamGateRouterV2`."""
    HOOKS = ()

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.2,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
                 fused_expert_threshold=8, shuffle_groups=2, refine=True, refine_reduction=8):
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         _hybrid_backend(num_experts, fused_expert_threshold, False), shuffle_groups,
                         hooks=self.HOOKS if refine else (), refine_reduction=refine_reduction,
                         fused_expert_threshold=fused_expert_threshold, router_v2=True)
        self.refine = bool(refine) and bool(self.HOOKS)


class OptimalHybridGateMoE(HybridAdaptiveGateMoEv2):
    """`OptimalHybridGateMoE(..., fused_expert_threshold=8, shuffle_groups=2, refine=True, refine_reduction=8)` (gated.py:1846-2023,
    v0_12): v0_11 plus a depthwise refinement gated by a global SE vector."""
    HOOKS = ("light_refine",)


class DiversifiedExpertMoE(OptimalHybridGateMoE):
    """`DiversifiedExpertMoE(..., fused_expert_threshold=8, shuffle_groups=2, refine=True, refine_reduction=8)` (gated.py:2499-2561,
    v0_14): `OptimalHybridGateMoE` whose expert group is replaced by a `DiversifiedExpertGroup` (expand_ratio 2, threshold 0)."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, *args, **kwargs):
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, *args, **kwargs)
        self.fused_experts = DiversifiedExpertGroup(self.dynamic_channels, self.out_dynamic, num_experts, expand_ratio=2.0, top_k=top_k,
                                                    weight_threshold=0.0, num_groups=num_groups)


class MultiHeadRouterMoE(_GatedMoE):
    """`MultiHeadRouterMoE(..., fused_expert_threshold=8, shuffle_groups=2, refine=True, refine_reduction=8, num_heads=4,
    expert_dropout=0.05)` (gated.py:2430-2496, v0_13): `OptimalHybridGateMoE` with `MultiHeadRouterV3`."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.2,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
                 fused_expert_threshold=8, shuffle_groups=2, refine=True, refine_reduction=8, num_heads=4, expert_dropout=0.05):
        router = MultiHeadRouterV3(int(in_channels * split_ratio), num_experts, top_k, temperature=initial_temperature,
                                   num_heads=num_heads, expert_dropout=expert_dropout)
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         _hybrid_backend(num_experts, fused_expert_threshold, False), shuffle_groups,
                         hooks=("light_refine",) if refine else (), refine_reduction=refine_reduction,
                         fused_expert_threshold=fused_expert_threshold, router=router)
        self.refine = bool(refine)


class GatedFusionMoE(_GatedMoE):
    """`GatedFusionMoE(..., fused_expert_threshold=8, shuffle_groups=2, refine=True, refine_reduction=8, drop_prob=0.05)`
    (gated.py:2564-2707, v0_15): `OptimalHybridGateMoE` with a `CrossPathGate` on the concatenated paths (stochastic depth is a
    training-time feature)."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=1.2,
                 final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
                 fused_expert_threshold=8, shuffle_groups=2, refine=True, refine_reduction=8, drop_prob=0.05):
        super().__init__(in_channels, out_channels, num_experts, top_k, split_ratio, num_groups, initial_temperature, final_temperature,
                         _hybrid_backend(num_experts, fused_expert_threshold, False), shuffle_groups,
                         hooks=("light_refine",) if refine else (), refine_reduction=refine_reduction,
                         fused_expert_threshold=fused_expert_threshold, router_v2=True, cross_gate=True)
        self.refine = bool(refine)


class ZeroCostRouter(nn.Module):
    """`ZeroCostRouter(in_channels, num_experts, top_k, temperature=1.0)` (gated.py:938-1000; `UltraLightRouter` :2710-2724 adds
    nothing at inference): Linear over [mean | std] -> Softmax, executed inside `ym_zero_cost_router`."""

    def __init__(self, in_channels, num_experts, top_k, temperature=1.0, use_cache=True):
        super().__init__()
        self.num_experts, self.top_k, self.temperature = num_experts, top_k, temperature
        self.router = nn.Sequential(nn.Linear(2 * in_channels, num_experts, bias=False), nn.Softmax(dim=1))


class _BalanceController(nn.Module):
    """State of `AdaptiveBalanceController` (gated.py:1767-1843): one parameter, used by the training loss only."""

    def __init__(self, num_experts):
        super().__init__()
        self.expert_importance = nn.Parameter(torch.ones(num_experts))


class UltimateOptimizedMoE(nn.Module, PackCache):
    """`UltimateOptimizedMoE(in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, use_routing_cache=True,
    capacity_factor=1.5, initial_temperature=2.0, final_temperature=0.5, entropy_coeff=0.01)` (moe/modules.py:1534-1742, v0_3 zoo):
    plain channel split, static path, ZeroCostRouter, complexity-scaled weights, fused expert group, projection + GroupNorm + x."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, use_routing_cache=True,
                 capacity_factor=1.5, initial_temperature=2.0, final_temperature=0.5, entropy_coeff=0.01):
        super().__init__()
        if in_channels != out_channels:
            raise ValueError("UltimateOptimizedMoE: the residual `proj(...) + x` needs in_channels == out_channels")
        self.in_channels, self.out_channels, self.num_experts, self.top_k = in_channels, out_channels, num_experts, top_k
        self.initial_temperature, self.final_temperature = initial_temperature, final_temperature
        self.dynamic_channels = int(in_channels * split_ratio)
        self.static_channels = in_channels - self.dynamic_channels
        self.out_dynamic = int(out_channels * split_ratio)
        self.out_static = out_channels - self.out_dynamic
        for n in (self.dynamic_channels, self.static_channels, self.out_dynamic, self.out_static):
            if n % 8:
                raise NotImplementedError("UltimateOptimizedMoE: channel halves must be multiples of 8 on the B200 path")
        sc = self.static_channels
        self.static_net = nn.Sequential(
            nn.Conv2d(sc, sc, 3, padding=1, groups=sc, bias=False), nn.BatchNorm2d(sc, eps=BN_EPS, momentum=0.03), nn.SiLU(inplace=True),
            nn.Conv2d(sc, self.out_static, 1, bias=False), nn.BatchNorm2d(self.out_static, eps=BN_EPS, momentum=0.03), nn.SiLU(inplace=True))
        self.routing = ZeroCostRouter(self.dynamic_channels, num_experts, top_k, temperature=initial_temperature, use_cache=use_routing_cache)
        self.fused_experts = FusedExpertGroup(self.dynamic_channels, self.out_dynamic, num_experts, num_groups)
        self.complexity_estimator = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(self.dynamic_channels, 1, 1), nn.Sigmoid())
        self.register_buffer("current_top_k", torch.tensor(num_experts))
        self.balance_controller = _BalanceController(num_experts)
        self.proj = nn.Conv2d(out_channels, out_channels, 1, bias=False)
        self.bn = _gn(out_channels, num_groups)
        self.expert_backend = "fused"
        self.last_routing_snapshot: dict = {}

    def _build_pack(self):
        pk = {}
        w, b = fold_bn(self.static_net[0].weight, None, self.static_net[1])
        pk["st_dw"], pk["st_dw_b"] = _pack_dw(w), b.contiguous()
        w, b = fold_bn(self.static_net[3].weight, None, self.static_net[4])
        pk["st_pw"] = (pack_gemm_weight(w), b.contiguous())
        pk["fc"] = _f32(self.routing.router[0].weight)
        ce = self.complexity_estimator[1]
        pk["cx_w"], pk["cx_b"] = _f32(ce.weight).reshape(-1).contiguous(), float(ce.bias.detach().float())
        pk["fused_w"] = pack_gemm_weight(self.fused_experts.dense_weight())
        pk["fused_gamma"], pk["fused_beta"] = _f32(self.fused_experts.expert_norm_weight), _f32(self.fused_experts.expert_norm_bias)
        pk["proj"], pk["bn"] = _pack_linear(self.proj.weight), _gn_args(self.bn)
        return pk

    def fwd_nhwc(self, x, out=None):
        require_eval(self)
        B, H, W, C = x.shape
        st_c = self.static_channels
        pk = self.get_pack()
        xs, xd = x[..., :st_c], x[..., st_c:]
        cat = ops.new_act(B, H, W, C, x.device)
        ts = ops.dwconv(xs, pk["st_dw"], pk["st_dw_b"], 3, True, st_c)
        ops.conv2d(ts, *pk["st_pw"], self.out_static, 1, 1, 1, 0, True, out=cat[..., :self.out_static])
        idx, w, probs = ops.zero_cost_router(xd, pk["fc"], self.routing.temperature, pk["cx_w"], pk["cx_b"], self.top_k)
        self.last_routing_snapshot = {"topk_indices": idx, "topk_weights": w, "router_probs": probs}
        E, oc = self.num_experts, self.out_dynamic
        fo = ops.conv2d(xd, pk["fused_w"], None, E * oc, 3, 3, 1, 1, False)
        ops.gated_select(fo, idx, w, pk["fused_gamma"], pk["fused_beta"], E, oc, self.fused_experts.norm_groups, 1e-5,
                         out=cat[..., self.out_static:])
        return _norm(ops.conv2d(cat, *pk["proj"], C, 1, 1, 1, 0, False), pk["bn"], add=x, out=out)

    def forward(self, x):
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.proj.weight.device)
