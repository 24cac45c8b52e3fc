"""Routed MoE-FFN blocks with the reference's names, signatures and state_dict keys.

Mirrors `ultralytics/nn/modules/moe/`: EfficientSpatialRouter routers.py:268-304, SimpleExpert experts.py:73-88,
OptimizedMOEImproved modules.py:957-1180, ABlockMoE :1199-1260, A2C2fMoE :1268-1297.

Routing is per IMAGE (routers.py:300), so an "expert batch" is a set of whole images.  Instead of the reference's
Python loop over experts (`mask.any()` host sync, `x[batch_idx]` gather copy, `index_add_`), the router writes a
device index table and one grouped GEMM per layer of the expert MLP runs all B*top_k (image, expert) problems with
weight-pointer indirection; GroupNorm statistics are accumulated in the GEMM epilogues and applied on the next load.
No host synchronisation happens anywhere in the block, so the forward is CUDA-graph capturable.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops
from ...utils.errors import MoERouterError, ShapeMismatchError
from ._base import PackCache, bn_affine, fold_bn, pack_gemm_weight, require_eval, to_nchw, to_nhwc
from .block import A2C2f, ABlock, C3k, _SeqNHWC

__all__ = ("UltraEfficientRouter", "UltraOptimizedMoE", "EfficientSpatialRouter", "SimpleExpert", "OptimizedMOEImproved", "ABlockMoE", "A2C2fMoE", "get_safe_groups",
           "DynamicRoutingLayer", "DepthwiseSeparableConv", "EfficientExpertGroup", "ES_MOE")


def get_safe_groups(channels: int, desired_groups: int = 8) -> int:
    """Largest num_groups <= desired_groups dividing channels (reference nn/modules/utils.py:108-115)."""
    if channels <= 0:
        return 1
    groups = min(desired_groups, channels)
    while channels % groups != 0:
        groups -= 1
    return max(1, groups)


def _validate_router_input(x: torch.Tensor, expected_channels: int, context: str = "") -> None:
    """routers.py:33-52: rank, channel count, finiteness (the last one reads the tensor and synchronises, as the reference does;
    it runs only at a router's public `forward`, never inside a block's fused path)."""
    tag = f" [{context}]" if context else ""
    if x.dim() != 4:
        raise MoERouterError(f"Router input must be 4-D (NCHW), got {x.dim()}-D shape {tuple(x.shape)}{tag}")
    if expected_channels > 0 and x.shape[1] != expected_channels:
        raise ShapeMismatchError(expected=f"(N, {expected_channels}, H, W)", actual=tuple(x.shape), context=context or "router input")
    if not bool(torch.isfinite(x).all()):
        raise MoERouterError(f"Router input contains NaN/Inf values{tag}")


class EfficientSpatialRouter(nn.Module, PackCache):
    """`EfficientSpatialRouter(in_channels, num_experts, reduction=8, top_k=2, noise_std=1.0, pool_scale=4)`."""

    def __init__(self, in_channels, num_experts, reduction=8, top_k=2, noise_std=1.0, pool_scale=4):
        super().__init__()
        self.num_experts = num_experts
        self.top_k = top_k
        self.noise_std = noise_std
        self.pool_scale = pool_scale
        self.capacity_factor = None
        self.softmax = nn.Softmax(dim=1)
        reduced_channels = max(in_channels // reduction, 8)
        self.router = nn.Sequential(
            nn.Conv2d(in_channels, reduced_channels, 3, padding=1, bias=False),
            nn.BatchNorm2d(reduced_channels, eps=1e-3, momentum=0.03),
            nn.SiLU(inplace=False),
            nn.Conv2d(reduced_channels, num_experts, 1, bias=False),
            nn.BatchNorm2d(num_experts, eps=1e-3, momentum=0.03),
        )

    def _build_pack(self):
        c0, bn1, c3, bn2 = self.router[0], self.router[1], self.router[3], self.router[4]
        Cr, C = c0.weight.shape[0], c0.weight.shape[1]
        s1, h1 = bn_affine(bn1)
        s2, h2 = bn_affine(bn2)
        return {
            "Cr": Cr, "E": c3.weight.shape[0],
            # [tap][c/4][r][4]: the Cr lanes of one pixel read consecutive float4
            "w1": c0.weight.detach().float().permute(2, 3, 1, 0).reshape(9, C // 4, 4, Cr).permute(0, 1, 3, 2).contiguous(),
            "scale1": s1, "shift1": h1,
            "w2": c3.weight.detach().float().reshape(c3.weight.shape[0], Cr).contiguous(),
            "scale2": s2, "shift2": h2,
        }

    def route_nhwc(self, x, top_k=None):
        k = self.top_k if top_k is None else max(1, min(int(top_k), self.num_experts))
        return ops.router_topk(x, self.get_pack(), k, self.pool_scale)

    def forward(self, x, top_k=None):
        """Returns (weights fp32 [B,k], indices int64 [B,k], {}) like the reference's eval branch."""
        require_eval(self)
        _validate_router_input(x, self.router[0].in_channels, context="EfficientSpatialRouter")
        if not math.isfinite(float(self.noise_std)):
            raise MoERouterError("EfficientSpatialRouter noise_std must be finite")
        idx, w, _ = self.route_nhwc(to_nhwc(x), top_k)
        if not bool(torch.isfinite(w).all()):     # routers.py:295-302: non-finite weights / statistics inside the router
            raise MoERouterError("EfficientSpatialRouter internal output contains NaN/Inf")
        return w, idx.long(), {}


# "tc": routed expert FFN on the tcgen05 kernels with the hidden kept on chip (csrc/tc_moe.cu) where the widths allow;
# "mma": the mma.sync implicit-GEMM chain of csrc/gemm_conv.cu (h through global memory) - kept as the A/B baseline
MOE_FFN_IMPL = "tc"
# GroupNorm-1 / GroupNorm-2 of the tcgen05 expert chain finalised inside the consumer kernels (ym_moe_ffn_gn, ym_moe_combine_tc_gn)
# instead of by two ym_gn_finalize_tiles launches per block: same bits, 12 launches less per yolo26-master-n forward
MOE_GN_FOLD = True
# the router's finish inside the prologue of the statistics pass (ym_moe_ffn_routed) instead of its own launch: 6 launches less per forward
MOE_ROUTER_FOLD = True


class SimpleExpert(nn.Module):
    """`SimpleExpert(in_channels, out_channels, expand_ratio=2, num_groups=8)`: 1x1 -> GN -> SiLU -> 1x1 -> GN.

    Parameter container only; the arithmetic runs batched over all routed (image, expert) problems in
    `OptimizedMOEImproved.fwd_nhwc`."""

    def __init__(self, in_channels, out_channels, expand_ratio=2, num_groups=8):
        super().__init__()
        hidden_dim = int(in_channels * expand_ratio)
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, hidden_dim, 1, bias=False),
            nn.GroupNorm(get_safe_groups(hidden_dim, num_groups), hidden_dim),
            nn.SiLU(inplace=True),
            nn.Conv2d(hidden_dim, out_channels, 1, bias=False),
            nn.GroupNorm(get_safe_groups(out_channels, num_groups), out_channels),
        )

    def forward(self, x):
        raise RuntimeError("SimpleExpert runs only inside OptimizedMOEImproved on the B200 path (grouped expert GEMM)")


class OptimizedMOEImproved(nn.Module, PackCache):
    """Same constructor as the reference (modules.py:960-976); only expert_type='simple', router_type='efficient'."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, expert_type="simple", router_type="efficient",
                 noise_std=1.0, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, expert_expand_ratio=2.0,
                 progressive_sparsity=True, detach_routing=False, add_residual=True):
        super().__init__()
        if expert_type != "simple" or router_type != "efficient":
            raise NotImplementedError("OptimizedMOEImproved: only expert_type='simple', router_type='efficient' are on the B200 path")
        if in_channels != out_channels:
            raise NotImplementedError("OptimizedMOEImproved: in_channels != out_channels is not on the B200 path")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_experts, self.top_k = num_experts, top_k
        self.balance_loss_coeff, self.router_z_loss_coeff = balance_loss_coeff, router_z_loss_coeff
        self.progressive_sparsity, self.add_residual, self.detach_routing = progressive_sparsity, add_residual, detach_routing
        self.routing = EfficientSpatialRouter(in_channels, num_experts, top_k=top_k, noise_std=noise_std)
        self.experts = nn.ModuleList(
            SimpleExpert(in_channels, out_channels, expand_ratio=expert_expand_ratio) for _ in range(num_experts))
        self.shared_expert = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels, eps=1e-3, momentum=0.03),
            nn.SiLU(inplace=True))
        self._init_weights()
        self.last_routing_snapshot = {}

    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.routing.router[3].weight, mean=0, std=0.05)

    def _pack_sources(self):  # the router keeps its own pack
        return list(self.experts.parameters()) + list(self.shared_expert.parameters()) + list(self.shared_expert.buffers())

    def _build_pack(self):
        ex = list(self.experts)
        w1 = torch.stack([pack_gemm_weight(e.conv[0].weight.detach().float()) for e in ex]).contiguous()  # [E][hid][Kpad]
        w2 = torch.stack([pack_gemm_weight(e.conv[3].weight.detach().float()) for e in ex]).contiguous()  # [E][C][Kpad]
        g1, g2 = ex[0].conv[1], ex[0].conv[4]
        ws, bs = fold_bn(self.shared_expert[0].weight, None, self.shared_expert[1])
        return {
            "w1": w1, "w2": w2, "hid": w1.shape[1],
            "gamma1": torch.stack([e.conv[1].weight.detach().float() for e in ex]).contiguous(),
            "beta1": torch.stack([e.conv[1].bias.detach().float() for e in ex]).contiguous(),
            "gamma2": torch.stack([e.conv[4].weight.detach().float() for e in ex]).contiguous(),
            "beta2": torch.stack([e.conv[4].bias.detach().float() for e in ex]).contiguous(),
            "G1": g1.num_groups, "G2": g2.num_groups, "eps1": g1.eps, "eps2": g2.eps,
            "ws": pack_gemm_weight(ws), "bs": bs.contiguous(),
        }

    def fwd_nhwc(self, x, out=None, outer_residual=False):
        """x: (B,H,W,C).  Returns shared(x) + sum_j w_j*expert_j(x) [+ x if add_residual or outer_residual]."""
        require_eval(self)
        B, H, W, C = x.shape
        HW, k = H * W, self.top_k
        pk = self.get_pack()
        P, hid = B * k, pk["hid"]
        ldx = ops.pitch(x)
        tc = MOE_FFN_IMPL == "tc" and pk["w1"].shape[2] == C and pk["w2"].shape[2] == hid and (hid // pk["G1"]) % 8 == 0 \
            and (C // pk["G2"]) % 8 == 0 and ops.moe_ffn_supported(C, hid, ldx)
        st1 = None
        if tc and MOE_ROUTER_FOLD and 1 <= k <= min(8, self.num_experts):
            # the router's finish (mean -> logits -> softmax -> top-k) runs in the prologue of the expert FFN's statistics pass, which
            # publishes idx / w / probs for the kernels after it: one launch less per block, same values
            rpack = self.routing.get_pack()
            partial, nblk, npix = ops.router_partial(x, rpack, self.routing.pool_scale)
            idx, w, probs, st1, strips = ops.moe_ffn_stats_routed(x, k, pk["w1"], rpack, partial, nblk, npix)
        else:
            idx, w, probs = self.routing.route_nhwc(x, k)       # device index table, no host sync
        ridx, rw = idx.view(-1), w.view(-1)
        if tc:
            # tcgen05 path (csrc/tc_moe.cu): the hidden activation stays in tensor memory - pass 1 takes GroupNorm-1 statistics of
            # h = x W1[e]^T without storing it, pass 2 recomputes h, normalises, SiLU, and feeds the second GEMM from tensor memory
            if st1 is None:
                st1, strips = ops.moe_ffn_stats(x, k, pk["w1"], ridx)
            if MOE_GN_FOLD:
                # both GroupNorms are finalised by their consumer kernels from the partial sums (two launches less per block, same bits)
                o, st2 = ops.moe_ffn_fused_gn(x, k, pk["w1"], pk["w2"], ridx, st1, strips, pk["G1"], HW * (hid // pk["G1"]), pk["eps1"],
                                              pk["gamma1"], pk["beta1"])
                add_res = outer_residual or (self.add_residual and self.in_channels == self.out_channels)
                if ops.moe_combine_gn_supported(x, pk["ws"], o, k, out):
                    y = ops.moe_combine_gn(x, pk["ws"], pk["bs"], o, st2, strips, pk["G2"], HW * (C // pk["G2"]), pk["eps2"], pk["gamma2"],
                                           pk["beta2"], ridx, rw, k, add_residual=add_res, out=out)
                    self.last_routing_snapshot = {"topk_indices": idx, "topk_weights": w, "router_probs": probs}  # device tensors, lazy
                    return y
            else:
                sc1, sh1 = ops.gn_finalize_tiles(st1, P, strips, pk["G1"], hid, HW * (hid // pk["G1"]), pk["eps1"], pk["gamma1"], pk["beta1"], ridx)
                o, st2 = ops.moe_ffn_fused(x, k, pk["w1"], pk["w2"], ridx, sc1, sh1, strips)
            sc2, sh2 = ops.gn_finalize_tiles(st2, P, strips, pk["G2"], C, HW * (C // pk["G2"]), pk["eps2"], pk["gamma2"], pk["beta2"], ridx, route_w=rw)
        else:
            # GEMM1: h[p] = x[p // k] @ W1[e_p]^T, GroupNorm-1 statistics in the epilogue
            h, st1 = ops.moe_expert_gemm(x, ldx, k, P, HW, C, pk["w1"], ridx, hid, groups=pk["G1"])
            sc1, sh1 = ops.gn_finalize(st1, P, HW, pk["G1"], hid, HW * (hid // pk["G1"]), pk["eps1"], pk["gamma1"], pk["beta1"], ridx)
            # GEMM2: o[p] = SiLU(GN1(h[p])) @ W2[e_p]^T with GN+SiLU applied on the A-operand load; GN-2 statistics
            o, st2 = ops.moe_expert_gemm(h, hid, 1, P, HW, hid, pk["w2"], ridx, C, a_scale=sc1, a_shift=sh1, groups=pk["G2"])
            sc2, sh2 = ops.gn_finalize(st2, P, HW, pk["G2"], C, HW * (C // pk["G2"]), pk["eps2"], pk["gamma2"], pk["beta2"], ridx, route_w=rw)
        # combine: shared expert GEMM + sum_j w_j*GN2(o_j) (+ residual), fp32 accumulate, one rounding
        add_res = outer_residual or (self.add_residual and self.in_channels == self.out_channels)
        y = ops.moe_combine(x, pk["ws"], pk["bs"], o, sc2, sh2, k, add_residual=add_res, out=out)
        self.last_routing_snapshot = {"topk_indices": idx, "topk_weights": w, "router_probs": probs}  # device tensors, lazy
        return y

    def forward(self, x):
        """Standalone NCHW entry.  The reference guards four intermediate tensors (modules.py:1086,1144,1148,1151); the fused
        path has one place to look - the block output - and raises the same RuntimeError type there.  Inside A2C2fMoE the block
        runs through `fwd_nhwc` with no host synchronisation."""
        y = self.fwd_nhwc(to_nhwc(x))
        capturing = y.is_cuda and torch.cuda.is_current_stream_capturing()      # a host read would break a graph capture
        if not capturing and not bool(torch.isfinite(y).all()):
            raise RuntimeError("OptimizedMOEImproved final output contains NaN/Inf (shared expert / sparse expert aggregation / "
                               "dtype conversion)")
        return to_nchw(y)

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.shared_expert[0].weight.device)


class UltraEfficientRouter(nn.Module):
    """`UltraEfficientRouter(in_channels, num_experts, reduction=16, top_k=2, noise_std=1.0, temperature=1.0, pool_scale=8)`
    (moe/routers.py:58-137): depthwise / pointwise local stream on an 8x-pooled map, per-pixel softmax, spatial mean (`ym_pixel_router`)."""

    def __init__(self, in_channels, num_experts, reduction=16, top_k=2, noise_std=1.0, temperature=1.0, pool_scale=8):
        super().__init__()
        self.num_experts, self.top_k, self.noise_std = num_experts, top_k, noise_std
        self.temperature = max(float(temperature), 1e-3)
        self.pool_scale = pool_scale
        reduced = max(in_channels // reduction, 4)
        self.router = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, 3, padding=1, groups=in_channels, bias=False),
            nn.GroupNorm(get_safe_groups(in_channels, 8), in_channels), nn.SiLU(inplace=False),
            nn.Conv2d(in_channels, reduced, 1, bias=False), nn.GroupNorm(get_safe_groups(reduced, 4), reduced), nn.SiLU(inplace=False),
            nn.Conv2d(reduced, num_experts, 1, bias=True))
        self.softmax = nn.Softmax(dim=1)

    def pack(self):
        r = self.router
        C, R, E = r[0].weight.shape[0], r[3].weight.shape[0], self.num_experts
        f = lambda t: t.detach().float().contiguous()
        return {"E": E, "R": R, "pool": int(self.pool_scale), "G1": r[1].num_groups, "G2": r[4].num_groups, "eps": float(r[1].eps),
                "dw": f(r[0].weight).reshape(C, 9).contiguous(), "gn1_w": f(r[1].weight), "gn1_b": f(r[1].bias),
                "pw1": f(r[3].weight).reshape(R, C).contiguous(), "gn2_w": f(r[4].weight), "gn2_b": f(r[4].bias),
                "pw2": f(r[6].weight).reshape(E, R).contiguous(), "b2": f(r[6].bias), "temperature": float(self.temperature)}


class UltraOptimizedMoE(nn.Module, PackCache):
    """`UltraOptimizedMoE(in_channels, out_channels, num_experts=4, top_k=2, expert_type="simple", router_reduction=16,
    router_pool_scale=8, noise_std=1.0, router_temperature=1.0, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, num_groups=8,
    weight_threshold=0.01)` (moe/modules.py:121-320; v0_1 uomoe / v0_2 zoos): UltraEfficientRouter, GroupNorm shared expert and the
    routed 1x1 -> GN -> SiLU -> 1x1 -> GN experts on the grouped expert GEMM; routes with weight <= 0.01 are dropped; no residual."""

    def __init__(self, in_channels, out_channels, num_experts=4, top_k=2, expert_type="simple", router_reduction=16, router_pool_scale=8,
                 noise_std=1.0, router_temperature=1.0, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, num_groups=8, weight_threshold=0.01):
        super().__init__()
        if expert_type != "simple":
            raise NotImplementedError("UltraOptimizedMoE: only expert_type='simple' is on the B200 path")
        if in_channels != out_channels:
            raise NotImplementedError("UltraOptimizedMoE: in_channels != out_channels is not on the B200 path")
        self.in_channels, self.out_channels, self.num_experts, self.top_k = in_channels, out_channels, num_experts, top_k
        self.weight_threshold = weight_threshold
        self.routing = UltraEfficientRouter(in_channels, num_experts, reduction=router_reduction, top_k=top_k, noise_std=noise_std,
                                            temperature=router_temperature, pool_scale=router_pool_scale)
        self.experts = nn.ModuleList(SimpleExpert(in_channels, out_channels, expand_ratio=2, num_groups=num_groups) for _ in range(num_experts))
        self.shared_expert = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1, bias=False),
                                           nn.GroupNorm(get_safe_groups(out_channels, num_groups), out_channels), nn.SiLU(inplace=True))
        self.last_routing_snapshot = {}

    def _build_pack(self):
        ex = list(self.experts)
        g1, g2, gs = ex[0].conv[1], ex[0].conv[4], self.shared_expert[1]
        f = lambda t: t.detach().float().contiguous()
        return {
            "w1": torch.stack([pack_gemm_weight(e.conv[0].weight.detach().float()) for e in ex]).contiguous(),
            "w2": torch.stack([pack_gemm_weight(e.conv[3].weight.detach().float()) for e in ex]).contiguous(),
            "gamma1": torch.stack([f(e.conv[1].weight) for e in ex]).contiguous(), "beta1": torch.stack([f(e.conv[1].bias) for e in ex]).contiguous(),
            "gamma2": torch.stack([f(e.conv[4].weight) for e in ex]).contiguous(), "beta2": torch.stack([f(e.conv[4].bias) for e in ex]).contiguous(),
            "G1": g1.num_groups, "G2": g2.num_groups, "eps1": g1.eps, "eps2": g2.eps,
            "ws": pack_gemm_weight(self.shared_expert[0].weight.detach().float()), "gs": (gs.num_groups, f(gs.weight), f(gs.bias), float(gs.eps)),
            "router": self.routing.pack(),
        }

    def fwd_nhwc(self, x, out=None):
        require_eval(self)
        B, H, W, C = x.shape
        HW = H * W
        pk = self.get_pack()
        idx, w, probs = ops.pixel_router(x, pk["router"], self.top_k, self.weight_threshold)      # weights <= threshold arrive as 0
        self.last_routing_snapshot = {"topk_indices": idx, "topk_weights": w, "router_probs": probs}
        G, gw, gb, geps = pk["gs"]
        acc = ops.groupnorm(ops.conv2d(x, pk["ws"], None, C, 1, 1, 1, 0, False), G, gw, gb, eps=geps, act=True)   # shared expert
        hid, ldx = pk["w1"].shape[1], ops.pitch(x)
        for j in range(self.top_k):      # one grouped-GEMM chain per routing rank: a zero weight contributes zero
            rj, wj = idx[:, j].contiguous(), w[:, j].contiguous()
            h, st1 = ops.moe_expert_gemm(x, ldx, 1, B, HW, C, pk["w1"], rj, hid, groups=pk["G1"])
            sc1, sh1 = ops.gn_finalize(st1, B, HW, pk["G1"], hid, HW * (hid // pk["G1"]), pk["eps1"], pk["gamma1"], pk["beta1"], rj)
            o, st2 = ops.moe_expert_gemm(h, hid, 1, B, HW, hid, pk["w2"], rj, C, a_scale=sc1, a_shift=sh1, groups=pk["G2"])
            sc2, sh2 = ops.gn_finalize(st2, B, HW, pk["G2"], C, HW * (C // pk["G2"]), pk["eps2"], pk["gamma2"], pk["beta2"], rj, route_w=wj)
            last = j == self.top_k - 1
            acc = ops.ew(ops.EW_AFFINE, a=o.view(B, H, W, C), b=acc, p0=sc2, p1=sh2, rows_per_img=HW, act=False, out=out if last else None)
        return acc

    def forward(self, x):
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.shared_expert[0].weight.device)


class ABlockMoE(ABlock):
    """`ABlockMoE(dim, num_heads, mlp_ratio=1.2, area=1, num_experts=4, top_k=2, expert_type='simple')`."""

    def __init__(self, dim, num_heads, mlp_ratio=1.2, area=1, num_experts=4, top_k=2, expert_type="simple"):
        super().__init__(dim, num_heads, mlp_ratio, area)
        self.mlp = OptimizedMOEImproved(in_channels=dim, out_channels=dim, num_experts=num_experts, top_k=top_k,
                                        expert_type=expert_type, expert_expand_ratio=mlp_ratio, progressive_sparsity=True,
                                        add_residual=False)

    def fwd_nhwc(self, x, out=None):
        x = self.attn.fwd_nhwc(x, res=x)                              # x + attn(x)
        return self.mlp.fwd_nhwc(x, out=out, outer_residual=True)    # x + mlp(x): residual fused in the combine epilogue

    @property
    def aux_loss(self):
        return self.mlp.aux_loss


class A2C2fMoE(A2C2f):
    """`A2C2fMoE(c1, c2, n=1, a2=True, area=1, residual=False, mlp_ratio=2.0, e=0.5, g=1, shortcut=True,
    num_experts=4, top_k=2, expert_type='simple')`."""

    def __init__(self, c1, c2, n=1, a2=True, area=1, residual=False, mlp_ratio=2.0, e=0.5, g=1, shortcut=True,
                 num_experts=4, top_k=2, expert_type="simple"):
        super().__init__(c1, c2, n, a2, area, residual, mlp_ratio, e, g, shortcut)
        c_ = int(c2 * e)
        self.m = nn.ModuleList(
            _SeqNHWC(*(ABlockMoE(c_, c_ // 32, mlp_ratio, area, num_experts, top_k, expert_type) for _ in range(2)))
            if a2 else C3k(c_, c_, 2, shortcut, g)
            for _ in range(n)
        )

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.cv1.conv.weight.device)


# ======================================================================================================================
# ES_MOE family (moe/routers.py:429-527, moe/experts.py:280-311, moe/modules.py:410-741)
# ======================================================================================================================
class DynamicRoutingLayer(nn.Module):
    """`DynamicRoutingLayer(in_channels, num_experts=3, reduction=8, top_k=None)`: GAP -> 1x1 -> SiLU -> 1x1 (with biases).
    Parameter container: the router arithmetic runs inside `ES_MOE.fwd_nhwc` (ym_esmoe_route)."""

    def __init__(self, in_channels, num_experts=3, reduction=8, top_k=None):
        super().__init__()
        if num_experts < 1:
            raise ValueError(f"num_experts must be positive, got {num_experts}")
        if reduction < 1:
            raise ValueError(f"reduction must be positive, got {reduction}")
        if top_k is not None and not 1 <= top_k <= num_experts:
            raise ValueError(f"top_k must be in [1, {num_experts}], got {top_k}")
        reduced_channels = max(in_channels // reduction, 8)
        self.in_channels = in_channels
        self.num_experts = num_experts
        self.top_k = min(top_k, num_experts) if top_k is not None else num_experts
        self.use_top_k = top_k is not None
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.routing_network = nn.Sequential(
            nn.Conv2d(in_channels, reduced_channels, kernel_size=1), nn.SiLU(inplace=False),
            nn.Conv2d(reduced_channels, num_experts, kernel_size=1))


class DepthwiseSeparableConv(nn.Module):
    """`DepthwiseSeparableConv(in_channels, out_channels, kernel_size, stride=1)`: dw kxk -> 1x1 -> BN -> SiLU."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("DepthwiseSeparableConv: stride != 1 is not on the B200 path")
        padding = (kernel_size - 1) // 2
        self.depthwise = nn.Conv2d(in_channels, in_channels, kernel_size, stride=stride, padding=padding, groups=in_channels, bias=False)
        self.pointwise = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False)
        self.bn = nn.BatchNorm2d(out_channels, eps=1e-3, momentum=0.03)
        self.act = nn.SiLU(inplace=True)


class EfficientExpertGroup(nn.Module):
    """`EfficientExpertGroup(in_channels, out_channels, kernel_size, stride=1)`."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1):
        super().__init__()
        self.conv = DepthwiseSeparableConv(in_channels, out_channels, kernel_size, stride)


class ES_MOE(nn.Module, PackCache):
    """`ES_MOE(in_channels, out_channels=None, num_experts=4, reduction=8, top_k=2, use_sparse_inference=True,
    dynamic_threshold=0.4, max_kernel_size=15, expert_kernel_sizes=None)` — eval sparse path only."""

    def __init__(self, in_channels, out_channels=None, num_experts=4, reduction=8, top_k=2, use_sparse_inference=True,
                 dynamic_threshold=0.4, max_kernel_size=15, expert_kernel_sizes=None):
        super().__init__()
        if in_channels < 1 or (out_channels is not None and out_channels < 1):
            raise ValueError("in_channels and out_channels must be positive")
        if num_experts < 1:
            raise ValueError(f"num_experts must be positive, got {num_experts}")
        if top_k is not None and not 1 <= top_k <= num_experts:
            raise ValueError(f"top_k must be in [1, {num_experts}], got {top_k}")
        if not 0.0 <= dynamic_threshold <= 1.0:
            raise ValueError(f"dynamic_threshold must be in [0, 1], got {dynamic_threshold}")
        if max_kernel_size < 3:
            raise ValueError(f"max_kernel_size must be at least 3, got {max_kernel_size}")
        max_kernel_size = int(max_kernel_size)
        if max_kernel_size % 2 == 0:
            max_kernel_size -= 1
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.num_experts, self.reduction = in_channels, out_channels, num_experts, reduction
        self.top_k = min(top_k, num_experts) if top_k is not None else num_experts
        self.use_top_k = top_k is not None
        self.use_sparse_inference, self.dynamic_threshold, self.max_kernel_size = use_sparse_inference, dynamic_threshold, max_kernel_size
        self.routing = DynamicRoutingLayer(in_channels, num_experts, reduction, top_k)
        if expert_kernel_sizes is not None:
            if len(expert_kernel_sizes) != num_experts:
                raise ValueError(f"expert_kernel_sizes must have {num_experts} entries, got {len(expert_kernel_sizes)}")
            ks = [min(int(k) - (1 if int(k) % 2 == 0 else 0), max_kernel_size) for k in expert_kernel_sizes]
        else:
            default = [3, 5, 7]
            ks = [min(k, max_kernel_size) for k in default[:num_experts]] if num_experts <= len(default) \
                else [min(3 + 2 * i, max_kernel_size) for i in range(num_experts)]
        self.experts = nn.ModuleList([EfficientExpertGroup(in_channels, out_channels, kernel_size=k) for k in ks])
        self.norm = nn.Sequential(nn.BatchNorm2d(out_channels, eps=1e-3, momentum=0.03), nn.SiLU(inplace=True))
        self.register_buffer("load_balancing_loss", torch.tensor(0.0), persistent=False)
        self.register_buffer("expert_usage_counts", torch.zeros(num_experts), persistent=False)
        self.last_routing_snapshot = {}
        self.balance_loss_coeff = 1.0

    def _eager_sparse_enabled(self):
        return bool(self.use_sparse_inference and self.use_top_k and self.top_k < self.num_experts)

    def _pack_sources(self):
        return [t for t in list(self.parameters()) + list(self.buffers()) if t is not self.load_balancing_loss and t is not self.expert_usage_counts]

    def _build_pack(self):
        ks = [e.conv.depthwise.kernel_size[0] for e in self.experts]
        if any(k not in (3, 5, 7, 9) for k in ks):
            raise NotImplementedError(f"ES_MOE: expert kernel sizes {ks} not on the B200 path (3/5/7/9)")
        if self.in_channels % 16:
            raise NotImplementedError("ES_MOE: in_channels must be a multiple of 16 on the B200 path")
        r0, r2 = self.routing.routing_network[0], self.routing.routing_network[2]
        pw, pb = [], []
        for e in self.experts:
            w, b = fold_bn(e.conv.pointwise.weight, None, e.conv.bn)
            pw.append(pack_gemm_weight(w))
            pb.append(b)
        fs, fh = bn_affine(self.norm[0])
        C = self.in_channels
        return {
            "E": self.num_experts, "N": self.out_channels, "Cr": r0.weight.shape[0], "ks": ks,
            "rw1": r0.weight.detach().float().reshape(r0.weight.shape[0], C).contiguous(), "rb1": r0.bias.detach().float().contiguous(),
            "rw2": r2.weight.detach().float().reshape(self.num_experts, -1).contiguous(), "rb2": r2.bias.detach().float().contiguous(),
            "dw": [e.conv.depthwise.weight.detach().float().reshape(C, k * k).t().contiguous().half() for e, k in zip(self.experts, ks)],
            "pw": torch.stack(pw).contiguous(), "pb": torch.stack(pb).contiguous(), "fscale": fs, "fshift": fh,
        }

    def fwd_nhwc(self, x, out=None):
        require_eval(self)
        if not self._eager_sparse_enabled():
            raise NotImplementedError("ES_MOE: only the eval sparse path (top_k < num_experts, use_sparse_inference) is on the B200 path")
        y, idx, w, probs = ops.esmoe_forward(x, self.get_pack(), self.top_k, self.dynamic_threshold, out=out)
        self.last_routing_snapshot = {"topk_indices": idx, "topk_weights": w, "router_probs": probs}
        return y

    def forward(self, x):
        return to_nchw(self.fwd_nhwc(to_nhwc(x)))

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.norm[0].weight.device)
