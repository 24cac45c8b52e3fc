"""CPU oracle for the YOLO-Master detection forward pass.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 *restatement* of the reference's forward pass
(`/root/reference/ultralytics/nn/...`, cited per function as file:line).  It is
a checker: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
/ `--impl reference` legs may import it.  The product package never does.

Parity pinning: `tests/golden/make_golden.py` imports the real reference from
`/root/reference` in the build container, fills its parameters with the
deterministic generator in `yolo-master_b200/utils/synth.py`, and stores reference outputs in
`tests/golden/*.pt`; `tests/test_oracle_golden.py` checks this file against them.

Design: purely functional.  A model is (layer spec list, flat state_dict with
the reference's *unfused* key names, e.g. `model.4.m.0.0.mlp.experts.2.conv.0.weight`).
No nn.Module classes, no code shared with the product package.
"""
from __future__ import annotations

import math
from typing import Any

import torch
import torch.nn.functional as F

import contextlib

_SIM_FP16 = False  # see fp16_storage()


@contextlib.contextmanager
def fp16_storage():
    """Noise-floor model: same fp32 arithmetic, but every stored activation (conv / attention / expert outputs) is
    rounded to fp16, as ANY fp16 execution of this graph (the reference's `.half()` path included) must do.  Tests use
    |oracle_fp16_storage - oracle| as the yardstick for errors that accumulate across chained layers."""
    global _SIM_FP16
    old, _SIM_FP16 = _SIM_FP16, True
    try:
        yield
    finally:
        _SIM_FP16 = old


def _st(t):
    return t.half().float() if _SIM_FP16 else t


_W16 = False  # see fp16_weights()


@contextlib.contextmanager
def fp16_weights():
    """Model of the reference's deployed fp16 weights: `model.fuse().half()` folds BatchNorm into the conv weights
    (utils/torch_utils.py:315-349) and rounds them to fp16 (nn/backends/pytorch.py:44-67).  Arithmetic stays fp32.
    Used as the yardstick for single-kernel tests on activations with a large dynamic range, where fp16 weight
    rounding (common to the reference's fp16 path and the CUDA path) is visible against an fp32-weight oracle."""
    global _W16
    old, _W16 = _W16, True
    try:
        yield
    finally:
        _W16 = old


def _w(t):
    return t.half().float() if _W16 else t


BN_EPS = 1e-3  # utils/torch_utils.py:552-562 (`initialize_weights`, called by DetectionModel tasks.py:565) rewrites eps on every
#                nn.BatchNorm2d; ClassificationModel (tasks.py:947-964) never calls it, so its BatchNorms keep torch's default 1e-5:
#                `forward` switches this module constant for the duration of a classification model's pass
GN_EPS = 1e-5  # nn.GroupNorm default, untouched by initialize_weights


# ----------------------------------------------------------------------------
# YAML -> layer spec (restates nn/tasks.py:2022-2275 parse_model and
# nn/mixture_registry.py:84-156 adapt_mixture_args for the modules on the path)
# ----------------------------------------------------------------------------
def make_divisible(x: float, divisor: int) -> int:
    """utils/ops.py make_divisible."""
    return int(math.ceil(x / divisor) * divisor)


_BASE = {"Conv", "C2f", "C3k2", "SPPF", "C2PSA", "A2C2f", "DWConv", "C3", "Bottleneck", "Classify"}
_REPEAT = {"C2f", "C3k2", "C2PSA", "A2C2f", "C3"}
_MIX_BASE = {"A2C2fMoE", "ES_MOE"}
_MIX_REPEAT = {"A2C2fMoE"}


def parse_spec(d: dict, ch: int = 3, scale: str | None = None) -> dict:
    """Return {'layers': [{'i','f','type','args','n'}], 'save': [...], 'nc','reg_max','end2end'}."""
    nc, scales, end2end = d.get("nc"), d.get("scales"), d.get("end2end")
    reg_max = d.get("reg_max", 16)
    depth, width, max_channels = d.get("depth_multiple", 1.0), d.get("width_multiple", 1.0), float("inf")
    scale = scale or d.get("scale")
    if scales:
        if not scale:
            scale = next(iter(scales.keys()))
        depth, width, max_channels = scales[scale]
    chs = [ch]
    layers, save = [], []
    legacy = True
    for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
        args = list(args)
        for j, a in enumerate(args):
            if isinstance(a, str):
                if a == "nc":
                    args[j] = nc
                elif a == "None":
                    args[j] = None
        n = max(round(n * depth), 1) if n > 1 else n
        if m in _BASE:
            c1, c2 = chs[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [c1, c2, *args[1:]]
            if m in _REPEAT:
                args.insert(2, n)
                n = 1
            if m == "C3k2":
                legacy = False
                if scale and scale in "mlx":
                    args[3] = True
            if m == "A2C2f":
                legacy = False
                if scale and scale in "lx":
                    args.extend((True, 1.2))
        elif m in _MIX_BASE:
            c1, c2 = chs[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [c1, c2, *args[1:]]
            if m in _MIX_REPEAT:
                args.insert(2, n)
                n = 1
            if m == "A2C2fMoE":
                legacy = False
        elif m == "Concat":
            c2 = sum(chs[x] for x in f)
        elif m == "LatentMixture":     # multi-input mixture module (mixture_registry.py:96-141)
            c2 = args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [[chs[x] for x in f], c2, *args[1:]]
        elif m == "Detect":
            args = [args[0], reg_max, end2end, [chs[x] for x in f]]
            c2 = None
        elif m == "Pose":
            args = [args[0], tuple(d["kpt_shape"]), reg_max, end2end, [chs[x] for x in f]]
            c2 = None
        elif m == "OBB":
            args = [args[0], args[1], reg_max, end2end, [chs[x] for x in f]]
            c2 = None
        elif m == "Segment":
            args = [args[0], args[1], make_divisible(min(args[2], max_channels) * width, 8), reg_max, end2end, [chs[x] for x in f]]
            c2 = None
        else:  # nn.Upsample etc.
            c2 = chs[f]
        layers.append({"i": i, "f": f, "type": m, "args": args, "n": n, "legacy": legacy})
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        if i == 0:
            chs = []
        chs.append(c2)
    return {"layers": layers, "save": sorted(set(save)), "nc": nc, "reg_max": reg_max, "end2end": bool(end2end)}


# ----------------------------------------------------------------------------
# leaf ops
# ----------------------------------------------------------------------------
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def conv_block(sd, p, x, s=1, g=1, act=True, pad=None):
    """`Conv.forward` nn/modules/conv.py:69-78: act(bn(conv(x))), autopad :30-36."""
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    pad = k // 2 if pad is None else pad
    if _W16:  # fused + fp16-rounded weights, as deployed by the reference's fp16 backend
        sc = sd[p + ".bn.weight"] / torch.sqrt(sd[p + ".bn.running_var"] + BN_EPS)
        b0 = sd.get(p + ".conv.bias")
        b0 = torch.zeros_like(sc) if b0 is None else b0
        y = F.conv2d(x, _w(w * sc.view(-1, 1, 1, 1)), (b0 - sd[p + ".bn.running_mean"]) * sc + sd[p + ".bn.bias"], s, pad, 1, g)
    else:
        y = F.conv2d(x, w, sd.get(p + ".conv.bias"), s, pad, 1, g)
        y = _bn(sd, p + ".bn", y)
    return _st(F.silu(y) if act else y)


def dwconv_block(sd, p, x, act=True):
    """`DWConv` conv.py:185-199: groups = gcd(c1, c2)."""
    w = sd[p + ".conv.weight"]
    g = math.gcd(x.shape[1], w.shape[0])
    return conv_block(sd, p, x, 1, g, act)


def bottleneck(sd, p, x, shortcut=True, g=1):
    """`Bottleneck.forward` block.py:484-486 (k=(3,3); c1==c2 inside C3k2/C3k)."""
    y = conv_block(sd, p + ".cv2", conv_block(sd, p + ".cv1", x), 1, g)
    add = shortcut and x.shape[1] == y.shape[1]
    return _st(x + y) if add else y


def c3k(sd, p, x, n=2, shortcut=True, g=1):
    """`C3k`/`C3.forward` block.py:348-350,1114-1131."""
    a = conv_block(sd, p + ".cv1", x)
    for j in range(n):
        a = bottleneck(sd, f"{p}.m.{j}", a, shortcut, g)
    return conv_block(sd, p + ".cv3", torch.cat((a, conv_block(sd, p + ".cv2", x)), 1))


def attention(sd, p, x, num_heads, attn_ratio=0.5):
    """`Attention.forward` block.py:1313-1333."""
    B, C, H, W = x.shape
    N = H * W
    head_dim = C // num_heads
    key_dim = int(head_dim * attn_ratio)
    scale = key_dim ** -0.5
    qkv = conv_block(sd, p + ".qkv", x, act=False)
    q, k, v = qkv.view(B, num_heads, key_dim * 2 + head_dim, N).split([key_dim, key_dim, head_dim], dim=2)
    attn = (q * scale).transpose(-2, -1) @ k
    attn = attn.softmax(dim=-1)
    o = _st((v @ attn.transpose(-2, -1)).view(B, C, H, W)) + conv_block(sd, p + ".pe", v.reshape(B, C, H, W), 1, C, False)
    return conv_block(sd, p + ".proj", _st(o), act=False)


def psablock(sd, p, x, num_heads, shortcut=True):
    """`PSABlock.forward` block.py:1372-1383."""
    a = attention(sd, p + ".attn", x, num_heads)
    x = _st(x + a) if shortcut else a
    f = conv_block(sd, p + ".ffn.1", conv_block(sd, p + ".ffn.0", x), act=False)
    return _st(x + f) if shortcut else f


def aattn(sd, p, x, num_heads, area=1):
    """`AAttn.forward` block.py:1696-1732 (q scaled before q^T k; pe on V; head-interleaved qkv)."""
    B, C, H, W = x.shape
    N = H * W
    hd = C // num_heads
    qkv = conv_block(sd, p + ".qkv", x, act=False).flatten(2).transpose(1, 2)
    Bq = B
    if area > 1:
        qkv = qkv.reshape(B * area, N // area, C * 3)
        Bq, N = qkv.shape[0], qkv.shape[1]
    q, k, v = qkv.view(Bq, N, num_heads, hd * 3).permute(0, 2, 3, 1).split([hd, hd, hd], dim=2)
    attn = (q * (hd ** -0.5)).transpose(-2, -1) @ k
    attn = attn.softmax(dim=-1)
    o = _st(v @ attn.transpose(-2, -1)).permute(0, 3, 1, 2)
    v = v.permute(0, 3, 1, 2)
    if area > 1:
        o = o.reshape(B, N * area, C)
        v = v.reshape(B, N * area, C)
    o = o.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    v = v.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    o = _st(o + conv_block(sd, p + ".pe", v, 1, C, False, pad=3))
    return conv_block(sd, p + ".proj", o, act=False)


def get_safe_groups(channels: int, desired: int = 8) -> int:
    """nn/modules/utils.py:108-115."""
    if channels <= 0:
        return 1
    g = min(desired, channels)
    while channels % g != 0:
        g -= 1
    return max(1, g)


def simple_expert(sd, p, x):
    """`SimpleExpert` moe/experts.py:73-88: 1x1 -> GN -> SiLU -> 1x1 -> GN."""
    w1, w2 = sd[p + ".conv.0.weight"], sd[p + ".conv.3.weight"]
    h = _st(F.conv2d(x, _w(w1)))
    h = F.group_norm(h, get_safe_groups(w1.shape[0]), sd[p + ".conv.1.weight"], sd[p + ".conv.1.bias"], GN_EPS)
    h = F.silu(h)
    o = _st(F.conv2d(h, _w(w2)))
    return F.group_norm(o, get_safe_groups(w2.shape[0]), sd[p + ".conv.4.weight"], sd[p + ".conv.4.bias"], GN_EPS)


def efficient_spatial_router(sd, p, x, top_k, pool_scale=4):
    """`EfficientSpatialRouter.forward` moe/routers.py:283-304 + `_process_logits` :185-265 (eval).

    Returns (weights fp32 [B,k], indices int64 [B,k], probs fp32 [B,E]).
    """
    B, C, H, W = x.shape
    xin = F.avg_pool2d(x, pool_scale, pool_scale) if (H > pool_scale and W > pool_scale) else x
    h = F.conv2d(xin, sd[p + ".router.0.weight"], None, 1, 1)
    h = F.silu(_bn(sd, p + ".router.1", h))
    o = _bn(sd, p + ".router.4", F.conv2d(h, sd[p + ".router.3.weight"]))
    logits = o.float().mean(dim=[2, 3])
    probs = F.softmax(logits.float(), dim=1)
    vals, idx = torch.topk(probs, top_k, dim=1)
    vals = vals / vals.sum(dim=1, keepdim=True).clamp_min(1e-6)
    return vals, idx, probs


def optimized_moe_improved(sd, p, x, num_experts, top_k):
    """`OptimizedMOEImproved.forward` moe/modules.py:1069-1157 (eval; add_residual=False under ABlockMoE)."""
    B = x.shape[0]
    w, idx, _ = efficient_spatial_router(sd, p + ".routing", x, top_k)
    if _W16:
        sc = sd[p + ".shared_expert.1.weight"] / torch.sqrt(sd[p + ".shared_expert.1.running_var"] + BN_EPS)
        shared = F.silu(F.conv2d(x, _w(sd[p + ".shared_expert.0.weight"] * sc.view(-1, 1, 1, 1)),
                                 sd[p + ".shared_expert.1.bias"] - sd[p + ".shared_expert.1.running_mean"] * sc))
    else:
        shared = F.silu(_bn(sd, p + ".shared_expert.1", F.conv2d(x, sd[p + ".shared_expert.0.weight"])))
    out = torch.zeros_like(shared, dtype=torch.float32)
    for e in range(num_experts):
        mask = idx == e
        if mask.any():
            bi, ki = torch.where(mask)
            o = simple_expert(sd, f"{p}.experts.{e}", x[bi])
            out.index_add_(0, bi, o.float() * w[bi, ki].view(-1, 1, 1, 1))
    return _st((shared.float() + out).to(x.dtype))


def _conv1x1_gn(sd, p, x):
    """`_conv1x1` latent_mixture.py:48-53: 1x1 -> GroupNorm(1) -> SiLU."""
    return _st(F.silu(_gn(sd, p + ".1", F.conv2d(x, _w(sd[p + ".0.weight"])), 1)))


def latent_router(sd, p, tokens, temperature):
    """`LatentRouter.forward` latent_mixture.py:219-241 (per_token = False, fp32): mean over tokens of (token + scale embedding) ->
    LayerNorm -> Linear + SiLU -> Linear + SiLU -> expert head -> nan_to_num / clamp(+-30) -> softmax(/ max(T, 0.1))."""
    x = tokens.float()
    if p + ".scale_embedding" in sd:
        x = x + sd[p + ".scale_embedding"].unsqueeze(0)
    r = F.layer_norm(x.mean(1), (x.shape[-1],), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)
    h = F.silu(F.linear(r, sd[p + ".trunk.0.weight"], sd[p + ".trunk.0.bias"]))
    h = F.silu(F.linear(h, sd[p + ".trunk.2.weight"], sd[p + ".trunk.2.bias"]))
    logits = torch.nan_to_num(F.linear(h, sd[p + ".expert_head.weight"], sd[p + ".expert_head.bias"]), nan=0.0, posinf=30.0, neginf=-30.0).clamp(-30.0, 30.0)
    return logits, F.softmax(logits / max(float(temperature), 0.1), dim=-1)


def layer_latent_mixture(sd, p, xs, in_channels, c2, num_experts=4, expert_ratio=0.25, router_hidden_dim=None, temperature=1.0, *unused,
                         return_route=False):
    """`LatentMixture.forward` latent_mixture.py:721-800 (eval, dense dispatch, router_only fusion): projected tokens -> router ->
    base + residual_gain * sum_e probs[:, e] * DenseChannelExpert_e(base)."""
    proj = [x if c == c2 else _conv1x1_gn(sd, f"{p}.token_projs.{i}", x) for i, (x, c) in enumerate(zip(xs, in_channels))]
    base = xs[0] if in_channels[0] == c2 else _conv1x1_gn(sd, p + ".base_proj", xs[0])
    tokens = torch.stack([_st(t.mean((2, 3))).float() for t in proj], 1)
    T = float(sd[p + ".router._temperature"]) if p + ".router._temperature" in sd else float(temperature)
    logits, probs = latent_router(sd, p + ".router", tokens, T)
    mixed = torch.zeros_like(base)
    for e in range(num_experts):
        q = f"{p}.experts.{e}.net"
        hid = sd[q + ".0.weight"].shape[0]
        t = _st(F.silu(_gn(sd, q + ".1", F.conv2d(base, _w(sd[q + ".0.weight"])), 1)))
        t = _st(F.silu(_gn(sd, q + ".4", F.conv2d(t, _w(sd[q + ".3.weight"]), None, 1, 1, 1, hid), 1)))
        mixed = mixed + _st(F.conv2d(t, _w(sd[q + ".6.weight"]))) * probs[:, e].view(-1, 1, 1, 1)
    out = _st(base + sd[p + ".residual_gain"] * mixed)
    return (out, probs, logits) if return_route else out


def ultra_efficient_router(sd, p, x, top_k, temperature=1.0, pool_scale=8):
    """`UltraEfficientRouter.forward` moe/routers.py:96-118 (eval): pooled local stream, PER-PIXEL clamp / T / softmax over experts, spatial
    mean, top-k, renormalise (clamp_min 1e-6).  Returns (weights [B,k], indices [B,k], pooled probabilities [B,E])."""
    B, C, H, W = x.shape
    xl = F.avg_pool2d(x, pool_scale, pool_scale) if (H > pool_scale and W > pool_scale) else x
    t = F.conv2d(xl, sd[p + ".router.0.weight"], None, 1, 1, 1, C)
    t = F.silu(_gn(sd, p + ".router.1", t, get_safe_groups(C, 8)))
    t = F.conv2d(t, sd[p + ".router.3.weight"])
    t = F.silu(_gn(sd, p + ".router.4", t, get_safe_groups(t.shape[1], 4)))
    logits = F.conv2d(t, sd[p + ".router.6.weight"], sd[p + ".router.6.bias"])
    wts = F.softmax((logits.clamp(-30.0, 30.0) / temperature).float(), dim=1)
    pooled = wts.mean((2, 3))
    w, idx = torch.topk(pooled, top_k, dim=1)
    return w / w.sum(1, keepdim=True).clamp_min(1e-6), idx, pooled


def layer_ultra_optimized_moe(sd, p, x, c1, c2, num_experts=4, top_k=2, *unused, return_route=False):
    """`UltraOptimizedMoE.forward` moe/modules.py:205-214 (v0_1 uomoe / v0_2 zoos), eval: UltraEfficientRouter, shared expert
    (1x1 -> GroupNorm -> SiLU) + `BatchedExpertComputation.compute_sparse_experts_batched` over OptimizedSimpleExpert (the SimpleExpert
    structure: routes with weight <= 0.01 dropped, fp32 multiply, clamp +-1e4; moe/utils.py:119-209).  No residual."""
    w, idx, probs = ultra_efficient_router(sd, p + ".routing", x, top_k)
    shared = F.silu(_gn(sd, p + ".shared_expert.1", F.conv2d(x, _w(sd[p + ".shared_expert.0.weight"])), get_safe_groups(c2, 8)))
    out = torch.zeros_like(shared)
    valid = w > 0.01
    for e in range(num_experts):
        mask = (idx == e) & valid
        if mask.any():
            bi, ki = torch.where(mask)
            out.index_add_(0, bi, simple_expert(sd, f"{p}.experts.{e}", x[bi]).float() * w[bi, ki].view(-1, 1, 1, 1))
    y = _st(shared + out.clamp(-1e4, 1e4))
    return (y, w, idx, probs) if return_route else y


def layer_modular_router_expert_moe(sd, p, x, c1, c2, num_experts=4, top_k=2, *unused):
    """`ModularRouterExpertMoE` (= `OptimizedMOEImproved`, moe/modules.py:1745) as a top-level YAML layer (v0_1 zoo): the block
    owns its residual, `final_output + x` when in == out channels (modules.py:1156-1159)."""
    y = optimized_moe_improved(sd, p, x, num_experts, top_k)
    return _st(y + x) if c1 == c2 else y


def es_moe_kernel_sizes(num_experts, max_kernel_size=15, expert_kernel_sizes=None):
    """`ES_MOE.__init__` moe/modules.py:478-493."""
    if max_kernel_size % 2 == 0:
        max_kernel_size -= 1
    if expert_kernel_sizes is not None:
        return [min(int(k) - (1 if int(k) % 2 == 0 else 0), max_kernel_size) for k in expert_kernel_sizes]
    default = [3, 5, 7]
    if num_experts <= len(default):
        return [min(k, max_kernel_size) for k in default[:num_experts]]
    return [min(3 + 2 * i, max_kernel_size) for i in range(num_experts)]


def dynamic_routing_hard_topk(sd, p, x, top_k):
    """`DynamicRoutingLayer.forward` + `_hard_top_k` moe/routers.py:458-496,519-527 (eval, Top-K enabled).

    Returns per-image weights (B, E) fp32 (zero for unselected experts); they are spatially constant (:496)."""
    pooled = x.mean(dim=(2, 3), keepdim=True)
    h = F.silu(F.conv2d(pooled, sd[p + ".routing_network.0.weight"], sd[p + ".routing_network.0.bias"]))
    logits = F.conv2d(h, sd[p + ".routing_network.2.weight"], sd[p + ".routing_network.2.bias"])
    B, E = logits.shape[:2]
    w = F.softmax(logits.reshape(B, E).float().clamp(-30.0, 30.0), dim=1)
    vals, idx = torch.topk(w, top_k, dim=1)
    vals = vals / vals.sum(dim=1, keepdim=True).clamp_min(1e-6)        # stable_normalize _numeric.py:85-90
    return torch.zeros_like(w).scatter_(1, idx, vals)


def es_moe_retained_weights(rw, top_k, dynamic_threshold):
    """Sample-level part of `ES_MOE._sparse_forward` moe/modules.py:659-684: Top-K over the (spatially constant) importance,
    rank >= 1 experts dropped below `dynamic_threshold`, retained mass renormalised.  rw (B, E) -> (indices (B,k), w (B,E),
    retained mask (B,E)).  Known answer of the reference (tests/test_moe.py:409-420): (0.6, 0.4, 0), k=2, thr 0.5 -> (1, 0, 0)."""
    B, E = rw.shape
    tv, ti = torch.topk(rw, top_k, dim=1)
    ranks = torch.arange(top_k, device=rw.device).view(1, -1)
    keep = (ranks == 0) | (tv >= dynamic_threshold) if dynamic_threshold > 0 else torch.ones_like(ti, dtype=torch.bool)
    retained = torch.zeros(B, E, dtype=torch.bool, device=rw.device).scatter_(1, ti, keep)
    w = rw * retained
    w = w / w.sum(1, keepdim=True).clamp_min(torch.finfo(torch.float32).eps)
    return ti, w, retained


def es_moe(sd, p, x, c1, c2=None, num_experts=4, reduction=8, top_k=2, use_sparse_inference=True, dynamic_threshold=0.4,
           max_kernel_size=15, expert_kernel_sizes=None):
    """`ES_MOE.forward` (eval) moe/modules.py:535-583 with `_sparse_forward` :659-704: sample-level Top-K with the rank>=1
    experts dropped when their renormalised weight < dynamic_threshold, expert = dw kxk -> 1x1 -> BN -> SiLU
    (experts.py:280-296), fp32 here; final BN + SiLU (:496,581)."""
    assert use_sparse_inference and top_k is not None and top_k < num_experts, "dense ES_MOE paths are not on the oracle"
    c2 = c1 if c2 is None else c2
    rw = dynamic_routing_hard_topk(sd, p + ".routing", x, top_k)                   # (B, E)
    B, E = rw.shape
    ti, w, retained = es_moe_retained_weights(rw, top_k, dynamic_threshold)
    ks = es_moe_kernel_sizes(num_experts, max_kernel_size, expert_kernel_sizes)
    out = torch.zeros(B, c2, x.shape[2], x.shape[3])
    for e in range(E):
        bi = torch.where(retained[:, e])[0]
        if bi.numel() == 0:
            continue
        q = f"{p}.experts.{e}.conv"
        t = F.conv2d(x[bi], _w(sd[q + ".depthwise.weight"]), None, 1, (ks[e] - 1) // 2, 1, x.shape[1])
        t = _st(t)
        if _W16:
            sc = sd[q + ".bn.weight"] / torch.sqrt(sd[q + ".bn.running_var"] + BN_EPS)
            y = F.conv2d(t, _w(sd[q + ".pointwise.weight"] * sc.view(-1, 1, 1, 1)), sd[q + ".bn.bias"] - sd[q + ".bn.running_mean"] * sc)
        else:
            y = _bn(sd, q + ".bn", F.conv2d(t, sd[q + ".pointwise.weight"]))
        y = _st(F.silu(y))
        out.index_add_(0, bi, y * w[bi, e].view(-1, 1, 1, 1))
    return _st(F.silu(_bn(sd, p + ".norm.0", _st(out)))), (ti, w)


def layer_es_moe(sd, p, x, *args):
    return es_moe(sd, p, x, *args)[0]


def ablock_moe(sd, p, x, num_heads, area, num_experts, top_k):
    """`ABlockMoE.forward` moe/modules.py:1247-1260."""
    x = _st(x + aattn(sd, p + ".attn", x, num_heads, area))
    return _st(x + optimized_moe_improved(sd, p + ".mlp", x, num_experts, top_k))


# ----------------------------------------------------------------------------
# top-level layers
# ----------------------------------------------------------------------------
def layer_conv(sd, p, x, c1, c2, k=1, s=1, pad=None, g=1, d=1, act=True):
    return conv_block(sd, p, x, s, g, act is True, pad)


def layer_c3k2(sd, p, x, c1, c2, n=1, c3k_=False, e=0.5, attn=False, g=1, shortcut=True):
    """`C3k2` block.py:1074-1108 on top of `C2f.forward` :313-317."""
    c = int(c2 * e)
    y = list(conv_block(sd, p + ".cv1", x).chunk(2, 1))
    for j in range(n):
        q = f"{p}.m.{j}"
        if attn:
            t = bottleneck(sd, q + ".0", y[-1], shortcut, g)
            t = psablock(sd, q + ".1", t, max(c // 64, 1))
        elif c3k_:
            t = c3k(sd, q, y[-1], 2, shortcut, g)
        else:
            t = bottleneck(sd, q, y[-1], shortcut, g)
        y.append(t)
    return conv_block(sd, p + ".cv2", torch.cat(y, 1))


def layer_c2f(sd, p, x, c1, c2, n=1, shortcut=False, g=1, e=0.5):
    """`C2f.forward` block.py:313-317."""
    y = list(conv_block(sd, p + ".cv1", x).chunk(2, 1))
    for j in range(n):
        y.append(bottleneck(sd, f"{p}.m.{j}", y[-1], shortcut, g))
    return conv_block(sd, p + ".cv2", torch.cat(y, 1))


def layer_sppf(sd, p, x, c1, c2, k=5, n=3, shortcut=False):
    """`SPPF.forward` block.py:237-242 (cv1 has act=False)."""
    y = [conv_block(sd, p + ".cv1", x, act=False)]
    for _ in range(n):
        y.append(F.max_pool2d(y[-1], k, 1, k // 2))
    o = conv_block(sd, p + ".cv2", torch.cat(y, 1))
    return o + x if (shortcut and c1 == c2) else o


def layer_c2psa(sd, p, x, c1, c2, n=1, e=0.5):
    """`C2PSA.forward` block.py:1482-1493."""
    c = int(c1 * e)
    a, b = conv_block(sd, p + ".cv1", x).split((c, c), 1)
    for j in range(n):
        b = psablock(sd, f"{p}.m.{j}", b, c // 64)
    return conv_block(sd, p + ".cv2", torch.cat((a, b), 1))


def layer_a2c2f_moe(sd, p, x, c1, c2, n=1, a2=True, area=1, residual=False, mlp_ratio=2.0, e=0.5, g=1,
                    shortcut=True, num_experts=4, top_k=2, expert_type="simple"):
    """`A2C2fMoE` moe/modules.py:1268-1297 over `A2C2f.forward` block.py:1865-1879."""
    c_ = int(c2 * e)
    y = [conv_block(sd, p + ".cv1", x)]
    for j in range(n):
        t = y[-1]
        if a2:
            for r in range(2):
                t = ablock_moe(sd, f"{p}.m.{j}.{r}", t, c_ // 32, area, num_experts, top_k)
        else:
            t = c3k(sd, f"{p}.m.{j}", t, 2, shortcut, g)
        y.append(t)
    o = conv_block(sd, p + ".cv2", torch.cat(y, 1))
    if a2 and residual:
        return x + sd[p + ".gamma"].view(1, -1, 1, 1) * o
    return o


def ablock(sd, p, x, num_heads, area):
    """`ABlock.forward` block.py:1787-1797: x + attn(x); x + mlp(x), mlp = Conv(c, c*r, 1) -> Conv(c*r, c, 1, act=False)."""
    x = _st(x + aattn(sd, p + ".attn", x, num_heads, area))
    return _st(x + conv_block(sd, p + ".mlp.1", conv_block(sd, p + ".mlp.0", x), act=False))


def layer_a2c2f(sd, p, x, c1, c2, n=1, a2=True, area=1, residual=False, mlp_ratio=2.0, e=0.5, g=1, shortcut=True):
    """`A2C2f.forward` block.py:1865-1879 (n x Sequential(ABlock, ABlock) or C3k; optional layer-scale residual)."""
    c_ = int(c2 * e)
    y = [conv_block(sd, p + ".cv1", x)]
    for j in range(n):
        t = y[-1]
        if a2:
            for r in range(2):
                t = ablock(sd, f"{p}.m.{j}.{r}", t, c_ // 32, area)
        else:
            t = c3k(sd, f"{p}.m.{j}", t, 2, shortcut, g)
        y.append(t)
    o = conv_block(sd, p + ".cv2", torch.cat(y, 1))
    if a2 and residual:
        return _st(x + sd[p + ".gamma"].view(1, -1, 1, 1) * o)
    return o


def make_anchors(shapes, strides, offset=0.5, device=None, dtype=torch.float32):
    """utils/tal.py:398-411 (anchors in the feature dtype/device, as the reference builds them)."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, dtype=dtype, device=device) + offset
        sy = torch.arange(h, dtype=dtype, device=device) + offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=dtype, device=device))
    return torch.cat(pts), torch.cat(st)


def detect_head_raw(sd, p, feats, nc, reg_max, end2end, legacy=False):
    """`Detect.forward_head` head.py:146-155 on the one2one (end2end) or one2many towers."""
    bs = feats[0].shape[0]
    bp = p + (".one2one_cv2" if end2end else ".cv2")
    cp = p + (".one2one_cv3" if end2end else ".cv3")
    boxes, scores = [], []
    for i, x in enumerate(feats):
        b = conv_block(sd, f"{bp}.{i}.1", conv_block(sd, f"{bp}.{i}.0", x))
        b = F.conv2d(b, _w(sd[f"{bp}.{i}.2.weight"]), sd[f"{bp}.{i}.2.bias"])
        if legacy:
            c = conv_block(sd, f"{cp}.{i}.1", conv_block(sd, f"{cp}.{i}.0", x))
        else:
            c = conv_block(sd, f"{cp}.{i}.0.1", dwconv_block(sd, f"{cp}.{i}.0.0", x))
            c = conv_block(sd, f"{cp}.{i}.1.1", dwconv_block(sd, f"{cp}.{i}.1.0", c))
        c = F.conv2d(c, _w(sd[f"{cp}.{i}.2.weight"]), sd[f"{cp}.{i}.2.bias"])
        boxes.append(b.view(bs, 4 * reg_max, -1))
        scores.append(c.view(bs, nc, -1))
    return torch.cat(boxes, -1), torch.cat(scores, -1)


def pose_kpts(sd, p, feats, kpt_shape, shapes, strides):
    """`Pose.forward_head` head.py:613-621 (one2many keypoint towers) + `kpts_decode` :644-664.  Returns (B, nk, A)."""
    bs, nk, ndim = feats[0].shape[0], kpt_shape[0] * kpt_shape[1], kpt_shape[1]
    raw = []
    for i, x in enumerate(feats):
        t = conv_block(sd, f"{p}.cv4.{i}.1", conv_block(sd, f"{p}.cv4.{i}.0", x))
        raw.append(F.conv2d(t, _w(sd[f"{p}.cv4.{i}.2.weight"]), sd[f"{p}.cv4.{i}.2.bias"]).view(bs, nk, -1))
    y = torch.cat(raw, 2).clone()
    anchors, st = make_anchors(shapes, strides, device=y.device, dtype=y.dtype)
    anchors, st = anchors.t(), st.t()
    if ndim == 3:
        y[:, 2::ndim] = y[:, 2::ndim].sigmoid()
    y[:, 0::ndim] = (y[:, 0::ndim] * 2.0 + (anchors[0] - 0.5)) * st
    y[:, 1::ndim] = (y[:, 1::ndim] * 2.0 + (anchors[1] - 0.5)) * st
    return y


def segment_extras(sd, p, feats, nm):
    """`Segment.forward_head` head.py:337-345 (one2many mask-coefficient towers) and `Proto.forward` block.py:105-107.
    Returns (mask coefficients (B, nm, A), prototypes (B, nm, 2h, 2w))."""
    bs = feats[0].shape[0]
    mc = []
    for i, x in enumerate(feats):
        t = conv_block(sd, f"{p}.cv4.{i}.1", conv_block(sd, f"{p}.cv4.{i}.0", x))
        mc.append(F.conv2d(t, _w(sd[f"{p}.cv4.{i}.2.weight"]), sd[f"{p}.cv4.{i}.2.bias"]).view(bs, nm, -1))
    t = conv_block(sd, p + ".proto.cv1", feats[0])
    t = _st(F.conv_transpose2d(t, _w(sd[p + ".proto.upsample.weight"]), sd[p + ".proto.upsample.bias"], 2, 0))
    proto = conv_block(sd, p + ".proto.cv3", conv_block(sd, p + ".proto.cv2", t))
    return torch.cat(mc, 2), proto


def obb_decode(sd, p, feats, boxes, scores, shapes, strides, reg_max):
    """`OBB.forward_head` head.py:484-496 (angle towers, (sigmoid - 0.25) * pi), `OBB.decode_bboxes` = `dist2rbox` utils/tal.py:447-453 and
    `OBB._inference` head.py:477-482.  Returns (B, 4 + nc + 1, A): rotated cx, cy, w, h in pixels, sigmoid scores, angle."""
    bs = feats[0].shape[0]
    ang = []
    for i, x in enumerate(feats):
        t = conv_block(sd, f"{p}.cv4.{i}.1", conv_block(sd, f"{p}.cv4.{i}.0", x))
        ang.append(F.conv2d(t, _w(sd[f"{p}.cv4.{i}.2.weight"]), sd[f"{p}.cv4.{i}.2.bias"]).view(bs, 1, -1))
    angle = (torch.cat(ang, 2).sigmoid() - 0.25) * math.pi
    if reg_max > 1:
        b, _, a = boxes.shape
        boxes = (boxes.view(b, 4, reg_max, a).transpose(2, 1).softmax(1)
                 * torch.arange(reg_max, dtype=boxes.dtype).view(1, reg_max, 1, 1)).sum(1)
    anchors, st = make_anchors(shapes, strides, dtype=boxes.dtype)
    anchors, st = anchors.t().unsqueeze(0), st.t()
    lt, rb = boxes.split(2, 1)
    cos, sin = torch.cos(angle), torch.sin(angle)
    xf, yf = ((rb - lt) / 2).split(1, 1)
    xy = torch.cat([xf * cos - yf * sin, xf * sin + yf * cos], 1) + anchors
    dbox = torch.cat([xy, lt + rb], 1) * st
    return torch.cat([dbox, scores.sigmoid(), angle], 1)


def detect_decode(boxes, scores, shapes, strides, end2end, reg_max=1):
    """`Detect._inference` head.py:173-194, `decode_bboxes` :210-217, `dist2bbox` tal.py:414-423.

    Returns y (B, 4+nc, A): xyxy (end2end) or xywh boxes in pixels + sigmoid scores.
    """
    if reg_max > 1:  # `DFL.forward` block.py:80-85 (frozen arange weights; sd carries them as `dfl.conv.weight`)
        b, _, a = boxes.shape
        boxes = (boxes.view(b, 4, reg_max, a).transpose(2, 1).softmax(1)
                 * torch.arange(reg_max, dtype=boxes.dtype, device=boxes.device).view(1, reg_max, 1, 1)).sum(1)
    anchors, st = make_anchors(shapes, strides, device=boxes.device, dtype=boxes.dtype)
    anchors, st = anchors.t().unsqueeze(0), st.t()
    lt, rb = boxes.chunk(2, 1)
    x1y1, x2y2 = anchors - lt, anchors + rb
    if end2end:
        dbox = torch.cat((x1y1, x2y2), 1)
    else:
        dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1)
    return torch.cat((dbox * st, scores.sigmoid()), 1)


def detect_postprocess(y, nc, max_det=300):
    """`Detect.postprocess` head.py:219-233 + `get_topk_index` :235-258 (non-agnostic)."""
    preds = y.permute(0, 2, 1)
    boxes, scores = preds.split([4, nc], dim=-1)
    B, A, _ = scores.shape
    k = min(max_det, A)
    ori = scores.max(dim=-1)[0].topk(k)[1].unsqueeze(-1)
    sc = scores.gather(1, ori.repeat(1, 1, nc))
    sc, index = sc.flatten(1).topk(k)
    idx = ori[torch.arange(B, device=y.device)[..., None], index // nc]
    boxes = boxes.gather(1, idx.repeat(1, 1, 4))
    return torch.cat([boxes, sc[..., None], (index % nc)[..., None].float()], dim=-1), idx.squeeze(-1)


# ----------------------------------------------------------------------------
# whole model (`BaseModel._predict_once` tasks.py:182-218)
# ----------------------------------------------------------------------------
_LAYER_FN = {"ES_MOE": layer_es_moe, "Conv": layer_conv, "C3k2": layer_c3k2, "C2f": layer_c2f, "SPPF": layer_sppf, "C2PSA": layer_c2psa,
             "A2C2fMoE": layer_a2c2f_moe, "A2C2f": layer_a2c2f}


def forward_layer(spec: dict, sd: dict, i: int, xin):
    """Run ONE top-level layer (not Detect) on given input(s): used for teacher-forced per-layer parity."""
    L = spec["layers"][i]
    t, args = L["type"], L["args"]
    if t in _LAYER_FN:
        return _LAYER_FN[t](sd, f"model.{i}", xin, *args)
    if t == "nn.Upsample":
        return F.interpolate(xin, scale_factor=float(args[1]), mode=args[2])
    if t == "Concat":
        return torch.cat(xin, args[0] if args else 1)
    raise NotImplementedError(t)


def forward(spec: dict, sd: dict, x: torch.Tensor, img_hw=None, return_layers: bool = False,
            end2end: bool | None = None, dtype: torch.dtype = torch.float32) -> Any:
    """Eval forward.  Returns (B,300,6) for end2end models, else (B,4+nc,A).

    dtype=torch.float16 with CUDA tensors reproduces the reference's torch-eager `.half().cuda()` path (the yolo26-master-n
    layer types only); bench.py times that as the same-GPU baseline.  The parity oracle is the fp32 default."""
    global BN_EPS
    if spec["layers"][-1]["type"] == "Classify" and BN_EPS != 1e-5:
        BN_EPS = 1e-5
        try:
            return forward(spec, sd, x, img_hw, return_layers, end2end, dtype)
        finally:
            BN_EPS = 1e-3
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    x = x.to(dtype)
    H_in = x.shape[-2]
    outs, ys = [], {}
    for L in spec["layers"]:
        i, f, t, args = L["i"], L["f"], L["type"], L["args"]
        if isinstance(f, int):
            xin = x if f == -1 else ys[f % i if f < 0 else f]
        else:
            xin = [x if j == -1 else ys[j] for j in f]
        p = f"model.{i}"
        if t in _LAYER_FN:
            x = _LAYER_FN[t](sd, p, xin, *args)
        elif t == "nn.Upsample":
            x = F.interpolate(xin, scale_factor=float(args[1]), mode=args[2])
        elif t == "Concat":
            x = torch.cat(xin, args[0] if args else 1)
        elif t == "Detect":
            nc, reg_max, e2e, _ = args
            e2e = bool(e2e) if end2end is None else end2end
            shapes = [tuple(v.shape[2:]) for v in xin]
            strides = [H_in / s[0] for s in shapes]
            braw, sraw = detect_head_raw(sd, p, xin, nc, reg_max, e2e, L.get("legacy", False))
            y = detect_decode(braw, sraw, shapes, strides, e2e, reg_max)
            x = detect_postprocess(y, nc)[0] if e2e else y
            ys["detect_raw"] = (braw, sraw, y)
        elif t == "Classify":  # head.py:825-832: Conv -> global average pool -> Linear -> softmax; the model output is the probabilities
            c1, c2 = args[0], args[1]
            t_ = conv_block(sd, p + ".conv", xin)
            logits = F.linear(_st(t_.mean((2, 3))), _w(sd[p + ".linear.weight"]), sd[p + ".linear.bias"])
            x = logits.softmax(1)
            ys["logits"] = logits
        elif t == "OBB":       # rotated boxes + angle row (head.py:477-500); one2many head
            nc, ne, reg_max, e2e, _ = args
            shapes = [tuple(v.shape[2:]) for v in xin]
            strides = [H_in / s[0] for s in shapes]
            braw, sraw = detect_head_raw(sd, p, xin, nc, reg_max, False, L.get("legacy", False))
            x = obb_decode(sd, p, xin, braw, sraw, shapes, strides, reg_max)
            ys["detect_raw"] = (braw, sraw, x)
        elif t == "Segment":   # Detect + mask coefficients appended to the dense prediction, prototypes alongside (head.py:317-335)
            nc, nm, npr, reg_max, e2e, _ = args
            shapes = [tuple(v.shape[2:]) for v in xin]
            strides = [H_in / s[0] for s in shapes]
            braw, sraw = detect_head_raw(sd, p, xin, nc, reg_max, False, L.get("legacy", False))
            y = detect_decode(braw, sraw, shapes, strides, False, reg_max)
            mc, proto = segment_extras(sd, p, xin, nm)
            x = torch.cat([y, mc], 1)
            ys["detect_raw"], ys["proto"] = (braw, sraw, y), proto
        elif t == "Pose":      # Detect + decoded keypoints appended to the dense prediction (head.py:608-611); one2many head
            nc, kpt_shape, reg_max, e2e, _ = args
            shapes = [tuple(v.shape[2:]) for v in xin]
            strides = [H_in / s[0] for s in shapes]
            braw, sraw = detect_head_raw(sd, p, xin, nc, reg_max, False, L.get("legacy", False))
            y = detect_decode(braw, sraw, shapes, strides, False, reg_max)
            x = torch.cat([y, pose_kpts(sd, p, xin, kpt_shape, shapes, strides)], 1)
            ys["detect_raw"] = (braw, sraw, y)
        else:
            raise NotImplementedError(t)
        ys[i] = x
        outs.append(x)
    return (x, ys) if return_layers else x


# ======================================================================================================================
# Mixture-of-Transformers (nn/modules/mot/*) and Mixture-of-Attention (nn/modules/moa/*): eval forward
# ======================================================================================================================
ROUTE_TAP = None   # tests set this to a dict to capture MoT / MoA router outputs by module path


def _gn(sd, p, x, groups, eps=1e-5):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _sdpa(q, k, v, scale):
    """F.scaled_dot_product_attention(q, k, v, scale=scale) (mot/experts.py:37-69, moa/heads.py:37-52), math form."""
    return ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1) @ v


def _pad_to_window(x, win):
    """`_WindowTransformerExpert._pad_to_window` mot/experts.py:236-244 on NHWC (zero pad bottom/right)."""
    _, H, W, _ = x.shape
    ph, pw = (win - H % win) % win, (win - W % win) % win
    return F.pad(x, (0, 0, 0, pw, 0, ph)) if (ph or pw) else x


def _win_part(x, win):
    """[B,H,W,C] -> [B*nH*nW, win*win, C]  mot/experts.py:246-251."""
    B, H, W, C = x.shape
    return x.view(B, H // win, win, W // win, win, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, win * win, C)


def _win_rev(w, win, H, W):
    """mot/experts.py:253-261."""
    B = w.shape[0] // ((H // win) * (W // win))
    return w.view(B, H // win, W // win, win, win, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def mot_local_expert(sd, p, x, nh, local_window=0):
    """`_LocalConvTransformerExpert.forward` mot/experts.py:122-171."""
    B, C, H, W = x.shape
    N, hd = H * W, C // nh
    g = get_safe_groups(C, 8)
    xn = _gn(sd, p + ".norm1", x, g)
    qkv = _st(F.conv2d(F.conv2d(xn, _w(sd[p + ".dw_mix.weight"]), None, 1, 1, 1, C), _w(sd[p + ".qkv.weight"])))
    q, k, v = qkv.split(C, dim=1)
    v = _st(v + F.conv2d(v, _w(sd[p + ".pe.weight"]), None, 1, 3, 1, C))
    scale = hd ** -0.5
    if local_window > 0 and N > local_window ** 2:
        win = local_window
        qw, kw, vw = (_win_part(_pad_to_window(t.permute(0, 2, 3, 1), win), win) for t in (q, k, v))
        Hp, Wp = H + (win - H % win) % win, W + (win - W % win) % win
        qw, kw, vw = (t.reshape(-1, win * win, nh, hd).permute(0, 2, 1, 3) for t in (qw, kw, vw))
        o = _sdpa(qw, kw, vw, scale).transpose(1, 2).reshape(-1, win * win, C)
        o = _win_rev(o, win, Hp, Wp)[:, :H, :W, :].permute(0, 3, 1, 2)
    else:
        th = lambda t: t.reshape(B, nh, hd, N).transpose(2, 3)
        o = _sdpa(th(q), th(k), th(v), scale).transpose(2, 3).reshape(B, C, H, W)
    x = _st(x + sd[p + ".ls1"] * F.conv2d(_st(o), _w(sd[p + ".proj.weight"])))
    xn = _gn(sd, p + ".norm2", x, g)
    ffn = _st(torch.sigmoid(conv_block(sd, p + ".ffn_gate.0", xn)) * conv_block(sd, p + ".ffn_val", xn))
    return _st(x + sd[p + ".ls2"] * conv_block(sd, p + ".ffn_out", ffn, act=False))


def mot_window_expert(sd, p, x, nh, win=7, shift=0):
    """`_WindowTransformerExpert.forward` mot/experts.py:263-315 (pad before LayerNorm, roll without mask)."""
    B, C, H0, W0 = x.shape
    hd = C // nh
    x = _pad_to_window(x.permute(0, 2, 3, 1), win)
    H, W = x.shape[1:3]
    if shift:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    w = _win_part(_ln(sd, p + ".norm1", x), win)
    Bw = w.shape[0]
    qkv = _st(F.linear(w, _w(sd[p + ".qkv.weight"]))).reshape(Bw, win * win, 3, nh, hd).permute(2, 0, 3, 1, 4)
    o = _sdpa(qkv[0], qkv[1], qkv[2], hd ** -0.5).transpose(1, 2).reshape(Bw, win * win, C)
    o = _win_rev(F.linear(_st(o), _w(sd[p + ".proj.weight"])), win, H, W)
    if shift:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    x = _st(x[:, :H0, :W0] + sd[p + ".ls1"] * o[:, :H0, :W0])
    h = _st(F.gelu(F.linear(_ln(sd, p + ".norm2", x), _w(sd[p + ".ffn.0.weight"]), sd[p + ".ffn.0.bias"])))
    x = _st(x + sd[p + ".ls2"] * F.linear(h, _w(sd[p + ".ffn.3.weight"]), sd[p + ".ffn.3.bias"]))
    return x.permute(0, 3, 1, 2).contiguous()


def mot_deform_expert(sd, p, x, nh, n_points=4, align_corners=True):
    """`_DeformableTransformerExpert.forward/_deform_attn` mot/experts.py:394-496."""
    B, C, H, W = x.shape
    N, hd = H * W, C // nh
    xf = x.flatten(2).transpose(1, 2)
    xn = _ln(sd, p + ".norm1", xf)
    q = _st(F.linear(xn, _w(sd[p + ".q_proj.weight"])))
    off = F.linear(q, sd[p + ".offset_proj.weight"], sd[p + ".offset_proj.bias"]).reshape(B, N, nh, n_points, 2).tanh()
    aw = F.linear(q, sd[p + ".attn_proj.weight"], sd[p + ".attn_proj.bias"]).reshape(B, N, nh, n_points).softmax(-1)
    idx = torch.arange(N)
    row = (idx // W).float() / max(H - 1, 1) * 2 - 1
    col = (idx % W).float() / max(W - 1, 1) * 2 - 1
    ref = torch.stack([col, row], -1)[None, :, None, None, :]
    locs = (ref + off * 0.25).clamp(-1.0, 1.0)
    v = _st(F.linear(xn, _w(sd[p + ".v_proj.weight"]))).permute(0, 2, 1).reshape(B * nh, hd, H, W)
    samp = F.grid_sample(v, locs.permute(0, 2, 1, 3, 4).reshape(B * nh, N, n_points, 2), mode="bilinear",
                         padding_mode="zeros", align_corners=align_corners)
    samp = samp.reshape(B, nh, hd, N, n_points).permute(0, 3, 1, 4, 2)
    o = _st((aw.unsqueeze(-1) * samp).sum(3).reshape(B, N, C))
    xf = _st(xf + sd[p + ".ls1"] * F.linear(o, _w(sd[p + ".out_proj.weight"])))
    h = _st(F.gelu(F.linear(_ln(sd, p + ".norm2", xf), _w(sd[p + ".ffn.0.weight"]), sd[p + ".ffn.0.bias"])))
    xf = _st(xf + sd[p + ".ls2"] * F.linear(h, _w(sd[p + ".ffn.3.weight"]), sd[p + ".ffn.3.bias"]))
    return xf.transpose(1, 2).reshape(B, C, H, W)


def mot_router(sd, p, x, top_k, num_experts=3):
    """`_MoTRouter.forward` mot/router.py:211-291 (spatial, fp32): returns dense weights [B,E,H,W], indices [B,k,H,W], logits."""
    hidden = sd[p + ".router.0.weight"].shape[0]
    h = F.conv2d(x.float(), sd[p + ".router.0.weight"])
    h = F.silu(_gn(sd, p + ".router.1", h, get_safe_groups(hidden, 4)))
    logits = F.conv2d(h, sd[p + ".router.3.weight"], sd[p + ".router.3.bias"])
    w = F.softmax(logits / sd[p + ".temperature"].float(), dim=1)
    if top_k < num_experts:
        vals, idx = w.topk(top_k, dim=1)
        vals = vals / vals.sum(1, keepdim=True).clamp_min(1e-6)
        w = torch.zeros_like(w).scatter_(1, idx, vals)
    else:
        idx = torch.arange(num_experts).view(1, -1, 1, 1).expand(x.shape[0], -1, x.shape[2], x.shape[3])
    return w, idx, logits


def mot_block(sd, p, x, nh, top_k=2, win=7, n_points=4, shift=False, local_window=0, return_route=False):
    """`MoTBlock.forward` mot/block.py:401-415 with the eval sample-sparse blend :347-364."""
    w, idx, logits = mot_router(sd, p + ".router", x, top_k)
    if ROUTE_TAP is not None:
        ROUTE_TAP[p + ".router"] = (w, idx)
    experts = (lambda t: mot_local_expert(sd, p + ".experts.0", t, nh, local_window),
               lambda t: mot_window_expert(sd, p + ".experts.1", t, nh, win, win // 2 if shift else 0),
               lambda t: mot_deform_expert(sd, p + ".experts.2", t, nh, n_points))
    out = torch.zeros_like(x)
    for e, fn in enumerate(experts):
        active = (idx == e).reshape(x.shape[0], -1).any(1)
        b = torch.nonzero(active, as_tuple=True)[0]
        if b.numel():
            out[b] = _st(out[b] + fn(x[b]) * _st(w[b, e:e + 1]))
    C = x.shape[1]
    o = _st(_gn(sd, p + ".out_norm", F.conv2d(out, _w(sd[p + ".out_proj.weight"])), get_safe_groups(C, 8)) + x)
    return (o, w, idx, logits) if return_route else o


def c2f_heads_mot(dim, num_heads):
    """Head clamping of `C2fMoT.__init__` mot/wrappers.py:74-84."""
    h = num_heads
    while h > 1 and (dim % h != 0 or dim // h < 8):
        h -= 1
    return max(1, h)


def layer_c2f_mot(sd, p, x, c1, c2, n=1, num_heads=6, top_k=2, window_size=7, n_points=4, mlp_ratio=2.0, temperature=1.0,
                  balance_loss_coeff=0.01, e=0.5, *unused):
    """`C2fMoT.forward` mot/wrappers.py:116-147."""
    c = int(c2 * e)
    nh = c2f_heads_mot(c, num_heads)
    y = list(conv_block(sd, p + ".cv1", x).chunk(2, 1))
    for j in range(n):
        y.append(mot_block(sd, f"{p}.m.{j}", y[-1], nh, top_k, window_size, n_points, shift=bool(j % 2)))
    return conv_block(sd, p + ".cv2", torch.cat(y, 1))


# ---- MoA -------------------------------------------------------------------------------------------------------------
def moa_router(sd, p, x, temperature):
    """`_MoARouter.forward` moa/router.py:50-62 (fp32 soft routing over the 3 head groups)."""
    hidden = sd[p + ".router.0.weight"].shape[0]
    h = F.silu(_gn(sd, p + ".router.1", F.conv2d(x.float(), sd[p + ".router.0.weight"]), get_safe_groups(hidden, 4)))
    logits = F.conv2d(h, sd[p + ".router.3.weight"], sd[p + ".router.3.bias"]) / max(temperature, 0.1)
    return F.softmax(logits, dim=1)


def _window_flash(q, k, v, scale, win, H, W):
    """`_window_flash_attn` moa/heads.py:88-121 on [B,nh,N,hd] (zero pad, no mask)."""
    B, nh, N, hd = q.shape
    win = max(1, min(int(win), H, W))
    ph, pw = (win - H % win) % win, (win - W % win) % win

    def part(t):
        t = F.pad(t.reshape(B, nh, H, W, hd), (0, 0, 0, pw, 0, ph))
        Hp, Wp = t.shape[2:4]
        return t.reshape(B, nh, Hp // win, win, Wp // win, win, hd).permute(0, 1, 2, 4, 3, 5, 6).reshape(-1, win * win, hd)

    Hp, Wp = H + ph, W + pw
    o = _sdpa(part(q), part(k), part(v), scale)
    o = o.reshape(B, nh, Hp // win, Wp // win, win, win, hd).permute(0, 1, 2, 4, 3, 5, 6).reshape(B, nh, Hp, Wp, hd)
    return o[:, :, :H, :W].reshape(B, nh, N, hd)


def moa_local_head(sd, p, x, nh, hd, win=7):
    """`_LocalAttnHead.forward` moa/heads.py:140-159."""
    B, C, H, W = x.shape
    N, inner = H * W, nh * hd
    qkv = _st(F.conv2d(F.conv2d(x, _w(sd[p + ".qkv_dw.weight"]), None, 1, 1, 1, C), _w(sd[p + ".qkv_pw.weight"])))
    q, k, v = qkv.split(inner, 1)
    v = _st(v + F.conv2d(v, _w(sd[p + ".pe.weight"]), None, 1, 3, 1, inner))
    th = lambda t: t.flatten(2).view(B, nh, hd, N).transpose(2, 3)
    o = _window_flash(th(q), th(k), th(v), hd ** -0.5, win, H, W).transpose(2, 3).reshape(B, inner, H, W)
    return _st(_gn(sd, p + ".norm", F.conv2d(_st(o), _w(sd[p + ".proj.weight"])), get_safe_groups(C, 8)))


def moa_region_head(sd, p, x, nh, hd, pool_stride=2, max_kv=4096):
    """`_RegionalAttnHead.forward` moa/heads.py:206-246."""
    B, C, H, W = x.shape
    inner = nh * hd
    if min(H, W) <= 1:
        kv = F.conv2d(x, _w(sd[p + ".kv_proj.weight"]))
    else:
        s = pool_stride
        if max_kv is not None:
            while max(1, H // s) * max(1, W // s) > max_kv:
                s *= 2
        kv = F.conv2d(_st(F.adaptive_avg_pool2d(x, (max(1, H // s), max(1, W // s)))), _w(sd[p + ".kv_proj.weight"]))
    kv = _st(kv)
    M = kv.shape[2] * kv.shape[3]
    k, v = kv.split(inner, 1)
    k = k.flatten(2).view(B, nh, hd, M).transpose(2, 3)
    v = v.flatten(2).view(B, nh, hd, M).transpose(2, 3)
    q = _st(F.conv2d(x, _w(sd[p + ".q_proj.weight"]))).flatten(2).view(B, nh, hd, H * W).transpose(2, 3)
    o = _sdpa(q, k, v, hd ** -0.5).transpose(2, 3).reshape(B, inner, H, W)
    return _st(_gn(sd, p + ".norm", F.conv2d(_st(o), _w(sd[p + ".proj.weight"])), get_safe_groups(C, 8)))


def _linear_attn(q, k, v, rf, limit=1e4, eps=1e-6):
    """`_GlobalAttnHead._linear_attn` moa/heads.py:318-352 (fp32, ReLU features on the persistent RF matrix)."""
    B, nh, N, hd = q.shape
    nb = rf.shape[0]
    s = nb ** -0.5
    qf = (F.relu(q @ rf.T * s) + eps).clamp(max=limit).reshape(B * nh, N, nb)
    kf = (F.relu(k @ rf.T * s) + eps).clamp(max=limit).reshape(B * nh, N, nb)
    vf = v.reshape(B * nh, N, hd)
    kv = kf.transpose(1, 2) @ vf
    numer = (qf @ kv).clamp(-limit, limit)
    denom = (qf @ kf.sum(1).unsqueeze(-1)).clamp(min=eps)
    return (numer / denom).reshape(B, nh, N, hd)


def moa_global_head(sd, p, x, nh, hd, threshold=512, blend=64):
    """`_GlobalAttnHead.forward` moa/heads.py:354-380."""
    B, C, H, W = x.shape
    N, inner = H * W, nh * hd
    qkv = _st(F.conv2d(x, _w(sd[p + ".qkv.weight"]))).flatten(2)
    th = lambda t: t.view(B, nh, hd, N).transpose(2, 3)
    q, k, v = (th(t) for t in qkv.split(inner, 1))
    rf = sd[p + "._rf_matrix"]
    if N <= threshold:
        o = _sdpa(q, k, v, hd ** -0.5)
        if N > threshold - blend:
            a = (N - (threshold - blend)) / blend
            o = (1 - a) * o + a * _linear_attn(q, k, v, rf)
    else:
        o = _linear_attn(q, k, v, rf)
    o = o.transpose(2, 3).reshape(B, inner, H, W)
    return _st(_gn(sd, p + ".norm", F.conv2d(_st(o), _w(sd[p + ".proj.weight"])), get_safe_groups(C, 8)))


def moa_block(sd, p, x, num_heads, temperature=1.0, shortcut=True, win=7, max_kv=4096, return_route=False):
    """`MoABlock.forward` moa/block.py:167-278 (dense soft mixture of the three head groups, eval, sparse_inference=False)."""
    C = x.shape[1]
    hd, hpg = max(C // num_heads, 16), num_heads // 3
    w = moa_router(sd, p + ".router", x, temperature)
    if ROUTE_TAP is not None:
        ROUTE_TAP[p + ".router"] = (w,)
    mixed = _st(w[:, 0:1]) * moa_local_head(sd, p + ".local_head", x, hpg, hd, win)
    mixed = mixed + _st(w[:, 1:2]) * moa_region_head(sd, p + ".region_head", x, hpg, hd, 2, max_kv)
    mixed = _st(mixed + _st(w[:, 2:3]) * moa_global_head(sd, p + ".global_head", x, hpg, hd))
    mixed = conv_block(sd, p + ".fusion", mixed, act=False)
    if shortcut:
        x = _st(x + sd[p + ".ls_attn"] * mixed)
        x = _st(x + sd[p + ".ls_ffn"] * conv_block(sd, p + ".ffn.1", conv_block(sd, p + ".ffn.0", x), act=False))
    else:
        x = _st(sd[p + ".ls_attn"] * mixed)
        x = _st(sd[p + ".ls_ffn"] * conv_block(sd, p + ".ffn.1", conv_block(sd, p + ".ffn.0", x), act=False))
    return (x, w) if return_route else x


def c2f_heads_moa(c, num_heads):
    """Head adjustment of `C2fMoA.__init__` moa/wrappers.py:96-121."""
    h = num_heads
    while h % 3 != 0:
        h += 1
    while c // h < 16 and h > 3:
        h -= 3
    return max(h, 3)


def layer_c2f_moa(sd, p, x, c1, c2, n=1, num_heads=6, mlp_ratio=2.0, temperature=1.0, shortcut=True, e=0.5,
                  aux_loss_coeff=0.01, local_window_size=7, sequential_heads=True, regional_max_kv_tokens=4096, *unused):
    """`C2fMoA.forward` moa/wrappers.py:147-186."""
    c = int(c2 * e)
    nh = c2f_heads_moa(c, num_heads)
    y = list(conv_block(sd, p + ".cv1", x).chunk(2, 1))
    for j in range(n):
        y.append(moa_block(sd, f"{p}.m.{j}", y[-1], nh, temperature, shortcut, local_window_size, regional_max_kv_tokens))
    return conv_block(sd, p + ".cv2", torch.cat(y, 1))


_LAYER_FN.update({"C2fMoT": layer_c2f_mot, "C2fMoA": layer_c2f_moa})
_MIX_BASE.update({"C2fMoT", "C2fMoA"})
_MIX_REPEAT.update({"C2fMoT", "C2fMoA"})


# ======================================================================================================================
# Gated MoE family (nn/modules/moe/gated.py): `VisualEnhancedAdaptiveGateMoE` as used by the v0_10 model zoo - eval forward.
# restated and pinned ahead of the CUDA path (no product code consumes this yet).
# ======================================================================================================================
def _gn_na(x, groups, eps=1e-5):
    return F.group_norm(x, groups, None, None, eps)


def gated_se_gate(sd, p, x):
    """`se_gate` gated.py:325-332: GAP -> Linear(no bias) -> SiLU -> Linear -> Sigmoid.  Returns [B, C]."""
    g = x.mean((2, 3))
    g = F.silu(F.linear(g, sd[p + ".2.weight"]))
    return torch.sigmoid(F.linear(g, sd[p + ".4.weight"], sd[p + ".4.bias"]))


def visual_detail_gate(sd, p, x, groups=8):
    """`VisualDetailGate.forward` gated.py:1154-1178: x * (1 + tanh(scale) * gate(x - avgpool3(x)))."""
    C = x.shape[1]
    d = x - F.avg_pool2d(x, 3, 1, 1)
    t = F.conv2d(d, _w(sd[p + ".detail_filter.0.weight"]), None, 1, 1, 1, C)
    t = F.silu(_gn(sd, p + ".detail_filter.1", t, get_safe_groups(C, groups)))
    t = F.silu(F.conv2d(t, _w(sd[p + ".detail_filter.3.weight"])))
    gate = torch.sigmoid(F.conv2d(t, _w(sd[p + ".detail_filter.5.weight"]), sd[p + ".detail_filter.5.bias"]))
    return _st(x * (1 + torch.tanh(sd[p + ".detail_scale"]) * gate))


def _multi_head_global_logits(sd, p, stats):
    """Global branch of `MultiHeadRouterV3.forward` gated.py:2123-2140 (v0_13): a dense projection blended with per-head projections
    of consecutive chunks of the normalised statistics."""
    B = stats.shape[0]
    nh = sum(1 for k in sd if k.startswith(p + ".heads.") and k.endswith(".weight"))
    hd = sd[p + ".heads.0.weight"].shape[1]
    hw = torch.sigmoid(sd[p + ".head_alpha"])
    hw = hw / (hw.sum() + 1e-6)
    gw = torch.sigmoid(sd[p + ".global_weight"])
    if stats.shape[1] < hd * nh:
        padded = F.pad(stats, (0, hd * nh - stats.shape[1]))
    else:
        padded = stats[:, :hd * nh]
    chunks = padded.view(B, nh, hd)
    out = gw * F.linear(stats, sd[p + ".global_proj.weight"])
    for i in range(nh):
        out = out + (1 - gw) * hw[i] * F.linear(chunks[:, i], sd[f"{p}.heads.{i}.weight"])
    return out


def cross_path_gate(sd, p, cat):
    """`CrossPathGate.forward` gated.py:2396-2412 (v0_15) on the concatenated [static | dynamic] outputs: per-(image, channel) gate
    0.5 + tanh(gate_scale) * 0.5 * sigmoid(MLP(GAP(cat))); only the first C of the MLP's 2C outputs are used."""
    C = cat.shape[1]
    g = F.silu(F.linear(cat.mean((2, 3)), sd[p + ".gate_net.2.weight"]))
    g = F.linear(g, sd[p + ".gate_net.4.weight"], sd[p + ".gate_net.4.bias"])
    gate = 0.5 + torch.tanh(sd[p + ".gate_scale"]) * 0.5 * torch.sigmoid(g)
    return _st(cat * gate[:, :C, None, None])


def dual_stream_gate_router(sd, p, x, top_k, temperature, pool_scale=4):
    """`DualStreamGateRouter.forward` gated.py:129-160 (fp32): global mean/std statistics + pooled local conv stream, blended by
    sigmoid(alpha), clamp +-30, softmax(/T), top-k, renormalise with +1e-6.  Returns (weights [B,k], indices [B,k], probs)."""
    B, C, H, W = x.shape
    xf = x.float()
    mean = xf.mean((2, 3))
    std = xf.std((2, 3), unbiased=False) if H * W > 1 else torch.zeros_like(mean)
    stats = torch.cat([mean, std], 1)
    if p + ".stat_norm.weight" in sd:      # DualStreamGateRouterV2 gated.py:181-260 (v0_11 / v0_12): LayerNorm on the statistics
        stats = F.layer_norm(stats, (2 * C,), sd[p + ".stat_norm.weight"], sd[p + ".stat_norm.bias"], 1e-5)
    gl = _multi_head_global_logits(sd, p, stats) if p + ".heads.0.weight" in sd else F.linear(stats, sd[p + ".global_fc.weight"])
    xl = F.avg_pool2d(xf, pool_scale, pool_scale) if (H > pool_scale and W > pool_scale) else xf
    t = F.conv2d(xl, sd[p + ".local_conv.0.weight"], None, 1, 1, 1, C)
    t = F.silu(_gn(sd, p + ".local_conv.1", t, get_safe_groups(C, 8)))
    t = F.conv2d(t, sd[p + ".local_conv.3.weight"])
    t = F.silu(_gn(sd, p + ".local_conv.4", t, get_safe_groups(t.shape[1], 4)))
    ll = F.conv2d(t, sd[p + ".local_conv.6.weight"], sd[p + ".local_conv.6.bias"]).mean((2, 3))
    a = torch.sigmoid(sd[p + ".alpha"])
    logits = a * gl + (1 - a) * ll
    if p + ".expert_prior" in sd:          # ... and a learnable per-expert prior, added before the clamp
        logits = logits + sd[p + ".expert_prior"].view(1, -1)
    logits = logits.clamp(-30.0, 30.0)
    probs = F.softmax(logits / temperature, dim=1)
    w, idx = torch.topk(probs, top_k, dim=1)
    return w / (w.sum(1, keepdim=True) + 1e-6), idx, probs


def complexity_gate(w, complexity):
    """`AdaptiveGateMoE._apply_complexity_gate` gated.py:469-490: keep the round(c * k) best ranks (c clamped to [0.3, 1.5]),
    renormalise.  `complexity` is ONE scalar for the whole batch (gated.py:455-461)."""
    k = w.shape[1]
    if k <= 1:
        return w
    c = torch.nan_to_num(complexity, nan=1.0, posinf=1.0, neginf=1.0).clamp(0.3, 1.5)
    keep = torch.round(c * k).clamp(1, k)
    mask = (torch.arange(1, k + 1, dtype=keep.dtype).view(1, k) <= keep).to(w.dtype)
    w = w * mask
    return w / w.sum(1, keepdim=True).clamp_min(1e-6)


def fused_expert_group(sd, p, x, w, idx, num_experts, out_channels, num_groups=8):
    """`FusedExpertGroup.forward` gated.py:1061-1081: one grouped 3x3 conv computes every expert, the top-k outputs are gathered,
    GroupNorm (no affine) + per-expert affine + SiLU, weighted sum."""
    B, C, H, W = x.shape
    fw = sd[p + ".fused_conv.weight"]
    groups = C // fw.shape[1]
    fo = _st(F.conv2d(x, _w(fw), None, 1, 1, 1, groups)).view(B, num_experts, out_channels, H, W)
    k = idx.shape[1]
    sel = torch.gather(fo, 1, idx.view(B, k, 1, 1, 1).expand(B, k, out_channels, H, W))
    nrm = _gn_na(sel.reshape(B * k, out_channels, H, W), get_safe_groups(out_channels, num_groups)).view(B, k, out_channels, H, W)
    nrm = nrm * sd[p + ".expert_norm_weight"][idx].view(B, k, out_channels, 1, 1) + sd[p + ".expert_norm_bias"][idx].view(B, k, out_channels, 1, 1)
    return _st((F.silu(nrm) * w.view(B, k, 1, 1, 1)).sum(1))


def low_rank_fused_expert_group(sd, p, x, w, idx, num_experts, out_channels, num_groups=8):
    """`LowRankFusedExpertGroup.forward` gated.py:1101-1147: shared 1x1 bottleneck -> GN -> SiLU -> FusedExpertGroup."""
    t = F.conv2d(x, _w(sd[p + ".bottleneck.0.weight"]))
    t = _st(F.silu(_gn(sd, p + ".bottleneck.1", t, get_safe_groups(t.shape[1], num_groups))))
    return fused_expert_group(sd, p + ".fused", t, w, idx, num_experts, out_channels, num_groups)


def shared_inverted_expert_group(sd, p, x, w, idx, num_experts, out_channels):
    """`SharedInvertedExpertGroup.forward` moe/experts.py:235-269: shared expand 1x1 -> GN -> SiLU -> dw3x3 -> GN -> SiLU, then a
    1x1 + GN projection per ACTIVE expert, weighted index_add (routes with weight <= 0 are dropped)."""
    B, C, H, W = x.shape
    t = F.conv2d(x, _w(sd[p + ".shared_feature.0.weight"]))
    hid = t.shape[1]
    t = F.silu(_gn(sd, p + ".shared_feature.1", t, get_safe_groups(hid, 8)))
    t = F.conv2d(_st(t), _w(sd[p + ".shared_feature.3.weight"]), None, 1, 1, 1, hid)
    feat = _st(F.silu(_gn(sd, p + ".shared_feature.4", t, get_safe_groups(hid, 8))))
    out = torch.zeros(B, out_channels, H, W)
    valid = w > 0.0
    for e in torch.unique(idx[valid]).tolist():
        bi, ki = torch.where((idx == e) & valid)
        y = _gn(sd, f"{p}.expert_projections.{e}.1", F.conv2d(feat[bi], _w(sd[f"{p}.expert_projections.{e}.0.weight"])),
                get_safe_groups(out_channels, 8))
        out.index_add_(0, bi, _st(y) * w[bi, ki].view(-1, 1, 1, 1))
    return _st(out)


def diversified_expert_group(sd, p, x, w, idx, num_experts, out_channels):
    """`DiversifiedExpertGroup.forward` moe/gated.py:2297-2333 (v0_14): shared expand 1x1 -> GN -> SiLU, then per ACTIVE expert e a
    depthwise 3x3 with dilation 1 + e // 2 (:2270-2276; the `dw_dilations` parameters are stored but never read) -> GN -> SiLU,
    1x1 + GN projection, weighted index_add (routes with weight <= 0 are dropped)."""
    B, C, H, W = x.shape
    t = F.conv2d(x, _w(sd[p + ".shared_expand.0.weight"]))
    hid = t.shape[1]
    feat = _st(F.silu(_gn(sd, p + ".shared_expand.1", t, get_safe_groups(hid, 8))))
    out = torch.zeros(B, out_channels, H, W)
    valid = w > 0.0
    for e in torch.unique(idx[valid]).tolist():
        d = 1 + e // 2
        bi, ki = torch.where((idx == e) & valid)
        t = F.conv2d(feat[bi], _w(sd[f"{p}.dw_layers.{e}.0.weight"]), None, 1, d, d, hid)
        t = _st(F.silu(_gn(sd, f"{p}.dw_layers.{e}.1", _st(t), get_safe_groups(hid, 8))))
        y = _gn(sd, f"{p}.expert_projections.{e}.1", F.conv2d(t, _w(sd[f"{p}.expert_projections.{e}.0.weight"])),
                get_safe_groups(out_channels, 8))
        out.index_add_(0, bi, _st(y) * w[bi, ki].view(-1, 1, 1, 1))
    return _st(out)


def pyramid_context_mixer(sd, p, x, groups=8, pool_scales=(2, 4)):
    """`PyramidContextMixer.forward` gated.py:1210-1221."""
    B, C, H, W = x.shape
    g = get_safe_groups(C, groups)
    ctx = [F.silu(_gn(sd, p + ".local_context.1", F.conv2d(x, _w(sd[p + ".local_context.0.weight"]), None, 1, 1, 1, C), g))]
    for i, s in enumerate(pool_scales):
        h, w = max(1, H // s), max(1, W // s)
        pooled = x if (H, W) == (h, w) else F.adaptive_avg_pool2d(x, (h, w))
        t = F.silu(_gn(sd, f"{p}.pool_projections.{i}.1", F.conv2d(pooled, _w(sd[f"{p}.pool_projections.{i}.0.weight"])), g))
        ctx.append(F.interpolate(t, size=(H, W), mode="nearest"))
    c = torch.stack(ctx, 0).mean(0)
    gate = torch.sigmoid(F.conv2d(c, _w(sd[p + ".context_gate.0.weight"]), sd[p + ".context_gate.0.bias"]))
    return _st(x + torch.tanh(sd[p + ".context_scale"]) * c * gate)


def feature_refine(sd, p, x, groups=8):
    """`FeatureRefinementHook` moe/hooks.py:50-57: x + tanh(scale) * refiner(x) * gate(x)."""
    C = x.shape[1]
    r = F.silu(_gn(sd, p + ".feature_refiner.1", F.conv2d(x, _w(sd[p + ".feature_refiner.0.weight"]), None, 1, 1, 1, C), get_safe_groups(C, groups)))
    g = F.silu(F.conv2d(x.mean((2, 3), keepdim=True), _w(sd[p + ".feature_gate.1.weight"])))
    g = torch.sigmoid(F.conv2d(g, _w(sd[p + ".feature_gate.3.weight"]), sd[p + ".feature_gate.3.bias"]))
    return _st(x + torch.tanh(sd[p + ".refine_scale"]) * r * g)


def light_refine(sd, p, x, groups=8):
    """`OptimalHybridGateMoE._apply_refine` gated.py:1958-1961 (v0_12): x + tanh(scale) * GN(dw3x3(x)) * SE(x) - no activation on
    the depthwise branch, unlike `feature_refine`."""
    C = x.shape[1]
    r = _gn(sd, p + ".refine_dw.1", F.conv2d(x, _w(sd[p + ".refine_dw.0.weight"]), None, 1, 1, 1, C), get_safe_groups(C, groups))
    g = F.silu(F.conv2d(x.mean((2, 3), keepdim=True), _w(sd[p + ".refine_gate.1.weight"])))
    g = torch.sigmoid(F.conv2d(g, _w(sd[p + ".refine_gate.3.weight"]), sd[p + ".refine_gate.3.bias"]))
    return _st(x + torch.tanh(sd[p + ".refine_scale"]) * r * g)


def gated_moe_forward(sd, p, x, c1, c2, num_experts, top_k, split_ratio, num_groups, temperature, backend, shuffle_groups, hooks,
                      return_route=False):
    """Eval forward shared by the whole AdaptiveGateMoE line: `AdaptiveGateMoE.forward` gated.py:508-555 (v0.4, v0.5),
    `HybridAdaptiveGateMoE.forward` :1340-1386 (v0.6, v0.7: + channel shuffle) and `run_visual_hybrid_moe_forward`
    moe/_gated_visual.py:33-86 (v0.8 ... v0.10: + hooks).  `backend` in {"shared_inverted", "fused", "low_rank_fused"}; `hooks` is the
    ordered subset of ("detail", "context", "refine") the class configures (detail runs before routing, the others after the
    concatenation; the complexity scalar is taken from the dynamic half the router sees)."""
    dyn = int(c1 * split_ratio)
    st_c = c1 - dyn
    out_dyn = int(c2 * split_ratio)
    gate = gated_se_gate(sd, p + ".se_gate", x)
    xs = _st(x[:, :st_c] * gate[:, :st_c, None, None])
    xd = _st(x[:, st_c:] * gate[:, st_c:, None, None])
    if "detail" in hooks:
        xd = visual_detail_gate(sd, p + ".detail_gate", xd, num_groups)
    # static path: dw3x3 -> BN -> SiLU -> 1x1 -> BN -> SiLU (gated.py:335-344)
    t = F.silu(_bn(sd, p + ".static_net.1", F.conv2d(xs, _w(sd[p + ".static_net.0.weight"]), None, 1, 1, 1, st_c)))
    out_s = _st(F.silu(_bn(sd, p + ".static_net.4", F.conv2d(_st(t), _w(sd[p + ".static_net.3.weight"])))))
    # batch-level complexity scalar (gated.py:455-461): sigmoid(conv1x1(GAP(x_dynamic))) averaged over the WHOLE batch
    cx = torch.sigmoid(F.conv2d(xd.mean((2, 3), keepdim=True), sd[p + ".complexity_estimator.1.weight"], sd[p + ".complexity_estimator.1.bias"])).mean()
    cx = cx.clamp(0.3, 1.5) if bool(torch.isfinite(cx)) else torch.tensor(1.0)
    w, idx, probs = dual_stream_gate_router(sd, p + ".routing", xd, top_k, max(float(temperature), 1e-3))
    w = complexity_gate(w, cx)
    if p + ".fused_experts.dw_layers.0.0.weight" in sd:     # DiversifiedExpertMoE gated.py:2552-2560 (v0_14): group replaced
        out_d = diversified_expert_group(sd, p + ".fused_experts", xd, w, idx, num_experts, out_dyn)
    elif backend == "low_rank_fused":
        out_d = low_rank_fused_expert_group(sd, p + ".fused_experts", xd, w, idx, num_experts, out_dyn, num_groups)
    elif backend == "fused":
        out_d = fused_expert_group(sd, p + ".fused_experts", xd, w, idx, num_experts, out_dyn, num_groups)
    else:
        out_d = shared_inverted_expert_group(sd, p + ".fused_experts", xd, w, idx, num_experts, out_dyn)
    cat = torch.cat([out_s, out_d], 1)
    if p + ".cross_gate.gate_scale" in sd:     # GatedFusionMoE gated.py:2652-2653 (v0_15): content-aware gate on both paths
        cat = cross_path_gate(sd, p + ".cross_gate", cat)
    sg = shuffle_groups if (shuffle_groups and c2 % shuffle_groups == 0) else 1
    if sg > 1:
        B, C, H, W = cat.shape
        cat = cat.view(B, sg, C // sg, H, W).transpose(1, 2).reshape(B, C, H, W)
    for h in hooks:
        if h == "context":
            cat = pyramid_context_mixer(sd, p + ".context_mixer", cat, num_groups)
        elif h == "refine":
            cat = feature_refine(sd, p, cat, num_groups)
        elif h == "light_refine":
            cat = light_refine(sd, p, cat, num_groups)
    out = _st(_gn(sd, p + ".bn", F.conv2d(cat, _w(sd[p + ".proj.weight"])), get_safe_groups(c2, num_groups)) + x)
    return (out, w, idx, probs) if return_route else out


# class name -> (default initial_temperature, hybrid back-end choice?, low-rank?, shuffle?, hooks): the constructor defaults and
# the forward each class of the line uses (gated.py:268-1766)
GATED_VARIANTS = {
    "AdaptiveGateMoE": (1.0, False, False, False, ()),                                               # v0.4
    "FusedAdaptiveGateMoE": (1.0, None, False, False, ()),                                           # v0.5: always fused
    "HybridAdaptiveGateMoE": (1.2, True, False, True, ()),                                           # v0.6
    "LowRankHybridAdaptiveGateMoE": (1.2, True, True, True, ()),                                     # v0.7
    "RefinedLowRankHybridAdaptiveGateMoE": (1.2, True, True, True, ("refine",)),                     # v0.8
    "DetailAwareLowRankHybridAdaptiveGateMoE": (1.2, True, True, True, ("detail",)),                 # v0.9
    "ContextRefinedLowRankHybridAdaptiveGateMoE": (1.2, True, True, True, ("context", "refine")),
    "VisualEnhancedAdaptiveGateMoE": (1.2, True, True, True, ("detail", "context", "refine")),       # v0.10
    "HybridAdaptiveGateMoEv2": (1.2, True, False, True, ()),                 # v0.11: v0.6 + DualStreamGateRouterV2 (keys in the sd)
    "OptimalHybridGateMoE": (1.2, True, False, True, ("light_refine",)),     # v0.12: + depthwise refinement
    "MultiHeadRouterMoE": (1.2, True, False, True, ("light_refine",)),       # v0.13: v0.12 + MultiHeadRouterV3 (keys in the sd)
    "DiversifiedExpertMoE": (1.2, True, False, True, ("light_refine",)),     # v0.14: v0.12 + DiversifiedExpertGroup (keys in the sd)
    "GatedFusionMoE": (1.2, True, False, True, ("light_refine",)),           # v0.15: v0.12 + CrossPathGate (keys in the sd)
    "SharedExpertMoE": (1.2, True, True, True, ()),       # moe/shared_expert_moe.py: v0.7 blocks sharing one expert group per pool_id
}


def zero_cost_router(sd, p, x, top_k, temperature):
    """`ZeroCostRouter.forward` gated.py:953-968 (fp32): `router` = Linear -> Softmax, whose OUTPUT is divided by T, clamped and
    sent through softmax again; top-k; renormalise with +1e-6.  Returns (weights [B,k], indices [B,k], probs)."""
    B, C, H, W = x.shape
    xf = x.float()
    mean = xf.mean((2, 3))
    std = xf.std((2, 3), unbiased=False) if H * W > 1 else torch.zeros_like(mean)
    p0 = F.softmax(F.linear(torch.cat([mean, std], 1), sd[p + ".router.0.weight"]), dim=1)
    probs = F.softmax((p0 / temperature).clamp(-30.0, 30.0), dim=1)
    w, idx = torch.topk(probs, top_k, dim=1)
    return w / (w.sum(1, keepdim=True) + 1e-6), idx, probs


def layer_ultimate_optimized_moe(sd, p, x, c1, c2, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, use_routing_cache=True,
                                 capacity_factor=1.5, initial_temperature=2.0, *unused, return_route=False):
    """`UltimateOptimizedMoE.forward` moe/modules.py:1653-1680 (v0_3 zoo), eval: un-gated channel split, static path, ZeroCostRouter
    on the dynamic half, routing weights SCALED by the batch-level complexity (clamped to [0.3, 1.5]), FusedExpertGroup,
    concatenation, 1x1 projection, GroupNorm, residual."""
    dyn = int(c1 * split_ratio)
    st_c = c1 - dyn
    out_dyn = int(c2 * split_ratio)
    xs, xd = x[:, :st_c], x[:, st_c:]
    cx = torch.sigmoid(F.conv2d(xd.mean((2, 3), keepdim=True), sd[p + ".complexity_estimator.1.weight"], sd[p + ".complexity_estimator.1.bias"])).mean()
    cx = torch.nan_to_num(cx, nan=1.0, posinf=1.5, neginf=0.3).clamp(0.3, 1.5)
    t = F.silu(_bn(sd, p + ".static_net.1", F.conv2d(xs, _w(sd[p + ".static_net.0.weight"]), None, 1, 1, 1, st_c)))
    out_s = _st(F.silu(_bn(sd, p + ".static_net.4", F.conv2d(_st(t), _w(sd[p + ".static_net.3.weight"])))))
    w, idx, probs = zero_cost_router(sd, p + ".routing", xd, top_k, float(initial_temperature))
    w = w * cx
    out_d = fused_expert_group(sd, p + ".fused_experts", xd, w, idx, num_experts, out_dyn, num_groups)
    cat = torch.cat([out_s, out_d], 1)
    out = _st(_gn(sd, p + ".bn", F.conv2d(cat, _w(sd[p + ".proj.weight"])), get_safe_groups(c2, num_groups)) + x)
    return (out, w, idx, probs) if return_route else out


_LAYER_FN["UltimateOptimizedMoE"] = layer_ultimate_optimized_moe
_MIX_BASE.add("UltimateOptimizedMoE")


def gated_backend(name, num_experts, fused_expert_threshold=8):
    _, hybrid, low_rank, _, _ = GATED_VARIANTS[name]
    if hybrid is None:
        return "fused"
    if not hybrid or num_experts > fused_expert_threshold:
        return "shared_inverted"
    return "low_rank_fused" if low_rank else "fused"


def _gated_layer(name):
    temp0, _, _, shuffle, hooks = GATED_VARIANTS[name]

    def layer(sd, p, x, c1, c2, num_experts=4, top_k=2, split_ratio=0.5, num_groups=8, initial_temperature=temp0,
              final_temperature=0.5, balance_loss_coeff=1.0, router_z_loss_coeff=1.0, entropy_loss_coeff=0.01,
              fused_expert_threshold=8, shuffle_groups=2, *unused, return_route=False):
        return gated_moe_forward(sd, p, x, c1, c2, num_experts, top_k, split_ratio, num_groups, initial_temperature,
                                 gated_backend(name, num_experts, fused_expert_threshold), shuffle_groups if shuffle else 1, hooks,
                                 return_route=return_route)
    layer.__name__ = "layer_" + name
    layer.__doc__ = f"`{name}.forward` (nn/modules/moe/gated.py), eval; see gated_moe_forward."
    return layer


layer_visual_enhanced_gate_moe = _gated_layer("VisualEnhancedAdaptiveGateMoE")
_LAYER_FN["LatentMixture"] = layer_latent_mixture
_LAYER_FN["UltraOptimizedMoE"] = layer_ultra_optimized_moe
_MIX_BASE.add("UltraOptimizedMoE")
for _name in ("ModularRouterExpertMoE", "OptimizedMOEImproved"):
    _LAYER_FN[_name] = layer_modular_router_expert_moe
    _MIX_BASE.add(_name)
for _name in GATED_VARIANTS:
    _LAYER_FN[_name] = _gated_layer(_name)
    _MIX_BASE.add(_name)


