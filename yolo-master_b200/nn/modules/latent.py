"""`LatentMixture` with the reference's constructor and state_dict keys (`ultralytics/nn/modules/latent_mixture.py:452-800`; the
yolo26-master-latent-n* zoo): several aligned feature maps in, one out.

Every input is projected to the output width (1x1 -> GroupNorm(1) -> SiLU, or identity), pooled to one token per image; the tokens'
mean goes through `LatentRouter` (LayerNorm -> two-layer MLP -> expert head -> softmax, fp32: `ym_latent_router`) and the dense
mixture of `DenseChannelExpert`s (1x1 -> GN -> SiLU -> dw3x3 -> GN -> SiLU -> 1x1) of the base feature, weighted per image, is added
to the base through `residual_gain`.  Eval, dense dispatch (`inference_top_k` = num_experts), `router_only` value fusion."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ._base import PackCache, pack_gemm_weight, require_eval, to_nchw, to_nhwc
from .mot import _f32, _pack_dw, _pack_linear

__all__ = ("LatentMixture", "LatentRouter", "DenseChannelExpert")


def _make_divisible(x, divisor):
    return int(-(-x // divisor) * divisor)


def _conv1x1(c1, c2):
    return nn.Sequential(nn.Conv2d(c1, c2, kernel_size=1, bias=False), nn.GroupNorm(1, c2), nn.SiLU(inplace=True))


class DenseChannelExpert(nn.Module):
    """`DenseChannelExpert(channels, expert_ratio=0.25)` (latent_mixture.py:108-130)."""

    def __init__(self, channels, expert_ratio=0.25):
        super().__init__()
        hidden = _make_divisible(max(8, int(round(channels * expert_ratio))), 8)
        self.net = nn.Sequential(
            nn.Conv2d(channels, hidden, 1, bias=False), nn.GroupNorm(1, hidden), nn.SiLU(inplace=True),
            nn.Conv2d(hidden, hidden, 3, padding=1, groups=hidden, bias=False), nn.GroupNorm(1, hidden), nn.SiLU(inplace=True),
            nn.Conv2d(hidden, channels, 1, bias=False))


class LatentRouter(nn.Module):
    """`LatentRouter(latent_dim, num_experts, router_hidden_dim=None, temperature=1.0, noise_std=0.0, router_init_std=0.0,
    num_tokens=None, per_token=False)` (latent_mixture.py:133-241); parameter container, executed by `ym_latent_router`."""

    def __init__(self, latent_dim, num_experts, router_hidden_dim=None, temperature=1.0, noise_std=0.0, router_init_std=0.0,
                 num_tokens=None, per_token=False):
        super().__init__()
        if per_token:
            raise NotImplementedError("LatentRouter: per_token routing is not on the B200 path")
        hidden = latent_dim if router_hidden_dim is None else int(router_hidden_dim)
        self.latent_dim, self.num_experts, self.num_tokens = latent_dim, num_experts, num_tokens
        self.norm = nn.LayerNorm(latent_dim)
        self.trunk = nn.Sequential(nn.Linear(latent_dim, hidden), nn.SiLU(), nn.Linear(hidden, latent_dim), nn.SiLU())
        self.expert_head = nn.Linear(latent_dim, num_experts)
        if num_tokens is None:
            self.register_parameter("scale_embedding", None)
        else:
            self.scale_embedding = nn.Parameter(torch.zeros(num_tokens, latent_dim))
        self.register_buffer("_temperature", torch.tensor(float(temperature)), persistent=True)
        self.register_buffer("_noise_std", torch.tensor(float(noise_std)), persistent=True)

    def pack(self):
        pk = {"ln_w": _f32(self.norm.weight), "ln_b": _f32(self.norm.bias), "ln_eps": float(self.norm.eps),
              "w1": _f32(self.trunk[0].weight), "b1": _f32(self.trunk[0].bias), "w2": _f32(self.trunk[2].weight), "b2": _f32(self.trunk[2].bias),
              "wh": _f32(self.expert_head.weight), "bh": _f32(self.expert_head.bias), "temperature": float(self._temperature)}
        if self.scale_embedding is not None:
            pk["emb"] = _f32(self.scale_embedding)
        return pk


class LatentMixture(nn.Module, PackCache):
    """`LatentMixture(in_channels, out_channels, num_experts=4, expert_ratio=0.25, router_hidden_dim=None, temperature=1.0,
    balance_loss_coeff=1e-2, router_z_loss_coeff=1e-3, residual_init=0.0, noise_std=0.0, router_init_std=0.0, inference_top_k=None,
    value_fusion_mode="router_only", value_fusion_weights=None, require_inference_calibration=False)`."""

    def __init__(self, in_channels, out_channels, num_experts=4, expert_ratio=0.25, router_hidden_dim=None, temperature=1.0,
                 balance_loss_coeff=1e-2, router_z_loss_coeff=1e-3, residual_init=0.0, noise_std=0.0, router_init_std=0.0,
                 inference_top_k=None, value_fusion_mode="router_only", value_fusion_weights=None, require_inference_calibration=False):
        super().__init__()
        if isinstance(in_channels, int):
            in_channels = [in_channels]
        self.in_channels, self.out_channels = tuple(int(c) for c in in_channels), int(out_channels)
        self.num_inputs, self.num_experts = len(self.in_channels), int(num_experts)
        if value_fusion_mode != "router_only" or value_fusion_weights is not None:
            raise NotImplementedError("LatentMixture: only value_fusion_mode='router_only' is on the B200 path")
        if inference_top_k is not None and int(inference_top_k) != self.num_experts:
            raise NotImplementedError("LatentMixture: sparse inference (inference_top_k < num_experts) is not on the B200 path")
        if self.num_inputs > 4 or self.out_channels % 8 or self.out_channels > 256:
            raise NotImplementedError("LatentMixture: up to 4 inputs, out_channels a multiple of 8 and <= 256 (GroupNorm(1) width) on the B200 path")
        self.value_fusion_mode, self.top_k, self.inference_top_k = "router_only", self.num_experts, self.num_experts
        self.register_buffer("value_fusion_weights", torch.ones(self.num_inputs) / self.num_inputs, persistent=False)
        oc = self.out_channels
        self.base_proj = nn.Identity() if self.in_channels[0] == oc else _conv1x1(self.in_channels[0], oc)
        self.token_projs = nn.ModuleList([nn.Identity() if c == oc else _conv1x1(c, oc) for c in self.in_channels])
        self.router = LatentRouter(oc, self.num_experts, router_hidden_dim=router_hidden_dim, temperature=temperature, noise_std=noise_std,
                                   router_init_std=router_init_std, num_tokens=self.num_inputs, per_token=False)
        self.experts = nn.ModuleList(DenseChannelExpert(oc, expert_ratio) for _ in range(self.num_experts))
        self.residual_gain = nn.Parameter(torch.tensor(float(residual_init)))
        self.last_routing_snapshot: dict = {}

    # the reference keeps routing configuration in the module's extra state (latent_mixture.py:532-560): accepted and ignored here
    def get_extra_state(self):
        return {"schema_version": 1, "value_fusion_mode": self.value_fusion_mode, "inference_top_k": int(self.inference_top_k)}

    def set_extra_state(self, state):
        pass

    def _load_from_state_dict(self, state_dict, prefix, *args):
        state_dict.setdefault(prefix + "_extra_state", self.get_extra_state())
        super()._load_from_state_dict(state_dict, prefix, *args)

    @staticmethod
    def _proj_pack(seq):
        if isinstance(seq, nn.Identity):
            return None
        return (_pack_linear(seq[0].weight), (1, _f32(seq[1].weight), _f32(seq[1].bias), float(seq[1].eps)))

    def _build_pack(self):
        C = self.out_channels
        pk = {"base": self._proj_pack(self.base_proj), "tok": [self._proj_pack(p) for p in self.token_projs], "router": self.router.pack(),
              "gain": self.residual_gain.detach().float().reshape(1).expand(C).contiguous(), "experts": []}
        for e in self.experts:
            n = e.net
            gn = lambda m: (1, _f32(m.weight), _f32(m.bias), float(m.eps))
            pk["experts"].append({"c0": _pack_linear(n[0].weight), "g1": gn(n[1]), "dw": _pack_dw(n[3].weight), "g4": gn(n[4]),
                                  "c6": _pack_linear(n[6].weight), "hid": n[0].weight.shape[0]})
        return pk

    @staticmethod
    def _project(x, p):
        if p is None:
            return x
        (w, b), (G, gw, gb, eps) = p
        return ops.groupnorm(ops.conv2d(x, w, b, w.shape[0], 1, 1, 1, 0, False), G, gw, gb, eps=eps, act=True)

    def fwd_nhwc(self, xs, out=None):
        require_eval(self)
        pk = self.get_pack()
        B, H, W, _ = xs[0].shape
        C, HW = self.out_channels, H * W
        for x, c in zip(xs, self.in_channels):
            if tuple(x.shape[:3]) != (B, H, W) or x.shape[3] != c:
                raise ValueError(f"LatentMixture: inputs must share the spatial size and carry {self.in_channels} channels")
        tokens = [ops.gap(self._project(x, p)) for x, p in zip(xs, pk["tok"])]
        base = self._project(xs[0], pk["base"])
        probs, logits = ops.latent_router(tokens, pk["router"])
        self.last_routing_snapshot = {"router_probs": probs, "router_logits": logits}
        zero = torch.zeros((B, C), dtype=torch.float32, device=base.device)
        mixed = None
        for e, ep in enumerate(pk["experts"]):
            G, gw, gb, eps = ep["g1"]
            t = ops.groupnorm(ops.conv2d(base, *ep["c0"], ep["hid"], 1, 1, 1, 0, False), G, gw, gb, eps=eps, act=True)
            G, gw, gb, eps = ep["g4"]
            t = ops.groupnorm(ops.dwconv(t, ep["dw"], None, 3, False, ep["hid"]), G, gw, gb, eps=eps, act=True)
            y = ops.conv2d(t, *ep["c6"], C, 1, 1, 1, 0, False)
            gate = probs[:, e:e + 1].expand(B, C).contiguous()                      # per-image gate, the same for every channel
            mixed = ops.ew(ops.EW_AFFINE, a=y, b=mixed, p0=gate, p1=zero, rows_per_img=HW)
        return ops.ew(ops.EW_SCALE_RES, a=base, b=mixed, p0=pk["gain"], out=out)    # base + residual_gain * mixed

    def forward(self, xs):
        return to_nchw(self.fwd_nhwc([to_nhwc(x) for x in xs]))

    @property
    def aux_loss(self):
        return torch.zeros((), device=self.residual_gain.device)
