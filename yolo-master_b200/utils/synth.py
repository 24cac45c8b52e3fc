"""Deterministic synthetic parameters for parity tests and benchmarks.

There is no network access for released checkpoints, and the reference's default init
(identity BatchNorm statistics, std=0.05 router heads) gives near-uniform routing that hides
bugs.  `fill_state_dict_` overwrites every tensor of a state_dict with
values drawn from a generator seeded by the tensor's *key name*, so the reference model, the
oracle and this package all get bit-identical weights from nothing but the key names and shapes.
"""
from __future__ import annotations

import re
import zlib

import torch

_DET_CLS = re.compile(r"cv3\.\d+\.2\.(weight|bias)$")
_DET_BOX = re.compile(r"cv2\.\d+\.2\.(weight|bias)$")


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


@torch.no_grad()
def fill_state_dict_(sd: dict, seed: int = 0, conv_gain: float = 1.0) -> dict:
    """In-place deterministic fill.  Returns `sd` (tensors keep dtype/device/shape)."""
    for key in sorted(sd.keys()):
        t = sd[key]
        if not torch.is_tensor(t) or not t.is_floating_point() or t.numel() == 0:
            continue
        if key.endswith("dfl.conv.weight"):   # frozen arange(reg_max) in the reference (block.py:73-76): not a free parameter
            continue
        g = _gen(key, seed)
        shape = tuple(t.shape)
        head_cls = _DET_CLS.search(key) is not None   # Detect's last 1x1 (no norm after it): keep logits un-saturated
        head_box = _DET_BOX.search(key) is not None
        if t.dim() == 4:  # conv weight [Co, Ci/g, kh, kw]
            fan_in = shape[1] * shape[2] * shape[3]
            gain = 0.25 if head_cls else 0.5 if head_box else conv_gain
            v = torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)
        elif head_cls and key.endswith("bias"):
            v = -2.0 + 0.5 * torch.randn(shape, generator=g)
        elif head_box and key.endswith("bias"):
            v = 1.0 + 0.3 * torch.randn(shape, generator=g)
        elif t.dim() == 2:  # linear
            v = torch.randn(shape, generator=g) * (1.0 / shape[1] ** 0.5)
        elif key.endswith("running_var"):
            v = 0.5 + torch.rand(shape, generator=g)
        elif key.endswith("running_mean"):
            v = 0.2 * torch.randn(shape, generator=g)
        elif key.rsplit(".", 1)[-1] in ("ls1", "ls2", "ls_attn", "ls_ffn"):   # layer-scale (init 0.1 in the reference)
            v = 0.05 + 0.15 * torch.rand(shape, generator=g)
        elif key.endswith(".gamma"):
            v = 0.005 + 0.015 * torch.rand(shape, generator=g)
        elif key.endswith("weight"):  # norm scale
            v = 0.6 + 0.8 * torch.rand(shape, generator=g)
        elif key.endswith("bias"):
            v = 0.2 * torch.randn(shape, generator=g)
        else:
            continue
        t.copy_(v.to(t.dtype))
    return sd


def synth_images(batch: int, h: int, w: int, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    """Uniform [0,1) NCHW images (matches the predictor's /255 range, engine/predictor.py:175)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    return torch.rand((batch, 3, h, w), generator=g).to(dtype)


def load_norm_stats_(sd: dict, stats: dict) -> dict:
    """Overwrite BatchNorm running statistics with calibrated ones (tests/golden/*.bnstats.pt).

    The statistics were measured once on `synth_images` through the reference model (tests/golden/make_golden.py) so that,
    like a trained checkpoint, every BatchNorm sees the data distribution it normalises and activations stay O(1) in fp16."""
    for k, v in stats.items():
        if k not in sd:
            raise KeyError(f"calibrated statistic {k} has no counterpart in the state_dict")
        sd[k].copy_(v.to(sd[k].dtype))
    return sd
