"""Shared plumbing for the accelerated modules: layout conversion at the module boundary and weight-pack caching."""
from __future__ import annotations

import torch
import torch.nn as nn

BN_EPS = 1e-3  # every nn.BatchNorm2d on the path runs with eps=1e-3 (reference utils/torch_utils.py:552-562)


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """NCHW-logical tensor -> (B,H,W,C) fp16 view (zero-copy when x is channels_last fp16)."""
    if not x.is_cuda:
        raise RuntimeError(
            "yolo_master_b200 modules run on CUDA tensors only (hand-written sm_100a kernels; there is no CPU fallback)")
    if x.dtype != torch.float16:
        x = x.half()
    v = x.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


def to_nchw(y: torch.Tensor) -> torch.Tensor:
    """(B,H,W,C) -> NCHW-logical view (channels_last memory)."""
    return y.permute(0, 3, 1, 2)


def fold_bn(weight: torch.Tensor, conv_bias, bn: nn.BatchNorm2d | None):
    """fp32 (w', b') with BatchNorm folded in: reference fuse_conv_and_bn utils/torch_utils.py:315-349."""
    w = weight.detach().float()
    b = torch.zeros(w.shape[0], device=w.device) if conv_bias is None else conv_bias.detach().float()
    if bn is not None:
        s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * s.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return w, b


def pack_gemm_weight(w: torch.Tensor) -> torch.Tensor:
    """[Co,Ci,KH,KW] fp32 -> fp16 [Co, Kpad] with k = (ky*KW+kx)*Ci + ci, zero padded to a multiple of 32."""
    Co = w.shape[0]
    m = w.permute(0, 2, 3, 1).reshape(Co, -1)
    K = m.shape[1]
    Kpad = (K + 31) // 32 * 32
    out = torch.zeros((Co, Kpad), dtype=torch.float16, device=w.device)
    out[:, :K] = m.half()
    return out.contiguous()


def pad_channels8(w: torch.Tensor, b: torch.Tensor, cout_to: int = 8):
    """Zero-pad a folded conv weight [Co,Ci,KH,KW] / bias [Co] to multiples of 8 channels on both sides (1-class heads, the OBB angle,
    the 51-channel Pose towers): kernels move 16-byte (8-channel) vectors, and zero weights keep the pad channels at exactly zero.
    `cout_to` = 2 for a head's last layer (the GEMM kernel writes any even width; only an odd one needs a pad channel).
    Returns (w, b, cin_p, cout_p)."""
    Co, Ci = w.shape[0], w.shape[1]
    cin_p, cout_p = (Ci + 7) // 8 * 8, (Co + cout_to - 1) // cout_to * cout_to
    if cin_p != Ci or cout_p != Co:
        wp = torch.zeros((cout_p, cin_p, *w.shape[2:]), dtype=w.dtype, device=w.device)
        wp[:Co, :Ci] = w
        bp = torch.zeros(cout_p, dtype=b.dtype, device=b.device)
        bp[:Co] = b
        w, b = wp, bp
    return w, b.contiguous(), cin_p, cout_p


def plain_conv_run(m: nn.Conv2d, pk: dict, x: torch.Tensor, out=None, out_f32=False):
    """The GEMM kernel for a bare nn.Conv2d from its pack; a padded output width is cut back to a dense tensor (the decode kernels
    that follow read dense rows)."""
    from ... import ops
    padded = pk["cin_p"] != m.in_channels or pk["cout_p"] != m.out_channels
    if padded and (x.shape[-1] != pk["cin_p"] or out is not None):
        raise NotImplementedError(f"nn.Conv2d({m.in_channels}->{m.out_channels}): odd channel widths run zero-padded inside a tower only")
    y = ops.conv2d(x, pk["w"], pk["bias"], pk["cout_p"], m.kernel_size[0], m.kernel_size[1], m.stride[0], m.padding[0], False,
                   out=out, out_f32=out_f32)
    return y[..., :m.out_channels].contiguous() if pk["cout_p"] != m.out_channels else y


def bn_affine(bn: nn.BatchNorm2d):
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return s.contiguous(), (bn.bias.detach().float() - bn.running_mean.detach().float() * s).contiguous()


class PackCache:
    """Derived (folded / repacked / fp16) weights, rebuilt when any source tensor changes version, storage or device."""

    def _pack_sources(self):
        return list(self.parameters(recurse=True)) + list(self.buffers(recurse=True))

    def get_pack(self):
        key = tuple((t.data_ptr(), t._version, str(t.device)) for t in self._pack_sources())
        cached = self.__dict__.get("_ym_pack")
        if cached is None or cached[0] != key:
            with torch.no_grad():
                cached = (key, self._build_pack())
            self.__dict__["_ym_pack"] = cached
        return cached[1]

    def _build_pack(self):
        raise NotImplementedError


def cached_f32(owner: nn.Module, name: str, t: torch.Tensor, shape=None) -> torch.Tensor:
    """Contiguous fp32 copy of a small parameter (layer-scale, norm affine ...), rebuilt when the source changes."""
    key = (t.data_ptr(), t._version, str(t.device))
    slot = owner.__dict__.setdefault("_ym_f32", {})
    hit = slot.get(name)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            v = t.detach().float().reshape(shape if shape is not None else t.shape).contiguous()
        hit = (key, v)
        slot[name] = hit
    return hit[1]


def cached_pack(owner: nn.Module, name: str, sources, build):
    """`build()` result cached on `owner` until any tensor in `sources` changes version, storage or device."""
    key = tuple((t.data_ptr(), t._version, str(t.device)) for t in sources)
    slot = owner.__dict__.setdefault("_ym_packs", {})
    hit = slot.get(name)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            hit = (key, build())
        slot[name] = hit
    return hit[1]


def require_eval(m: nn.Module):
    if m.training:
        raise RuntimeError(
            f"{type(m).__name__}: the B200 path implements the inference forward only; call .eval() first "
            "(training is out of scope, SURVEY.md §8)")


def fwd_child(m: nn.Module, x: torch.Tensor, out=None, **kw):
    """`m.fwd_nhwc(x, out=...)` for a child that may be a plain `nn.Sequential` (instances built by the reference's own
    constructors inside `yolo_master_b200.integration`): the chain is walked here, the last member writes into `out`."""
    f = getattr(m, "fwd_nhwc", None)
    if f is not None:
        return f(x, out=out, **kw)
    if isinstance(m, nn.Sequential):
        mods = list(m)
        for j, mm in enumerate(mods):
            x = fwd_child(mm, x, out=out if j == len(mods) - 1 else None)
        return x
    raise NotImplementedError(f"{type(m).__name__} has no NHWC forward on the B200 path")


def plain_conv_fwd(m: nn.Conv2d, x: torch.Tensor, out=None, out_f32=False):
    """Bare `nn.Conv2d` (bias, no norm / activation: Detect's last 1x1, head.py:104-119) on the GEMM kernel; the folded fp16 pack
    is cached on the module (same cache as `PlainConv2d`, which mirrors this layer when the model is built by this package)."""
    from ... import ops
    f = getattr(m, "fwd_nhwc", None)
    if f is not None:
        return f(x, out=out, out_f32=out_f32)
    if m.groups != 1 or m.dilation != (1, 1) or m.kernel_size[0] != m.kernel_size[1] or m.stride[0] != m.stride[1]:
        raise NotImplementedError("nn.Conv2d: groups / dilation / non-square geometry is not on the B200 path")

    def build():
        w, b = fold_bn(m.weight, m.bias, None)
        w, b, cin_p, cout_p = pad_channels8(w, b.contiguous(), cout_to=2)
        return {"w": pack_gemm_weight(w), "bias": b, "cin_p": cin_p, "cout_p": cout_p}

    pk = cached_pack(m, "plain", [m.weight] + ([m.bias] if m.bias is not None else []), build)
    return plain_conv_run(m, pk, x, out, out_f32)


_SIDE_STREAMS: dict = {}


def run_branches(fns):
    """Run independent kernel chains.  While a CUDA graph is being captured each chain gets its own stream (forked from / joined to
    the capturing stream), so the graph holds them as parallel branches - the six towers of the Detect head are ~24 latency-bound
    launches that otherwise run back to back.  Eager execution stays on the current stream (no cross-stream allocator traffic)."""
    if len(fns) <= 1 or not torch.cuda.is_available() or not torch.cuda.is_current_stream_capturing():
        return [f() for f in fns]
    cur = torch.cuda.current_stream()
    key = (cur.device.index, len(fns) - 1)
    side = _SIDE_STREAMS.get(key)
    if side is None:
        side = [torch.cuda.Stream(device=cur.device) for _ in range(len(fns) - 1)]
        _SIDE_STREAMS[key] = side
    for s in side:
        s.wait_stream(cur)
    out = [None] * len(fns)
    for s, (j, f) in zip(side, list(enumerate(fns))[1:]):
        with torch.cuda.stream(s):
            out[j] = f()
    out[0] = fns[0]()
    for s in side:
        cur.wait_stream(s)
    return out
