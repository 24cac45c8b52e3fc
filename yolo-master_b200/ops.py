"""Tensor-level wrappers over the C ABI (include/ym_b200.h).

All activations are fp16 CUDA tensors of logical shape (B, H, W, C) whose channel dimension is dense
(stride 1) — either contiguous NHWC or a channel slice of a wider NHWC buffer (pitch = stride of W).
Nothing here falls back to torch: a missing library or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_L = None
CONV_IMPL = "tc"  # "tc": TMA + tcgen05 conv where supported (else the mma.sync implicit GEMM); "legacy": always mma.sync
LAUNCHES = 0  # number of kernel-launching C-ABI calls issued (bench.py reports kernels via LAUNCH_KERNELS)
KERNELS = 0   # number of device kernels launched (a call may launch several)


REQUIRE_CUDA = True   # tests/test_cuda_host_emu.py clears this to drive the host-emulated build of the library with CPU tensors


def _dev(t) -> bool:
    """True where a tensor is acceptable as device memory for the C ABI."""
    return t.is_cuda or not REQUIRE_CUDA


def lib():
    global _L
    if _L is None:
        _L = _lib.load()
    return _L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _count(nkernels: int = 1):
    global LAUNCHES, KERNELS
    LAUNCHES += 1
    KERNELS += nkernels


def pitch(t: torch.Tensor, dtype=torch.float16) -> int:
    """Row pitch (elements) of an NHWC activation view; validates the layout."""
    if t.dim() != 4 or t.dtype != dtype or not _dev(t):
        raise ValueError(f"expected a 4-D fp16 CUDA NHWC activation, got {tuple(t.shape)} {t.dtype} {t.device}")
    B, H, W, Cc = t.shape
    sb, sh, sw, sc = t.stride()
    if Cc > 1 and sc != 1:
        raise ValueError("NHWC activation must have a dense channel dimension")
    ld = sw if W > 1 else (sh if H > 1 else (sb if B > 1 else Cc))
    if (W > 1 and H > 1 and sh != W * ld) or (B > 1 and sb != H * W * ld):
        raise ValueError(f"activation is not a (sliced) contiguous NHWC tensor: shape {tuple(t.shape)} strides {t.stride()}")
    return ld


def new_act(B, H, W, Cc, device) -> torch.Tensor:
    return torch.empty((B, H, W, Cc), dtype=torch.float16, device=device)


def conv2d(x, w_packed, bias, Cout, KH, KW, stride, pad, act, out=None, res=None, out_f32=False):
    """ym_conv2d_nhwc.  x: (B,H,W,Cin) view.  Returns `out` (B,Ho,Wo,Cout)."""
    B, H, W, Cin = x.shape
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    if out is None:
        out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32 if out_f32 else torch.float16, device=x.device)
        ldo = Cout
    else:
        ldo = pitch(out, torch.float32 if out_f32 else torch.float16)
    if tuple(out.shape) != (B, Ho, Wo, Cout):
        raise ValueError(f"conv2d: out shape {tuple(out.shape)} != {(B, Ho, Wo, Cout)}")
    ldr = 0
    rp = None
    if res is not None:
        if tuple(res.shape) != (B, Ho, Wo, Cout):
            raise ValueError("conv2d: residual shape mismatch")
        ldr = pitch(res)
        rp = res.data_ptr()
    ldx = pitch(x)
    L = lib()
    fn, name = L.ym_conv2d_nhwc, "ym_conv2d_nhwc"
    # 3x3 convs over <= 16 input channels stay on the mma.sync implicit GEMM: a k-tile is then 128 rows of 32 bytes and the TMA
    # unit, fed one 32-byte row at a time for each of the 9 taps, is slower than 16-byte cp.async gathers (measured at bs32:
    # 93.8 vs 82.7 us for 16->32 s2 @320^2, 86.0 vs 62.3 us for 16->8 @160^2; profiles/r01_conv_impl_table.json).  The rule
    # depends on the layer only, never on the batch, so results stay bit-identical across batch sizes.
    small_k = KH == 3 and Cin <= 16
    if CONV_IMPL == "tc" and not small_k and Cout % 8 == 0 and L.ym_conv2d_tc_supported(Cin, Cout, KH, KW, stride, pad, ldx):
        fn, name = L.ym_conv2d_tc, "ym_conv2d_tc"
    _lib.check(fn(x.data_ptr(), ldx, B, H, W, Cin, w_packed.data_ptr(), w_packed.shape[1],
                  None if bias is None else bias.data_ptr(), Cout, KH, KW, stride, pad,
                  out.data_ptr(), ldo, 1 if out_f32 else 0, rp, ldr, 1 if act else 0, _stream()), name)
    _count()
    return out


def stem_conv(img, wgt, bias, Cout, out=None):
    """ym_stem_conv_nchw.  img: contiguous NCHW fp16/fp32/uint8; wgt / bias: contiguous fp32 HOST tensors."""
    if wgt.is_cuda or bias.is_cuda or wgt.dtype != torch.float32 or bias.dtype != torch.float32 or not wgt.is_contiguous():
        raise ValueError("stem_conv: weights and bias must be contiguous fp32 host tensors")
    if not img.is_cuda or img.dim() != 4:
        raise ValueError("stem_conv: expected a 4-D CUDA NCHW image batch")
    img = img.contiguous()
    dt = {torch.float16: 0, torch.float32: 1, torch.uint8: 2}.get(img.dtype)
    if dt is None:
        raise ValueError(f"stem_conv: unsupported image dtype {img.dtype}")
    B, Cin, H, W = img.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if out is None:
        out = new_act(B, Ho, Wo, Cout, img.device)
    _lib.check(lib().ym_stem_conv_nchw(img.data_ptr(), dt, B, Cin, H, W, wgt.data_ptr(), bias.data_ptr(), Cout,
                                       out.data_ptr(), pitch(out), _stream()), "ym_stem_conv_nchw")
    _count()
    return out


def dwconv(x, w_taps, bias, ksize, act, C_out, grp_w=None, grp_stride=None, grp_off=0, add=None, out=None):
    """ym_dwconv_nhwc.  x: (B,H,W,Cx) source view; output has C_out channels gathered per the group mapping."""
    B, H, W, _ = x.shape
    if grp_w is None:
        grp_w, grp_stride = C_out, C_out
    if out is None:
        out = new_act(B, H, W, C_out, x.device)
    _lib.check(lib().ym_dwconv_nhwc(x.data_ptr(), pitch(x), grp_w, grp_stride, grp_off, w_taps.data_ptr(),
                                    None if bias is None else bias.data_ptr(), B, H, W, C_out, ksize, 1 if act else 0,
                                    None if add is None else add.data_ptr(), 0 if add is None else pitch(add),
                                    out.data_ptr(), pitch(out), _stream()), "ym_dwconv_nhwc")
    _count()
    return out


def sppf_pool(buf, Cslot, k):
    B, H, W, Ct = buf.shape
    _lib.check(lib().ym_sppf_pool_nhwc(buf.data_ptr(), pitch(buf), B, H, W, Cslot, k, _stream()), "ym_sppf_pool_nhwc")
    _count()
    return buf


def concat2(a, b, up=1, out=None):
    """cat([upsample(a, up), b], channel).  b may be None (pure upsample / copy)."""
    Ba, Ha, Wa, Ca = a.shape
    H, W = Ha * up, Wa * up
    Cb = 0 if b is None else b.shape[3]
    if b is not None and tuple(b.shape[:3]) != (Ba, H, W):
        raise ValueError(f"concat2: spatial mismatch {tuple(a.shape)} x{up} vs {tuple(b.shape)}")
    if out is None:
        out = new_act(Ba, H, W, Ca + Cb, a.device)
    _lib.check(lib().ym_concat2_nhwc(a.data_ptr(), pitch(a), Ca, up, None if b is None else b.data_ptr(),
                                     8 if b is None else pitch(b), Cb, out.data_ptr(), pitch(out), Ba, H, W, _stream()),
               "ym_concat2_nhwc")
    _count()
    return out


def attention(qkv, batch, N, heads, head_stride, q_off, k_off, v_off, d_qk, d_v, scale, out=None):
    """ym_attention_fwd over the rows of `qkv` (B,H,W,Ctot) reinterpreted as (batch, N, Ctot)."""
    B, H, W, _ = qkv.shape
    if B * H * W != batch * N:
        raise ValueError("attention: batch*N must equal the number of token rows")
    if out is None:
        out = new_act(B, H, W, heads * d_v, qkv.device)
    _lib.check(lib().ym_attention_fwd(qkv.data_ptr(), pitch(qkv), batch, N, heads, head_stride, q_off, k_off, v_off, d_qk,
                                      d_v, float(scale), out.data_ptr(), pitch(out), _stream()), "ym_attention_fwd")
    _count()
    return out


def router_topk(x, pack, topk, pool=4):
    """ym_router_topk.  Returns (idx int32 [B,k], w fp32 [B,k], probs fp32 [B,E])."""
    B, H, W, Cc = x.shape
    Cr, E = pack["Cr"], pack["E"]
    n = lib().ym_router_scratch_floats(B, H, W, Cc, Cr, pool)
    scratch = torch.empty((n,), dtype=torch.float32, device=x.device)
    idx = torch.empty((B, topk), dtype=torch.int32, device=x.device)
    w = torch.empty((B, topk), dtype=torch.float32, device=x.device)
    probs = torch.empty((B, E), dtype=torch.float32, device=x.device)
    _lib.check(lib().ym_router_topk(x.data_ptr(), pitch(x), B, H, W, Cc, pool, pack["w1"].data_ptr(), Cr,
                                    pack["scale1"].data_ptr(), pack["shift1"].data_ptr(), pack["w2"].data_ptr(),
                                    pack["scale2"].data_ptr(), pack["shift2"].data_ptr(), E, topk, scratch.data_ptr(),
                                    idx.data_ptr(), w.data_ptr(), probs.data_ptr(), _stream()), "ym_router_topk")
    _count(2)
    return idx, w, probs


def router_partial(x, pack, pool=4):
    """ym_router_partial: the fused pool + conv + BN + SiLU pass of the router.  Returns (partial sums, nblk, npix) for
    `moe_ffn_stats_routed`, whose prologue finishes the routing."""
    B, H, W, Cc = x.shape
    Cr = pack["Cr"]
    n = lib().ym_router_scratch_floats(B, H, W, Cc, Cr, pool)
    scratch = torch.empty((n,), dtype=torch.float32, device=x.device)
    npix = C.c_int(0)
    nblk = lib().ym_router_blocks(H, W, pool, C.addressof(npix))
    _lib.check(lib().ym_router_partial(x.data_ptr(), pitch(x), B, H, W, Cc, pool, pack["w1"].data_ptr(), Cr, pack["scale1"].data_ptr(),
                                       pack["shift1"].data_ptr(), scratch.data_ptr(), _stream()), "ym_router_partial")
    _count()
    return scratch, nblk, npix.value


def moe_ffn_stats_routed(x, topk, w1, rpack, partial, nblk, npix):
    """ym_moe_ffn_routed: stage 1 of the expert FFN with the router's finish in its prologue.
    Returns (idx int32 [B,k], w fp32 [B,k], probs fp32 [B,E], GroupNorm-1 partial sums, strips)."""
    B, H, W, Cc = x.shape
    E, HID, _ = w1.shape
    P = B * topk
    strips = lib().ym_moe_ffn_strips(H * W, P)
    stats = torch.empty((lib().ym_moe_ffn_stats_floats(P, strips, HID),), dtype=torch.float32, device=x.device)
    idx = torch.empty((B, topk), dtype=torch.int32, device=x.device)
    w = torch.empty((B, topk), dtype=torch.float32, device=x.device)
    probs = torch.empty((B, E), dtype=torch.float32, device=x.device)
    _lib.check(lib().ym_moe_ffn_routed(x.data_ptr(), pitch(x), B, H * W, Cc, HID, topk, w1.data_ptr(), E, partial.data_ptr(), nblk,
                                       rpack["Cr"], npix, rpack["w2"].data_ptr(), rpack["scale2"].data_ptr(), rpack["shift2"].data_ptr(),
                                       idx.data_ptr(), w.data_ptr(), probs.data_ptr(), stats.data_ptr(), strips, _stream()),
               "ym_moe_ffn_routed")
    _count()
    return idx, w, probs, stats, strips


def moe_expert_gemm(a, lda, a_div, P, HW, K, w_all, route_idx, N, a_scale=None, a_shift=None, groups=0):
    """ym_moe_expert_gemm.  Returns (out fp16 [P,HW,N], partial GroupNorm stats fp32 or None)."""
    dev = route_idx.device
    out = torch.empty((P, HW, N), dtype=torch.float16, device=dev)
    stats = torch.empty((lib().ym_moe_stats_floats(P, HW, N),), dtype=torch.float32, device=dev) if groups else None
    E, Nw, Kpad = w_all.shape
    _lib.check(lib().ym_moe_expert_gemm(a.data_ptr(), lda, a_div, P, HW, K, w_all.data_ptr(), Kpad, Nw * Kpad,
                                        route_idx.data_ptr(), N, out.data_ptr(), N,
                                        None if a_scale is None else a_scale.data_ptr(),
                                        None if a_shift is None else a_shift.data_ptr(),
                                        None if stats is None else stats.data_ptr(), groups, _stream()),
               "ym_moe_expert_gemm")
    _count()
    return out, stats


def gn_finalize(stats, P, HW, G, C_, count, eps, gamma, beta, route_idx, route_w=None):
    scale = torch.empty((P, C_), dtype=torch.float32, device=stats.device)
    shift = torch.empty((P, C_), dtype=torch.float32, device=stats.device)
    _lib.check(lib().ym_gn_finalize(stats.data_ptr(), P, HW, G, C_, float(count), float(eps), gamma.data_ptr(), beta.data_ptr(),
                                    route_idx.data_ptr(), None if route_w is None else route_w.data_ptr(),
                                    scale.data_ptr(), shift.data_ptr(), _stream()), "ym_gn_finalize")
    _count()
    return scale, shift


def moe_ffn_supported(C, HID, ldx) -> bool:
    """True when the tcgen05 expert-FFN kernels (csrc/tc_moe.cu) serve this width pair."""
    return bool(lib().ym_moe_ffn_supported(C, HID, ldx))


def moe_ffn_stats(x, topk, w1, route_idx):
    """ym_moe_ffn stage 1.  x: (B,H,W,C) fp16 view; w1 fp16 [E][HID][C]; route_idx int32 [B*topk].
    Returns (GroupNorm-1 partial sums, strips) - the hidden activation itself is never written."""
    B, H, W, Cc = x.shape
    E, HID, _ = w1.shape
    P = B * topk
    strips = lib().ym_moe_ffn_strips(H * W, P)
    stats = torch.empty((lib().ym_moe_ffn_stats_floats(P, strips, HID),), dtype=torch.float32, device=x.device)
    _lib.check(lib().ym_moe_ffn(1, x.data_ptr(), pitch(x), B, H * W, Cc, HID, topk, w1.data_ptr(), None, E, route_idx.data_ptr(),
                                None, None, None, stats.data_ptr(), strips, _stream()), "ym_moe_ffn(stats)")
    _count()
    return stats, strips


def moe_ffn_fused(x, topk, w1, w2, route_idx, a_scale, a_shift, strips):
    """ym_moe_ffn stage 2: o[p] = SiLU(GN1(x W1[e]^T)) W2[e]^T, fp16 [P][HW][C], + GroupNorm-2 partial sums."""
    B, H, W, Cc = x.shape
    E, HID, _ = w1.shape
    P = B * topk
    o = torch.empty((P, H * W, Cc), dtype=torch.float16, device=x.device)
    stats = torch.empty((lib().ym_moe_ffn_stats_floats(P, strips, Cc),), dtype=torch.float32, device=x.device)
    _lib.check(lib().ym_moe_ffn(2, x.data_ptr(), pitch(x), B, H * W, Cc, HID, topk, w1.data_ptr(), w2.data_ptr(), E, route_idx.data_ptr(),
                                a_scale.data_ptr(), a_shift.data_ptr(), o.data_ptr(), stats.data_ptr(), strips, _stream()),
               "ym_moe_ffn(fused)")
    _count()
    return o, stats


def moe_ffn_fused_gn(x, topk, w1, w2, route_idx, st1, strips, G1, count1, eps1, gamma1, beta1):
    """ym_moe_ffn_gn: stage 2 with GroupNorm-1 finalised inside the kernel from stage 1's partial sums `st1`."""
    B, H, W, Cc = x.shape
    E, HID, _ = w1.shape
    P = B * topk
    o = torch.empty((P, H * W, Cc), dtype=torch.float16, device=x.device)
    stats = torch.empty((lib().ym_moe_ffn_stats_floats(P, strips, Cc),), dtype=torch.float32, device=x.device)
    _lib.check(lib().ym_moe_ffn_gn(x.data_ptr(), pitch(x), B, H * W, Cc, HID, topk, w1.data_ptr(), w2.data_ptr(), E, route_idx.data_ptr(),
                                   st1.data_ptr(), G1, float(count1), float(eps1), gamma1.data_ptr(), beta1.data_ptr(), o.data_ptr(),
                                   stats.data_ptr(), strips, _stream()), "ym_moe_ffn_gn")
    _count()
    return o, stats


def moe_combine_gn_supported(x, ws_packed, o, topk, out=None) -> bool:
    Cc = x.shape[3]
    return (MOE_COMBINE_IMPL == "tc" and topk <= 2 and ws_packed.shape[1] == Cc and o.shape[2] == Cc
            and bool(lib().ym_moe_combine_tc_supported(Cc, pitch(x), pitch(x) if out is None else pitch(out))))


def moe_combine_gn(x, ws_packed, bias_s, o, st2, strips, G2, count2, eps2, gamma2, beta2, route_idx, route_w, topk, add_residual=True, out=None):
    """ym_moe_combine_tc_gn: the combine with GroupNorm-2 (times the routing weight) finalised inside the kernel from `st2`."""
    B, H, W, Cc = x.shape
    if out is None:
        out = new_act(B, H, W, Cc, x.device)
    _lib.check(lib().ym_moe_combine_tc_gn(x.data_ptr(), pitch(x), B, H * W, Cc, ws_packed.data_ptr(), bias_s.data_ptr(), o.data_ptr(),
                                          st2.data_ptr(), strips, G2, float(count2), float(eps2), gamma2.data_ptr(), beta2.data_ptr(),
                                          route_idx.data_ptr(), route_w.data_ptr(), topk, out.data_ptr(), pitch(out),
                                          1 if add_residual else 0, _stream()), "ym_moe_combine_tc_gn")
    _count()
    return out


def gn_finalize_tiles(stats, P, tiles, G, C_, count, eps, gamma, beta, route_idx, route_w=None):
    """ym_gn_finalize_tiles: partial sums [P][tiles][C_/8][2] -> (scale, shift) fp32 [P][C_] of the routed expert's GroupNorm."""
    scale = torch.empty((P, C_), dtype=torch.float32, device=stats.device)
    shift = torch.empty((P, C_), dtype=torch.float32, device=stats.device)
    _lib.check(lib().ym_gn_finalize_tiles(stats.data_ptr(), P, tiles, G, C_, float(count), float(eps), gamma.data_ptr(), beta.data_ptr(),
                                          route_idx.data_ptr(), None if route_w is None else route_w.data_ptr(),
                                          scale.data_ptr(), shift.data_ptr(), _stream()), "ym_gn_finalize_tiles")
    _count()
    return scale, shift


MOE_COMBINE_IMPL = "tc"    # "tc": tcgen05 kernel (csrc/tc_moe.cu) where the width allows; "mma": the mma.sync kernel of csrc/gemm_conv.cu


def moe_combine(x, ws_packed, bias_s, o, o_scale, o_shift, topk, add_residual=True, out=None):
    B, H, W, Cc = x.shape
    if out is None:
        out = new_act(B, H, W, Cc, x.device)
    L = lib()
    if MOE_COMBINE_IMPL == "tc" and topk <= 2 and ws_packed.shape[1] == Cc and o.shape[2] == Cc and L.ym_moe_combine_tc_supported(Cc, pitch(x), pitch(out)):
        _lib.check(L.ym_moe_combine_tc(x.data_ptr(), pitch(x), B, H * W, Cc, ws_packed.data_ptr(), bias_s.data_ptr(), o.data_ptr(),
                                       o_scale.data_ptr(), o_shift.data_ptr(), topk, out.data_ptr(), pitch(out),
                                       1 if add_residual else 0, _stream()), "ym_moe_combine_tc")
        _count()
        return out
    _lib.check(L.ym_moe_combine(x.data_ptr(), pitch(x), B, H * W, Cc, ws_packed.data_ptr(), ws_packed.shape[1],
                                bias_s.data_ptr(), o.data_ptr(), o.shape[2], o_scale.data_ptr(), o_shift.data_ptr(),
                                topk, out.data_ptr(), pitch(out), 1 if add_residual else 0, _stream()),
               "ym_moe_combine")
    _count()
    return out


def _level_arrays(boxes, logits, strides):
    nl = len(boxes)
    VP = C.c_void_p * nl
    bp = VP(*[b.data_ptr() for b in boxes])
    cp = VP(*[c.data_ptr() for c in logits])
    hs = (C.c_int * nl)(*[b.shape[1] for b in boxes])
    ws = (C.c_int * nl)(*[b.shape[2] for b in boxes])
    st = (C.c_float * nl)(*[float(s) for s in strides])
    return nl, bp, cp, hs, ws, st


def detect_topk(boxes, logits, strides, nc, max_det=300, return_anchor=False):
    """boxes[l]: fp32 (B,h,w,4); logits[l]: fp32 (B,h,w,nc).  Returns (B,k,6) fp32 [+ anchor ids (B,k) int32]."""
    B = boxes[0].shape[0]
    A = sum(b.shape[1] * b.shape[2] for b in boxes)
    k = min(max_det, A)
    out = torch.empty((B, k, 6), dtype=torch.float32, device=boxes[0].device)
    anc = torch.empty((B, k), dtype=torch.int32, device=boxes[0].device) if return_anchor else None
    nl, bp, cp, hs, ws, st = _level_arrays(boxes, logits, strides)
    scratch = torch.empty((B * A,), dtype=torch.int32, device=boxes[0].device)
    _lib.check(lib().ym_detect_topk(nl, bp, cp, hs, ws, st, B, nc, max_det, out.data_ptr(),
                                    None if anc is None else anc.data_ptr(), scratch.data_ptr(), _stream()), "ym_detect_topk")
    _count(2)
    return (out, anc) if return_anchor else out


def detect_dense(boxes, logits, strides, nc, xyxy, reg_max=1):
    """boxes[l]: fp32 (B,h,w,4*reg_max) (DFL bin logits when reg_max > 1); returns y fp32 (B, 4+nc, A)."""
    B = boxes[0].shape[0]
    A = sum(b.shape[1] * b.shape[2] for b in boxes)
    y = torch.empty((B, 4 + nc, A), dtype=torch.float32, device=boxes[0].device)
    nl, bp, cp, hs, ws, st = _level_arrays(boxes, logits, strides)
    _lib.check(lib().ym_detect_dense(nl, bp, cp, hs, ws, st, B, nc, reg_max, 1 if xyxy else 0, y.data_ptr(), _stream()),
               "ym_detect_dense")
    _count()
    return y


def tc_gemm_nt(a, b, bias=None, res=None, act=False, out=None):
    """ym_tc_gemm_nt (tcgen05): a [M,K] fp16 (row pitch = stride(0)), b [N,K] fp16.  Returns out [M,N] fp16."""
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float16, device=a.device)
    _lib.check(lib().ym_tc_gemm_nt(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                   None if bias is None else bias.data_ptr(), None if res is None else res.data_ptr(),
                                   0 if res is None else res.stride(0), out.data_ptr(), out.stride(0), M, N, K,
                                   1 if act else 0, _stream()), "ym_tc_gemm_nt")
    _count()
    return out


# "v3": CTA-pair kernel (tcgen05.mma.cta_group::2, half the weight traffic per SM) where supported, else "v2": persistent
# single-CTA TMA + tcgen05 kernel where supported, else "v1": one tile per CTA (cp.async + tcgen05)
DISPATCH_IMPL = "v3"


def moe_dispatch(x, w_all, route_idx, route_w, w_min=0.01, clamp=1e4, out=None):
    """ES-MoE dispatch.  x: (B,H,W,C) fp16 NHWC; w_all: [E,N,C] fp16; route_idx int32 [B,k]; route_w fp32 [B,k]."""
    B, H, W, Cc = x.shape
    E, N, Kw = w_all.shape
    k = route_idx.shape[1]
    if out is None:
        out = new_act(B, H, W, N, x.device)
    L = lib()
    if DISPATCH_IMPL == "v3" and w_all.is_contiguous() and L.ym_moe_dispatch_v3_supported(H * W, Cc, N, k, pitch(x), w_all.stride(1), pitch(out)):
        _lib.check(L.ym_moe_dispatch_v3(x.data_ptr(), pitch(x), B, H * W, Cc, w_all.data_ptr(), w_all.stride(1), E, route_idx.data_ptr(),
                                        route_w.data_ptr(), k, N, float(w_min), float(clamp), out.data_ptr(), pitch(out), _stream()),
                   "ym_moe_dispatch_v3")
        _count()
        return out
    if DISPATCH_IMPL in ("v2", "v3") and w_all.is_contiguous() and L.ym_moe_dispatch_v2_supported(H * W, Cc, N, k, pitch(x), w_all.stride(1), pitch(out)):
        _lib.check(L.ym_moe_dispatch_v2(x.data_ptr(), pitch(x), B, H * W, Cc, w_all.data_ptr(), w_all.stride(1), E, route_idx.data_ptr(),
                                        route_w.data_ptr(), k, N, float(w_min), float(clamp), out.data_ptr(), pitch(out), _stream()),
                   "ym_moe_dispatch_v2")
        _count()
        return out
    _lib.check(lib().ym_moe_dispatch_tc(x.data_ptr(), pitch(x), B, H * W, Cc, w_all.data_ptr(), w_all.stride(1), w_all.stride(0),
                                        route_idx.data_ptr(), route_w.data_ptr(), k, N, float(w_min), float(clamp),
                                        out.data_ptr(), pitch(out), _stream()), "ym_moe_dispatch_tc")
    _count()
    return out


def nms_batched(pred, conf_thres, iou_thres, max_det=300, max_nms=30000, max_wh=7680.0, mode=0, sigma=0.1, frame_wh=(0.0, 0.0)):
    """ym_nms_batched.  pred: fp32 (B, 4+nc, A).  Returns (out (B,max_det,6), count (B,) int32, idx (B,max_det) int32)."""
    if pred.dtype != torch.float32 or not pred.is_cuda or pred.dim() != 3:
        raise ValueError("nms_batched: expected an fp32 CUDA tensor of shape (B, 4+nc, A)")
    pred = pred.contiguous()
    B, no, A = pred.shape
    out = torch.empty((B, max_det, 6), dtype=torch.float32, device=pred.device)
    cnt = torch.empty((B,), dtype=torch.int32, device=pred.device)
    idx = torch.empty((B, max_det), dtype=torch.int32, device=pred.device)
    scratch = torch.empty((lib().ym_nms_scratch_bytes(B, A),), dtype=torch.uint8, device=pred.device)
    _lib.check(lib().ym_nms_batched(pred.data_ptr(), B, no - 4, A, float(conf_thres), float(iou_thres), max_det, max_nms,
                                    float(max_wh), mode, float(sigma), float(frame_wh[0]), float(frame_wh[1]), out.data_ptr(),
                                    cnt.data_ptr(), idx.data_ptr(), scratch.data_ptr(), _stream()), "ym_nms_batched")
    _count(2)
    return out, cnt, idx, scratch


def nms_batched_large(pred, conf_thres, iou_thres, max_det=300, max_nms=30000, max_wh=7680.0):
    """ym_nms_batched_large: mode 0 of nms_batched without the 16384-candidate limit.  Returns (out, count, idx)."""
    if pred.dtype != torch.float32 or not _dev(pred) or pred.dim() != 3:
        raise ValueError("nms_batched_large: expected an fp32 CUDA tensor of shape (B, 4+nc, A)")
    pred = pred.contiguous()
    B, no, A = pred.shape
    out = torch.empty((B, max_det, 6), dtype=torch.float32, device=pred.device)
    cnt = torch.empty((B,), dtype=torch.int32, device=pred.device)
    idx = torch.empty((B, max_det), dtype=torch.int32, device=pred.device)
    scratch = torch.empty((lib().ym_nms_large_scratch_bytes(B, A),), dtype=torch.uint8, device=pred.device)
    _lib.check(lib().ym_nms_batched_large(pred.data_ptr(), B, no - 4, A, float(conf_thres), float(iou_thres), max_det, max_nms,
                                          float(max_wh), out.data_ptr(), cnt.data_ptr(), idx.data_ptr(), scratch.data_ptr(), _stream()),
               "ym_nms_batched_large")
    _count(2)
    return out, cnt, idx


def dwconv3_routed(x, w_taps, route, dil):
    """ym_dwconv3_routed_nhwc.  x: (B,H,W,C) fp16 view; w_taps fp16 [E,9,C]; route int32 (B,) view (image -> expert); dil int32 [E]."""
    B, H, W, Cc = x.shape
    E = w_taps.shape[0]
    if route.dtype != torch.int32 or route.dim() != 1 or route.shape[0] != B or dil.dtype != torch.int32 or dil.numel() != E:
        raise ValueError("dwconv3_routed: route int32 (B,), dil int32 (E,)")
    out = new_act(B, H, W, Cc, x.device)
    _lib.check(lib().ym_dwconv3_routed_nhwc(x.data_ptr(), pitch(x), w_taps.data_ptr(), route.data_ptr(), max(route.stride(0), 1),
                                            dil.data_ptr(), E, B, H, W, Cc, out.data_ptr(), pitch(out), _stream()),
               "ym_dwconv3_routed_nhwc")
    _count()
    return out


def route_affine(scale, shift, gamma, beta, route, route_w=None):
    """ym_route_affine, in place: (scale, shift) fp32 (B, C) from unit-gamma GroupNorm statistics -> expert route[b]'s affine
    (times the routing weight route_w[b] when given)."""
    B, Cc = scale.shape
    _lib.check(lib().ym_route_affine(scale.data_ptr(), shift.data_ptr(), gamma.data_ptr(), beta.data_ptr(), route.data_ptr(),
                                     max(route.stride(0), 1), gamma.shape[0], B, Cc,
                                     None if route_w is None else route_w.data_ptr(), _stream()), "ym_route_affine")
    _count()
    return scale, shift


def process_mask(protos, dets, shape, upsample=True, coef_col=6):
    """ym_process_mask.  protos: (nm, mh, mw) fp16 / fp32; dets: fp32 (n, >= coef_col + nm) rows with the xyxy box (in `shape`
    coordinates) at columns 0..3 and the mask coefficients from `coef_col` -> uint8 (n, *shape) when upsample else (n, mh, mw)."""
    if protos.dim() != 3 or protos.dtype not in (torch.float16, torch.float32) or not _dev(protos):
        raise ValueError("process_mask: expected fp16 / fp32 CUDA prototypes of shape (nm, mh, mw)")
    if dets.dim() != 2 or dets.dtype != torch.float32 or not _dev(dets) or dets.stride(-1) != 1:
        raise ValueError("process_mask: expected fp32 CUDA detection rows (n, >= coef_col + nm)")
    nm, mh, mw = protos.shape
    n = dets.shape[0]
    oh, ow = (int(shape[0]), int(shape[1])) if upsample else (mh, mw)
    out = torch.empty((n, oh, ow), dtype=torch.uint8, device=protos.device)
    if n == 0:
        return out
    protos = protos.contiguous()
    scratch = torch.empty((lib().ym_process_mask_scratch_bytes(n, mh, mw),), dtype=torch.uint8, device=protos.device)
    _lib.check(lib().ym_process_mask(protos.data_ptr(), 1 if protos.dtype == torch.float16 else 2, nm, mh, mw, dets.data_ptr(),
                                     dets.stride(0), n, coef_col, int(shape[0]), int(shape[1]), 1 if upsample else 0, out.data_ptr(),
                                     scratch.data_ptr(), _stream()), "ym_process_mask")
    _count(2)
    return out


def nms_rotated(pred, conf_thres, iou_thres, max_det=300, max_nms=30000, max_wh=7680.0):
    """ym_nms_rotated.  pred: fp32 (B, 4+nc+1, A) xywh + class scores + angle.  Returns (out (B,max_det,7), count (B,) int32,
    idx (B,max_det) int32)."""
    if pred.dtype != torch.float32 or not _dev(pred) or pred.dim() != 3 or pred.shape[1] < 6:
        raise ValueError("nms_rotated: expected an fp32 CUDA tensor of shape (B, 4+nc+1, A)")
    pred = pred.contiguous()
    B, no, A = pred.shape
    out = torch.empty((B, max_det, 7), dtype=torch.float32, device=pred.device)
    cnt = torch.empty((B,), dtype=torch.int32, device=pred.device)
    idx = torch.empty((B, max_det), dtype=torch.int32, device=pred.device)
    scratch = torch.empty((lib().ym_nms_rotated_scratch_bytes(B, A),), dtype=torch.uint8, device=pred.device)
    _lib.check(lib().ym_nms_rotated(pred.data_ptr(), B, no - 5, A, float(conf_thres), float(iou_thres), max_det, max_nms, float(max_wh),
                                    out.data_ptr(), cnt.data_ptr(), idx.data_ptr(), scratch.data_ptr(), _stream()), "ym_nms_rotated")
    _count(4)
    return out, cnt, idx


def esmoe_forward(x, pack, topk, dyn_thr, out=None):
    """ES_MOE eval forward on the C ABI (ym_esmoe_route / _dwconv / _pointwise / _combine).  x: (B,H,W,C) fp16 view."""
    B, H, W, Cc = x.shape
    HW, E, N = H * W, pack["E"], pack["N"]
    dev = x.device
    L = lib()
    st = _stream()
    scratch = torch.empty((L.ym_esmoe_scratch_floats(B, HW, Cc),), dtype=torch.float32, device=dev)
    idx = torch.empty((B, topk), dtype=torch.int32, device=dev)
    w = torch.empty((B, topk), dtype=torch.float32, device=dev)
    probs = torch.empty((B, E), dtype=torch.float32, device=dev)
    _lib.check(L.ym_esmoe_route(x.data_ptr(), pitch(x), B, HW, Cc, pack["rw1"].data_ptr(), pack["rb1"].data_ptr(), pack["Cr"],
                                pack["rw2"].data_ptr(), pack["rb2"].data_ptr(), E, topk, float(dyn_thr), scratch.data_ptr(),
                                idx.data_ptr(), w.data_ptr(), probs.data_ptr(), st), "ym_esmoe_route")
    _count(2)
    t = torch.empty((B * topk, HW, Cc), dtype=torch.float16, device=dev)
    for e, (k, wdw) in enumerate(zip(pack["ks"], pack["dw"])):
        _lib.check(L.ym_esmoe_dwconv(x.data_ptr(), pitch(x), wdw.data_ptr(), B, H, W, Cc, k, idx.data_ptr(), topk, e,
                                     t.data_ptr(), Cc, st), "ym_esmoe_dwconv")
        _count()
    y = torch.empty((B * topk, HW, N), dtype=torch.float16, device=dev)
    pw = pack["pw"]
    _lib.check(L.ym_esmoe_pointwise(t.data_ptr(), Cc, B * topk, HW, Cc, pw.data_ptr(), pw.shape[2], pw.shape[1] * pw.shape[2],
                                    pack["pb"].data_ptr(), N, idx.data_ptr(), w.data_ptr(), y.data_ptr(), N, st), "ym_esmoe_pointwise")
    _count()
    if out is None:
        out = new_act(B, H, W, N, dev)
    _lib.check(L.ym_esmoe_combine(y.data_ptr(), N, idx.data_ptr(), topk, pack["fscale"].data_ptr(), pack["fshift"].data_ptr(),
                                  out.data_ptr(), pitch(out), B, HW, N, st), "ym_esmoe_combine")
    _count()
    return out, idx, w, probs


EW_SCALE_RES, EW_TOKEN_ACC, EW_GLU, EW_GELU, EW_AFFINE, EW_LERP = 0, 1, 2, 3, 4, 5


def _rows(t):
    return t.shape[0] * t.shape[1] * t.shape[2]


def ew(op, a=None, b=None, p0=None, p1=None, ldt=0, toff=0, rows_per_img=1, act=False, out=None, tok=None):
    """ym_ew_nhwc on (B,H,W,C) views; see include/ym_b200.h for the op table."""
    ref = a if a is not None else b
    B, H, W, Cc = ref.shape
    if out is None:
        out = new_act(B, H, W, Cc, ref.device)
    for p in (p0, p1, tok):
        if p is not None and (p.dtype != torch.float32 or not p.is_contiguous()):
            raise ValueError("ew: parameters must be contiguous fp32")
    _lib.check(lib().ym_ew_nhwc(op, None if a is None else a.data_ptr(), 0 if a is None else pitch(a),
                                None if b is None else b.data_ptr(), 0 if b is None else pitch(b),
                                None if p0 is None else p0.data_ptr(), None if p1 is None else p1.data_ptr(),
                                None if tok is None else tok.data_ptr(), ldt, toff, rows_per_img, 1 if act else 0, out.data_ptr(), pitch(out), B * H * W, Cc, _stream()), "ym_ew_nhwc")
    _count()
    return out


def groupnorm_stats(x, G, gamma, beta, eps=1e-5):
    """ym_groupnorm_stats over a (B,H,W,C) fp16 view -> (scale, shift) fp32 (B, C)."""
    B, H, W, Cc = x.shape
    sc = torch.empty((B, Cc), dtype=torch.float32, device=x.device)
    sh = torch.empty((B, Cc), dtype=torch.float32, device=x.device)
    _lib.check(lib().ym_groupnorm_stats(x.data_ptr(), 0, pitch(x), B, H * W, Cc, G, eps, gamma.data_ptr(), beta.data_ptr(),
                                        sc.data_ptr(), sh.data_ptr(), _stream()), "ym_groupnorm_stats")
    _count()
    return sc, sh


def groupnorm(x, G, gamma, beta, eps=1e-5, act=False, tok=None, ldt=0, toff=0, add=None, out=None):
    """GroupNorm (+SiLU) (* per-token weight) (+ add): statistics kernel + fused apply."""
    sc, sh = groupnorm_stats(x, G, gamma, beta, eps)
    return ew(EW_AFFINE, a=x, b=add, p0=sc, p1=sh, tok=tok, ldt=ldt, toff=toff, rows_per_img=x.shape[1] * x.shape[2], act=act, out=out)


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    B, H, W, Cc = x.shape
    if out is None:
        out = new_act(B, H, W, Cc, x.device)
    _lib.check(lib().ym_layernorm_nhwc(x.data_ptr(), pitch(x), gamma.data_ptr(), beta.data_ptr(), eps, out.data_ptr(), pitch(out),
                                       B * H * W, Cc, _stream()), "ym_layernorm_nhwc")
    _count()
    return out


def attn_small(q, k, v, heads, hdp, scale, out=None):
    """q: (B,H,W,heads*hdp) view; k, v: (B,h,w,heads*hdp) views (Nkv = h*w tokens per image)."""
    B, H, W, _ = q.shape
    Nkv = k.shape[1] * k.shape[2]
    if out is None:
        out = new_act(B, H, W, heads * hdp, q.device)
    _lib.check(lib().ym_attn_small(q.data_ptr(), pitch(q), k.data_ptr(), pitch(k), v.data_ptr(), pitch(v), B, heads, hdp, H * W, Nkv,
                                   scale, out.data_ptr(), pitch(out), _stream()), "ym_attn_small")
    _count()
    return out


def attn_window(q, k, v, heads, hdp, win, shift, scale, padk=None, padv=None, out=None):
    B, H, W, _ = q.shape
    if out is None:
        out = new_act(B, H, W, heads * hdp, q.device)
    _lib.check(lib().ym_attn_window(q.data_ptr(), pitch(q), k.data_ptr(), pitch(k), v.data_ptr(), pitch(v), B, H, W, heads, hdp, win,
                                    shift, None, None if padk is None else padk.data_ptr(), None if padv is None else padv.data_ptr(),
                                    scale, out.data_ptr(), pitch(out), _stream()), "ym_attn_window")
    _count()
    return out


def deform_sample(oa, v, heads, hd, n_points, align_corners=True, out=None):
    """oa: fp32 (B,H,W,heads*n_points*3); v: fp16 (B,H,W,heads*hd)."""
    B, H, W, Cc = v.shape
    if out is None:
        out = new_act(B, H, W, Cc, v.device)
    _lib.check(lib().ym_deform_sample(oa.data_ptr(), pitch(oa, torch.float32), v.data_ptr(), pitch(v), B, H, W, heads, hd, n_points,
                                      1 if align_corners else 0, out.data_ptr(), pitch(out), _stream()), "ym_deform_sample")
    _count()
    return out


def token_router(x, pk, topk, temp_dev=None, temp=1.0, want_idx=True):
    """ym_token_router.  pk: dict(w1 [HID,C], gn_w, gn_b, G, w2 [E,HID], b2) fp32.  Returns (weights fp32 (B,H,W,E), idx int32)."""
    B, H, W, Cc = x.shape
    HID, E = pk["w1"].shape[0], pk["w2"].shape[0]
    wts = torch.empty((B, H, W, E), dtype=torch.float32, device=x.device)
    idx = torch.empty((B, H, W, topk), dtype=torch.int32, device=x.device) if want_idx else None
    scratch = torch.empty((lib().ym_token_router_scratch_floats(B, H * W, HID),), dtype=torch.float32, device=x.device)
    _lib.check(lib().ym_token_router(x.data_ptr(), pitch(x), B, H * W, Cc, pk["w1"].data_ptr(), HID, pk["G"], pk["gn_w"].data_ptr(),
                                     pk["gn_b"].data_ptr(), 1e-5, pk["w2"].data_ptr(), pk["b2"].data_ptr(), E, topk,
                                     None if temp_dev is None else temp_dev.data_ptr(), float(temp), wts.data_ptr(),
                                     None if idx is None else idx.data_ptr(), scratch.data_ptr(), _stream()), "ym_token_router")
    _count(3)
    return wts, idx


def linear_attn(q, k, v, heads, hdp, hd, rf, eps=1e-6, limit=1e4, out=None):
    B, H, W, _ = q.shape
    N = H * W
    if out is None:
        out = new_act(B, H, W, heads * hdp, q.device)
    scratch = torch.empty((lib().ym_linear_attn_scratch_floats(B, heads, hdp, N),), dtype=torch.float32, device=q.device)
    _lib.check(lib().ym_linear_attn(q.data_ptr(), pitch(q), k.data_ptr(), pitch(k), v.data_ptr(), pitch(v), B, heads, hdp, hd,
                                    rf.shape[0], N, rf.data_ptr(), eps, limit, scratch.data_ptr(), out.data_ptr(), pitch(out),
                                    _stream()), "ym_linear_attn")
    _count(2)
    return out


def adaptive_avgpool(x, h, w, out=None):
    B, H, W, Cc = x.shape
    if out is None:
        out = new_act(B, h, w, Cc, x.device)
    _lib.check(lib().ym_adaptive_avgpool_nhwc(x.data_ptr(), pitch(x), B, H, W, Cc, h, w, out.data_ptr(), pitch(out), _stream()),
               "ym_adaptive_avgpool_nhwc")
    _count()
    return out


_LB_DTYPES = {torch.uint8: 0, torch.float16: 1, torch.float32: 2}


def letterbox(src, xtab, ytab, area2x, nw, nh, top, left, H, W, pad_value=114, swap_rb=True, chw=True, dtype=torch.uint8, out=None):
    """ym_letterbox_u8.  src: uint8 CUDA (B, sh, sw, 3) frames with dense rows; xtab / ytab: int32 CUDA (n, 2) tap tables
    (None when area2x).  Returns (B, 3, H, W) (chw) or (B, H, W, 3) of `dtype` (fp16 / fp32 are scaled by 1/255)."""
    if not _dev(src) or src.dtype != torch.uint8 or src.dim() != 4 or src.shape[3] != 3:
        raise ValueError(f"letterbox: expected uint8 CUDA frames (B, H, W, 3), got {tuple(src.shape)} {src.dtype} {src.device}")
    if src.stride(3) != 1 or src.stride(2) != 3:
        raise ValueError("letterbox: frames must have dense interleaved rows")
    B, sh, sw, _ = src.shape
    if dtype not in _LB_DTYPES:
        raise ValueError(f"letterbox: unsupported output dtype {dtype}")
    if not area2x:
        for t, n in ((xtab, nw), (ytab, nh)):
            if t is None or not _dev(t) or t.dtype != torch.int32 or tuple(t.shape) != (n, 2) or not t.is_contiguous():
                raise ValueError("letterbox: tap tables must be contiguous int32 CUDA tensors of shape (n, 2)")
    shape = (B, 3, H, W) if chw else (B, H, W, 3)
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=src.device)
    elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous() or not _dev(out):
        raise ValueError(f"letterbox: out must be a contiguous {dtype} CUDA tensor of shape {shape}")
    _lib.check(lib().ym_letterbox_u8(src.data_ptr(), src.stride(0) if B > 1 else 0, B, sh, sw, src.stride(1) if sh > 1 else 3 * sw,
                                     None if area2x else xtab.data_ptr(), None if area2x else ytab.data_ptr(), 1 if area2x else 0,
                                     nw, nh, top, left, int(pad_value), 1 if swap_rb else 0, out.data_ptr(), _LB_DTYPES[dtype],
                                     1 if chw else 0, H, W, _stream()), "ym_letterbox_u8")
    _count()
    return out


def scale_boxes(boxes, params, rows_per_img=0, row_img=None, padding=True, xywh=False):
    """ym_scale_boxes, in place.  boxes: fp32 CUDA (..., ld >= 4) rows with a dense last dimension; params: fp32 HOST (n_img, 5)
    = (gain, pad_x, pad_y, w0, h0); image of a row = row_img[row] (int32 CUDA) or row // rows_per_img."""
    if not _dev(boxes) or boxes.dtype != torch.float32 or boxes.dim() < 2 or boxes.shape[-1] < 4:
        raise ValueError("scale_boxes: expected an fp32 CUDA tensor (..., >= 4)")
    if boxes.dim() == 2 and boxes.stride(1) == 1 and (boxes.shape[0] == 1 or boxes.stride(0) >= boxes.shape[1]):
        ld = boxes.stride(0) if boxes.shape[0] > 1 else boxes.shape[1]   # e.g. the [:, :4] view of (n, 6) result rows
    elif boxes.is_contiguous():
        ld = boxes.shape[-1]
    else:
        raise ValueError("scale_boxes: rows must be dense with a constant pitch")
    if params.is_cuda or params.dtype != torch.float32 or params.dim() != 2 or params.shape[1] != 5 or not params.is_contiguous():
        raise ValueError("scale_boxes: params must be a contiguous fp32 HOST tensor (n_img, 5)")
    n = boxes.numel() // boxes.shape[-1]
    n_img = params.shape[0]
    if row_img is not None:
        if not _dev(row_img) or row_img.dtype != torch.int32 or row_img.numel() != n or not row_img.is_contiguous():
            raise ValueError("scale_boxes: row_img must be a contiguous int32 CUDA tensor with one entry per row")
    elif rows_per_img <= 0 or n > rows_per_img * n_img:
        raise ValueError("scale_boxes: rows_per_img * n_img must cover every row")
    if n_img > 128:
        raise ValueError("scale_boxes: at most 128 images per call")
    _lib.check(lib().ym_scale_boxes(boxes.data_ptr(), ld, n, rows_per_img, None if row_img is None else row_img.data_ptr(), n_img,
                                    params.data_ptr(), 1 if padding else 0, 1 if xywh else 0, _stream()), "ym_scale_boxes")
    _count()
    return boxes


def scale_coords(coords, params, padding=True, normalize=False):
    """ym_scale_coords, in place.  coords: fp32 CUDA contiguous (..., 2 | 3) points of ONE image; params: fp32 HOST (5,) row."""
    if not _dev(coords) or coords.dtype != torch.float32 or not coords.is_contiguous() or coords.shape[-1] < 2:
        raise ValueError("scale_coords: expected a contiguous fp32 CUDA tensor (..., >= 2)")
    if params.is_cuda or params.dtype != torch.float32 or params.numel() != 5 or not params.is_contiguous():
        raise ValueError("scale_coords: params must be a contiguous fp32 HOST tensor of 5 values")
    _lib.check(lib().ym_scale_coords(coords.data_ptr(), coords.shape[-1], coords.numel() // coords.shape[-1], params.data_ptr(),
                                     1 if padding else 0, 1 if normalize else 0, _stream()), "ym_scale_coords")
    _count()
    return coords


EW_SIGMOID, EW_MUL_GATE, EW_MUL = 6, 7, 8


def gate_router(x, pk, topk):
    """ym_gate_router.  x: (B,H,W,C) fp16 view (the dynamic channel half); pk: fp32 parameter pack of DualStreamGateRouter +
    complexity estimator (see VisualEnhancedAdaptiveGateMoE._build_pack).  Returns (idx int32 [B,k], w fp32 [B,k], probs [B,E])."""
    B, H, W, Cc = x.shape
    E, R = pk["E"], pk["R"]
    dev = x.device
    w = torch.empty((B, topk), dtype=torch.float32, device=dev)
    idx = torch.empty((B, topk), dtype=torch.int32, device=dev)
    probs = torch.empty((B, E), dtype=torch.float32, device=dev)
    scratch = torch.empty((lib().ym_gate_router_scratch_floats(B, H, W, Cc, R, E, pk["pool"]),), dtype=torch.float32, device=dev)
    ln, prior = pk.get("stat_norm"), pk.get("prior")          # DualStreamGateRouterV2: (weight, bias, eps) and the expert prior
    _lib.check(lib().ym_gate_router(x.data_ptr(), pitch(x), B, H, W, Cc, pk["pool"], pk["global_fc"].data_ptr(), pk["dw"].data_ptr(),
                                    pk["gn1_w"].data_ptr(), pk["gn1_b"].data_ptr(), pk["G1"], pk["pw1"].data_ptr(), R,
                                    pk["gn2_w"].data_ptr(), pk["gn2_b"].data_ptr(), pk["G2"], pk["pw2"].data_ptr(), pk["b2"].data_ptr(),
                                    E, pk["eps"], pk["alpha"], pk["temperature"], pk["cx_w"].data_ptr(), pk["cx_b"], topk,
                                    None if ln is None else ln[0].data_ptr(), None if ln is None else ln[1].data_ptr(),
                                    0.0 if ln is None else ln[2], None if prior is None else prior.data_ptr(),
                                    scratch.data_ptr(), w.data_ptr(), idx.data_ptr(), probs.data_ptr(), _stream()), "ym_gate_router")
    _count(6)
    return idx, w, probs


def zero_cost_router(x, fc, temperature, cx_w, cx_b, topk):
    """ym_zero_cost_router.  x: (B,H,W,C) fp16 view; fc fp32 [E,2C]; cx_w fp32 [C].  Returns (idx int32 [B,k], w fp32 [B,k], probs)."""
    B, H, W, Cc = x.shape
    E = fc.shape[0]
    w = torch.empty((B, topk), dtype=torch.float32, device=x.device)
    idx = torch.empty((B, topk), dtype=torch.int32, device=x.device)
    probs = torch.empty((B, E), dtype=torch.float32, device=x.device)
    scratch = torch.empty((lib().ym_zero_cost_router_scratch_floats(B, Cc),), dtype=torch.float32, device=x.device)
    _lib.check(lib().ym_zero_cost_router(x.data_ptr(), pitch(x), B, H, W, Cc, fc.data_ptr(), E, float(temperature), cx_w.data_ptr(),
                                         float(cx_b), topk, scratch.data_ptr(), w.data_ptr(), idx.data_ptr(), probs.data_ptr(), _stream()),
               "ym_zero_cost_router")
    _count(3)
    return idx, w, probs


def fc_gate(v, w1, w2, b2, scale=1.0, offset=0.0):
    """ym_fc_gate.  v: (B,1,1,Cin) fp16 pooled vector; w1 fp32 [Cr,Cin], w2 fp32 [Cout,Cr], b2 fp32 [Cout] or None -> fp32 (B, Cout)
    = offset + scale * sigmoid(w2 . silu(w1 . v) + b2)."""
    B, Cin = v.shape[0], v.shape[3]
    if v.shape[1] != 1 or v.shape[2] != 1:
        raise ValueError("fc_gate: expected a (B,1,1,C) pooled vector")
    out = torch.empty((B, w2.shape[0]), dtype=torch.float32, device=v.device)
    _lib.check(lib().ym_fc_gate(v.data_ptr(), pitch(v), B, Cin, w1.data_ptr(), w1.shape[0], w2.data_ptr(),
                                None if b2 is None else b2.data_ptr(), w2.shape[0], float(scale), float(offset), out.data_ptr(), _stream()),
               "ym_fc_gate")
    _count()
    return out


def gated_select(fo, idx, w, gamma, beta, E, oc, G, eps=1e-5, out=None):
    """ym_gated_select.  fo: (B,H,W,E*oc) fp16 all-expert conv output; idx int32 / w fp32 [B,k]; gamma, beta fp32 [E,oc]."""
    B, H, W, _ = fo.shape
    k = idx.shape[1]
    if out is None:
        out = new_act(B, H, W, oc, fo.device)
    scratch = torch.empty((lib().ym_gated_select_scratch_floats(B, k, oc),), dtype=torch.float32, device=fo.device)
    _lib.check(lib().ym_gated_select(fo.data_ptr(), pitch(fo), B, H * W, E, oc, G, eps, idx.data_ptr(), w.data_ptr(), k,
                                     gamma.data_ptr(), beta.data_ptr(), scratch.data_ptr(), out.data_ptr(), pitch(out), _stream()),
               "ym_gated_select")
    _count(3)
    return out


def ctx_mean3(a, b, c, out=None):
    """ym_ctx_mean3: (a + nearest_up(b) + nearest_up(c)) / 3 on (B,H,W,C) fp16 views."""
    B, H, W, Cc = a.shape
    if out is None:
        out = new_act(B, H, W, Cc, a.device)
    _lib.check(lib().ym_ctx_mean3(a.data_ptr(), pitch(a), b.data_ptr(), pitch(b), c.data_ptr(), pitch(c), B, H, W, Cc,
                                  b.shape[1], b.shape[2], c.shape[1], c.shape[2], out.data_ptr(), pitch(out), _stream()), "ym_ctx_mean3")
    _count()
    return out


def kpts_decode(kpts, strides, ndim):
    """ym_kpts_decode.  kpts[l]: fp32 (B,h,w,nk) pose-tower outputs -> fp32 (B, nk, A)."""
    B, nk = kpts[0].shape[0], kpts[0].shape[3]
    A = sum(k.shape[1] * k.shape[2] for k in kpts)
    y = torch.empty((B, nk, A), dtype=torch.float32, device=kpts[0].device)
    nl, kp, _, hs, ws, st = _level_arrays(kpts, kpts, strides)
    _lib.check(lib().ym_kpts_decode(nl, kp, hs, ws, st, B, nk, ndim, y.data_ptr(), _stream()), "ym_kpts_decode")
    _count()
    return y


def obb_finish(y, angles, strides, nc):
    """ym_obb_finish.  y: fp32 (B, 4+nc, A) xywh dense decode; angles[l]: fp32 (B,h,w,1) raw angle-tower outputs -> fp32 (B, 4+nc+1, A)."""
    B, rows, A = y.shape
    out = torch.empty((B, rows + 1, A), dtype=torch.float32, device=y.device)
    nl, ap, _, hs, ws, st = _level_arrays(angles, angles, strides)
    _lib.check(lib().ym_obb_finish(nl, ap, hs, ws, st, B, nc, y.contiguous().data_ptr(), out.data_ptr(), _stream()), "ym_obb_finish")
    _count()
    return out


def classify_head(v, w, b):
    """ym_classify_head.  v: (B,1,1,Cin) fp16 pooled features; w fp32 [nc,Cin], b fp32 [nc] -> (probs, logits) fp32 (B, nc)."""
    B, Cin, nc = v.shape[0], v.shape[3], w.shape[0]
    logits = torch.empty((B, nc), dtype=torch.float32, device=v.device)
    probs = torch.empty((B, nc), dtype=torch.float32, device=v.device)
    _lib.check(lib().ym_classify_head(v.data_ptr(), pitch(v), B, Cin, w.data_ptr(), None if b is None else b.data_ptr(), nc,
                                      logits.data_ptr(), probs.data_ptr(), _stream()), "ym_classify_head")
    _count()
    return probs, logits


def pixel_router(x, pk, topk, w_min=0.01):
    """ym_pixel_router (UltraEfficientRouter).  x: (B,H,W,C) fp16 view; pk: fp32 pack of the router's local stream.
    Returns (idx int32 [B,k], w fp32 [B,k] with weights <= w_min zeroed, probs fp32 [B,E])."""
    B, H, W, Cc = x.shape
    E, R = pk["E"], pk["R"]
    w = torch.empty((B, topk), dtype=torch.float32, device=x.device)
    idx = torch.empty((B, topk), dtype=torch.int32, device=x.device)
    probs = torch.empty((B, E), dtype=torch.float32, device=x.device)
    scratch = torch.empty((lib().ym_gate_router_scratch_floats(B, H, W, Cc, R, E, pk["pool"]),), dtype=torch.float32, device=x.device)
    _lib.check(lib().ym_pixel_router(x.data_ptr(), pitch(x), B, H, W, Cc, pk["pool"], pk["dw"].data_ptr(), pk["gn1_w"].data_ptr(),
                                     pk["gn1_b"].data_ptr(), pk["G1"], pk["pw1"].data_ptr(), R, pk["gn2_w"].data_ptr(), pk["gn2_b"].data_ptr(),
                                     pk["G2"], pk["pw2"].data_ptr(), pk["b2"].data_ptr(), E, pk["eps"], pk["temperature"], float(w_min), topk,
                                     scratch.data_ptr(), w.data_ptr(), idx.data_ptr(), probs.data_ptr(), _stream()), "ym_pixel_router")
    _count(5)
    return idx, w, probs


def latent_router(tokens, pk):
    """ym_latent_router.  tokens: list of (B,1,1,C) fp16 pooled vectors; pk: fp32 pack of LatentRouter.  Returns (probs, logits) (B, E)."""
    T, B, Cc = len(tokens), tokens[0].shape[0], tokens[0].shape[3]
    E = pk["wh"].shape[0]
    logits = torch.empty((B, E), dtype=torch.float32, device=tokens[0].device)
    probs = torch.empty((B, E), dtype=torch.float32, device=tokens[0].device)
    tp = (C.c_void_p * T)(*[t.data_ptr() for t in tokens])
    lds = (C.c_int * T)(*[pitch(t) for t in tokens])
    emb = pk.get("emb")
    _lib.check(lib().ym_latent_router(T, tp, lds, B, Cc, None if emb is None else emb.data_ptr(), pk["ln_w"].data_ptr(), pk["ln_b"].data_ptr(),
                                      pk["ln_eps"], pk["w1"].data_ptr(), pk["b1"].data_ptr(), pk["w1"].shape[0], pk["w2"].data_ptr(),
                                      pk["b2"].data_ptr(), pk["wh"].data_ptr(), pk["bh"].data_ptr(), E, pk["temperature"],
                                      logits.data_ptr(), probs.data_ptr(), _stream()), "ym_latent_router")
    _count()
    return probs, logits


def gap(x, out=None):
    """ym_gap_nhwc: global average pool of a (B,H,W,C) fp16 view -> (B,1,1,C) fp16."""
    B, H, W, Cc = x.shape
    if out is None:
        out = new_act(B, 1, 1, Cc, x.device)
    _lib.check(lib().ym_gap_nhwc(x.data_ptr(), pitch(x), B, H * W, Cc, out.data_ptr(), pitch(out), _stream()), "ym_gap_nhwc")
    _count()
    return out
