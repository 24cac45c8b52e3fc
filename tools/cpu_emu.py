"""DEV TOOL (never imported by the product): torch-CPU emulation of the C-ABI semantics of the MoT / MoA ops, monkey-patched
over `yolo_master_b200.ops` so that the HOST wiring of the modules (weight packing, slicing, folding, op order) can be checked
against the oracle in the GPU-less build container.  The kernels themselves are verified on the GPU by tests/test_gpu_mot.py.

    python tools/cpu_emu.py
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import yolo_master_oracle as O  # noqa: E402
from yolo_master_b200 import ops  # noqa: E402
from yolo_master_b200.nn.modules import _base, block, conv, moa, mot  # noqa: E402
from yolo_master_b200.utils.synth import fill_state_dict_  # noqa: E402


def _out(y, out, dtype=torch.float16):
    y = y.to(dtype)
    if out is None:
        return y.contiguous()
    out.copy_(y)
    return out


def conv2d(x, w_packed, bias, Cout, KH, KW, stride, pad, act, out=None, res=None, out_f32=False):
    B, H, W, Cin = x.shape
    w = w_packed.float()[:, :KH * KW * Cin].reshape(Cout, KH, KW, Cin).permute(0, 3, 1, 2)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, None if bias is None else bias.float(), stride, pad)
    if act:
        y = F.silu(y)
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + res.float()
    return _out(y, out, torch.float32 if out_f32 else torch.float16)


def dwconv(x, w_taps, bias, ksize, act, C_out, grp_w=None, grp_stride=None, grp_off=0, add=None, out=None):
    xs = x.float()
    if grp_w is not None and not (grp_w == C_out and grp_off == 0):   # source channel of c = (c/grp_w)*grp_stride + grp_off + c%grp_w
        src = torch.tensor([(c // grp_w) * grp_stride + grp_off + c % grp_w for c in range(C_out)])
        xs = xs[..., src]
    else:
        xs = xs[..., :C_out]
    w = w_taps.float().t().reshape(C_out, 1, ksize, ksize)
    y = F.conv2d(xs.permute(0, 3, 1, 2), w, None if bias is None else bias.float(), 1, ksize // 2, 1, C_out)
    if act:
        y = F.silu(y)
    y = y.permute(0, 2, 3, 1)
    if add is not None:
        y = y + add.float()
    return _out(y, out)


def ew(op, a=None, b=None, p0=None, p1=None, ldt=0, toff=0, rows_per_img=1, act=False, out=None, tok=None):
    ref = a if a is not None else b
    B, H, W, C = ref.shape
    fa = torch.zeros(ref.shape) if a is None else a.float()
    fb = torch.zeros(ref.shape) if b is None else b.float()
    tw = None if tok is None else tok.view(B, H, W, ldt)[..., toff:toff + 1]
    if op == ops.EW_SCALE_RES:
        y = fa + p0.view(1, 1, 1, C) * fb
    elif op == ops.EW_TOKEN_ACC:
        y = fa + tw * fb
    elif op == ops.EW_GLU:
        y = torch.sigmoid(fa) * fb
    elif op == ops.EW_GELU:
        y = F.gelu(fa)
    elif op == ops.EW_LERP:
        y = p0[0] * fa + (1 - p0[0]) * fb
    elif op == ops.EW_SIGMOID:
        y = torch.sigmoid(fa)
    elif op == ops.EW_MUL_GATE:
        y = fa * (1 + p0[0] * fb)
    elif op == ops.EW_MUL:
        y = fa * fb
    else:
        assert rows_per_img == H * W
        v = fa * p0.view(B, 1, 1, C) + p1.view(B, 1, 1, C)
        if act:
            v = F.silu(v)
        y = (v if tw is None else tw * v) + fb
    return _out(y, out)


def groupnorm_stats(x, G, gamma, beta, eps=1e-5):
    B, H, W, C = x.shape
    xf = x.float().reshape(B, H * W, G, C // G)
    mean = xf.mean((1, 3))
    var = xf.var((1, 3), unbiased=False)
    rstd = (var + eps).rsqrt()
    sc = rstd.repeat_interleave(C // G, 1) * gamma
    sh = beta - mean.repeat_interleave(C // G, 1) * sc
    return sc.contiguous(), sh.contiguous()


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    return _out(F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps), out)


def _h(t, nh, hdp):
    B, H, W, _ = t.shape
    return t.float().reshape(B, H * W, nh, hdp).permute(0, 2, 1, 3)


def attn_small(q, k, v, heads, hdp, scale, out=None):
    B, H, W, _ = q.shape
    o = O._sdpa(_h(q, heads, hdp), _h(k, heads, hdp), _h(v, heads, hdp), scale)
    return _out(o.permute(0, 2, 1, 3).reshape(B, H, W, heads * hdp), out)


def attn_window(q, k, v, heads, hdp, win, shift, scale, padk=None, padv=None, out=None):
    B, H, W, C = q.shape
    Hp, Wp = math.ceil(H / win) * win, math.ceil(W / win) * win

    def prep(t, pad):
        full = (torch.zeros(C) if pad is None else pad.float()).view(1, 1, 1, C).expand(B, Hp, Wp, C).clone()
        full[:, :H, :W] = t.float()
        if shift:
            full = torch.roll(full, (-shift, -shift), (1, 2))
        return O._win_part(full, win).reshape(-1, win * win, heads, hdp).permute(0, 2, 1, 3)

    o = O._sdpa(prep(q, None), prep(k, padk), prep(v, padv), scale).transpose(1, 2).reshape(-1, win * win, C)
    o = O._win_rev(o, win, Hp, Wp)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    return _out(o[:, :H, :W], out)


def deform_sample(oa, v, heads, hd, n_points, align_corners=True, out=None):
    B, H, W, C = v.shape
    N = H * W
    off = oa[..., :heads * n_points * 2].reshape(B, N, heads, n_points, 2).tanh()
    aw = oa[..., heads * n_points * 2:].reshape(B, N, heads, n_points).softmax(-1)
    idx = torch.arange(N)
    ref = torch.stack([(idx % W).float() / max(W - 1, 1) * 2 - 1, (idx // W).float() / max(H - 1, 1) * 2 - 1], -1)[None, :, None, None, :]
    locs = (ref + off * 0.25).clamp(-1, 1)
    v4 = v.float().reshape(B, N, heads, hd).permute(0, 2, 3, 1).reshape(B * heads, hd, H, W)
    samp = F.grid_sample(v4, locs.permute(0, 2, 1, 3, 4).reshape(B * heads, N, n_points, 2), mode="bilinear", padding_mode="zeros",
                         align_corners=align_corners)
    o = (aw.unsqueeze(-1) * samp.reshape(B, heads, hd, N, n_points).permute(0, 3, 1, 4, 2)).sum(3)
    return _out(o.reshape(B, H, W, C), out)


def token_router(x, pk, topk, temp_dev=None, temp=1.0, want_idx=True):
    B, H, W, C = x.shape
    h = x.float().reshape(B, H * W, C) @ pk["w1"].t()
    hid = h.shape[-1]
    hn = F.group_norm(h.transpose(1, 2), pk["G"], pk["gn_w"], pk["gn_b"], 1e-5).transpose(1, 2)
    lg = F.silu(hn) @ pk["w2"].t() + pk["b2"]
    T = float(temp_dev[0]) if temp_dev is not None else temp
    p = F.softmax(lg / T, -1)
    E = p.shape[-1]
    if topk < E:
        vals, idx = p.topk(topk, -1)
        vals = vals / vals.sum(-1, keepdim=True).clamp_min(1e-6)
        p = torch.zeros_like(p).scatter_(-1, idx, vals)
    else:
        idx = torch.arange(E).expand(B, H * W, E)
    return p.reshape(B, H, W, E).contiguous(), (idx.reshape(B, H, W, -1).int() if want_idx else None)


def linear_attn(q, k, v, heads, hdp, hd, rf, eps=1e-6, limit=1e4, out=None):
    B, H, W, _ = q.shape
    o = O._linear_attn(_h(q, heads, hdp)[..., :hd], _h(k, heads, hdp)[..., :hd], _h(v, heads, hdp)[..., :hd], rf, limit, eps)
    full = torch.zeros(B, heads, H * W, hdp)
    full[..., :hd] = o
    return _out(full.permute(0, 2, 1, 3).reshape(B, H, W, heads * hdp), out)


def gap(x, out=None):
    return _out(x.float().mean((1, 2), keepdim=True), out)


def adaptive_avgpool(x, h, w, out=None):
    return _out(F.adaptive_avg_pool2d(x.float().permute(0, 3, 1, 2), (h, w)).permute(0, 2, 3, 1), out)


def moe_expert_gemm(a, lda, a_div, P, HW, K, w_all, route_idx, N, a_scale=None, a_shift=None, groups=0):
    """out[p] = A[p // a_div] @ W[route_idx[p]]^T (optionally A := SiLU(A*scale + shift)); the "stats" handle is the stored output."""
    A = a.reshape(-1, HW, a.shape[-1])[..., :K].float()
    outs = []
    for p in range(P):
        Ap = A[p // a_div]
        if a_scale is not None:
            Ap = F.silu(Ap * a_scale[p].view(1, K) + a_shift[p].view(1, K))
        outs.append(Ap @ w_all[int(route_idx[p])].float()[:N, :K].t())
    out = torch.stack(outs).half()
    return out, (out if groups else None)


def moe_ffn_supported(C, HID, ldx):
    return (C, HID) in ((64, 128), (128, 256)) and ldx % 8 == 0


def moe_ffn_stats(x, topk, w1, route_idx):
    """Stage 1 of ym_moe_ffn: the "statistics" handle of the emulation is the fp16-rounded hidden activation itself."""
    B, H, W, Cc = x.shape
    h, _ = moe_expert_gemm(x, x.shape[-1], topk, B * topk, H * W, Cc, w1, route_idx, w1.shape[1], groups=1)
    return h, 1


def moe_ffn_fused(x, topk, w1, w2, route_idx, a_scale, a_shift, strips):
    B, H, W, Cc = x.shape
    P, HW, HID = B * topk, H * W, w1.shape[1]
    h, _ = moe_expert_gemm(x, x.shape[-1], topk, P, HW, Cc, w1, route_idx, HID, groups=1)
    o, _ = moe_expert_gemm(h, HID, 1, P, HW, HID, w2, route_idx, Cc, a_scale=a_scale, a_shift=a_shift, groups=1)
    return o, o


def gn_finalize_tiles(stats, P, tiles, G, C_, count, eps, gamma, beta, route_idx, route_w=None):
    return gn_finalize(stats, P, stats.shape[1], G, C_, count, eps, gamma, beta, route_idx, route_w)


def router_partial(x, pack, pool=4):
    return (x, pack, pool), 0, 0                  # the emulation finishes the routing in moe_ffn_stats_routed


def moe_ffn_stats_routed(x, topk, w1, rpack, partial, nblk, npix):
    idx, w, probs = router_topk(x, rpack, topk, partial[2])
    st1, strips = moe_ffn_stats(x, topk, w1, idx.view(-1))
    return idx, w, probs, st1, strips


def moe_ffn_fused_gn(x, topk, w1, w2, route_idx, st1, strips, G1, count1, eps1, gamma1, beta1):
    P, HID = x.shape[0] * topk, w1.shape[1]
    sc1, sh1 = gn_finalize_tiles(st1, P, strips, G1, HID, count1, eps1, gamma1, beta1, route_idx)
    return moe_ffn_fused(x, topk, w1, w2, route_idx, sc1, sh1, strips)


def moe_combine_gn_supported(x, ws_packed, o, topk, out=None):
    return True


def moe_combine_gn(x, ws_packed, bias_s, o, st2, strips, G2, count2, eps2, gamma2, beta2, route_idx, route_w, topk, add_residual=True, out=None):
    P, Cc = x.shape[0] * topk, x.shape[3]
    sc2, sh2 = gn_finalize_tiles(st2, P, strips, G2, Cc, count2, eps2, gamma2, beta2, route_idx, route_w)
    return moe_combine(x, ws_packed, bias_s, o, sc2, sh2, topk, add_residual=add_residual, out=out)


def gn_finalize(stats, P, HW, G, C_, count, eps, gamma, beta, route_idx, route_w=None):
    o = stats.float().reshape(P, HW, G, C_ // G)
    mean = o.mean((1, 3))
    rstd = 1.0 / torch.sqrt(o.var((1, 3), unbiased=False) + eps)
    e = route_idx.long()
    rw = torch.ones(P) if route_w is None else route_w.float()
    g, b = gamma[e].float(), beta[e].float()
    mean, rstd = mean.repeat_interleave(C_ // G, 1), rstd.repeat_interleave(C_ // G, 1)
    return (rw[:, None] * rstd * g).contiguous(), (rw[:, None] * (b - mean * rstd * g)).contiguous()


def install_gated(host):
    """Gated-family ops backed by the HOST build of their kernel bodies (tests/native/gated_host.cpp): `host` is its ctypes handle."""
    import ctypes as C
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    host.host_gate_router.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp, vp, vp, vp, ci, vp, ci, vp, vp, ci, vp, vp, ci, cf, cf, cf, vp, cf,
                                      ci, vp, vp, cf, vp, vp, vp, vp]
    host.host_pixel_router.argtypes = [vp, ci, ci, ci, ci, ci, ci, vp, vp, vp, ci, vp, ci, vp, vp, ci, vp, vp, ci, cf, cf, cf, ci, vp, vp, vp]
    host.host_zero_cost_router.argtypes = [vp, ci, ci, ci, ci, ci, vp, ci, cf, vp, cf, ci, vp, vp, vp]
    host.host_classify_head.argtypes = [vp, ci, ci, ci, vp, vp, ci, vp, vp]
    host.host_fc_gate.argtypes = [vp, ci, ci, ci, vp, ci, vp, vp, ci, cf, cf, vp]
    host.host_gated_select.argtypes = [vp, ci, ci, ci, ci, ci, ci, cf, vp, vp, ci, vp, vp, vp, ci]
    host.host_ctx_mean3.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, ci]

    def ld(t):   # row pitch of a (B,H,W,C) view with a dense channel dimension (what ops.pitch validates on the GPU path)
        assert t.dtype == torch.float16 and t.stride(3) == 1
        return t.stride(2) if t.shape[2] > 1 else (t.stride(1) if t.shape[1] > 1 else (t.stride(0) if t.shape[0] > 1 else t.shape[3]))

    def gate_router(x, pk, topk):
        B, H, W, Cc = x.shape
        w = torch.empty((B, topk), dtype=torch.float32)
        idx = torch.empty((B, topk), dtype=torch.int32)
        probs = torch.empty((B, pk["E"]), dtype=torch.float32)
        ln, prior = pk.get("stat_norm"), pk.get("prior")
        host.host_gate_router(x.data_ptr(), ld(x), B, H, W, Cc, pk["pool"], pk["global_fc"].data_ptr(), pk["dw"].data_ptr(),
                              pk["gn1_w"].data_ptr(), pk["gn1_b"].data_ptr(), pk["G1"], pk["pw1"].data_ptr(), pk["R"],
                              pk["gn2_w"].data_ptr(), pk["gn2_b"].data_ptr(), pk["G2"], pk["pw2"].data_ptr(), pk["b2"].data_ptr(), pk["E"],
                              pk["eps"], pk["alpha"], pk["temperature"], pk["cx_w"].data_ptr(), pk["cx_b"], topk,
                              None if ln is None else ln[0].data_ptr(), None if ln is None else ln[1].data_ptr(),
                              0.0 if ln is None else ln[2], None if prior is None else prior.data_ptr(), w.data_ptr(),
                              idx.data_ptr(), probs.data_ptr())
        return idx, w, probs

    def pixel_router(x, pk, topk, w_min=0.01):
        B, H, W, Cc = x.shape
        w = torch.empty((B, topk), dtype=torch.float32)
        idx = torch.empty((B, topk), dtype=torch.int32)
        probs = torch.empty((B, pk["E"]), dtype=torch.float32)
        host.host_pixel_router(x.data_ptr(), ld(x), B, H, W, Cc, pk["pool"], pk["dw"].data_ptr(), pk["gn1_w"].data_ptr(), pk["gn1_b"].data_ptr(),
                               pk["G1"], pk["pw1"].data_ptr(), pk["R"], pk["gn2_w"].data_ptr(), pk["gn2_b"].data_ptr(), pk["G2"],
                               pk["pw2"].data_ptr(), pk["b2"].data_ptr(), pk["E"], pk["eps"], pk["temperature"], float(w_min), topk,
                               w.data_ptr(), idx.data_ptr(), probs.data_ptr())
        return idx, w, probs

    def zero_cost_router(x, fc, temperature, cx_w, cx_b, topk):
        B, H, W, Cc = x.shape
        w = torch.empty((B, topk), dtype=torch.float32)
        idx = torch.empty((B, topk), dtype=torch.int32)
        probs = torch.empty((B, fc.shape[0]), dtype=torch.float32)
        host.host_zero_cost_router(x.data_ptr(), ld(x), B, H, W, Cc, fc.data_ptr(), fc.shape[0], float(temperature), cx_w.data_ptr(),
                                   float(cx_b), topk, w.data_ptr(), idx.data_ptr(), probs.data_ptr())
        return idx, w, probs

    def fc_gate(v, w1, w2, b2, scale=1.0, offset=0.0):
        B, Cin = v.shape[0], v.shape[3]
        out = torch.empty((B, w2.shape[0]), dtype=torch.float32)
        host.host_fc_gate(v.data_ptr(), ld(v), B, Cin, w1.data_ptr(), w1.shape[0], w2.data_ptr(), None if b2 is None else b2.data_ptr(),
                          w2.shape[0], float(scale), float(offset), out.data_ptr())
        return out

    def classify_head(v, w, b):
        B, Cin, nc = v.shape[0], v.shape[3], w.shape[0]
        logits, probs = torch.empty((B, nc), dtype=torch.float32), torch.empty((B, nc), dtype=torch.float32)
        host.host_classify_head(v.data_ptr(), ld(v), B, Cin, w.data_ptr(), None if b is None else b.data_ptr(), nc, logits.data_ptr(),
                                probs.data_ptr())
        return probs, logits

    def gated_select(fo, idx, w, gamma, beta, E, oc, G, eps=1e-5, out=None):
        B, H, W, _ = fo.shape
        if out is None:
            out = torch.empty((B, H, W, oc), dtype=torch.float16)
        host.host_gated_select(fo.data_ptr(), ld(fo), B, H * W, E, oc, G, eps, idx.data_ptr(), w.data_ptr(), idx.shape[1], gamma.data_ptr(),
                               beta.data_ptr(), out.data_ptr(), ld(out))
        return out

    def ctx_mean3(a, b, c, out=None):
        B, H, W, Cc = a.shape
        if out is None:
            out = torch.empty((B, H, W, Cc), dtype=torch.float16)
        host.host_ctx_mean3(a.data_ptr(), ld(a), b.data_ptr(), ld(b), c.data_ptr(), ld(c), B, H, W, Cc, b.shape[1], b.shape[2],
                            c.shape[1], c.shape[2], out.data_ptr(), ld(out))
        return out

    for name, fn in dict(gate_router=gate_router, pixel_router=pixel_router, zero_cost_router=zero_cost_router, fc_gate=fc_gate, classify_head=classify_head, gated_select=gated_select, ctx_mean3=ctx_mean3,
                         moe_expert_gemm=moe_expert_gemm, gn_finalize=gn_finalize).items():
        setattr(ops, name, fn)
    ops.pitch = lambda t, dtype=torch.float16: ld(t)
    from yolo_master_b200.nn.modules import gated
    gated.to_nhwc = _base.to_nhwc


# ---- the remaining ops of a whole model (stem, area attention, upsample+concat, SPPF pooling, image router, MoE combine, Detect):
# with these, DetectionModel forwards run end to end on the CPU emulation (tests/test_host_model_wiring.py)
def stem_conv(img, wgt, bias, Cout, out=None):
    B, Cin, H, W = img.shape
    x = img.float() / 255 if img.dtype == torch.uint8 else img.float()
    w = wgt.float().reshape(Cin, 3, 3, Cout).permute(3, 0, 1, 2)                  # [Cin*9][Cout], k = (ci*3+ky)*3+kx
    y = F.silu(F.conv2d(x, w, bias.float(), 2, 1)).permute(0, 2, 3, 1)
    return _out(y, out)


def attention(qkv, batch, N, heads, head_stride, q_off, k_off, v_off, d_qk, d_v, scale, out=None):
    B, H, W, Ct = qkv.shape
    rows = qkv.reshape(batch, N, Ct).float()
    outs = []
    for h in range(heads):
        b0 = h * head_stride
        q, k, v = rows[..., b0 + q_off:b0 + q_off + d_qk], rows[..., b0 + k_off:b0 + k_off + d_qk], rows[..., b0 + v_off:b0 + v_off + d_v]
        outs.append(torch.softmax((q * scale) @ k.transpose(1, 2), -1) @ v)
    y = torch.cat(outs, -1).reshape(B, H, W, heads * d_v)
    return _out(y, out)


def concat2(a, b, up=1, out=None):
    ya = a.float()
    if up > 1:
        ya = ya.repeat_interleave(up, 1).repeat_interleave(up, 2)
    y = ya if b is None else torch.cat([ya, b.float()], 3)
    return _out(y, out)


def sppf_pool(buf, Cslot, k):
    t = buf[..., :Cslot].float().permute(0, 3, 1, 2)
    for i in range(1, 4):
        t = F.max_pool2d(t, k, 1, k // 2)
        buf[..., i * Cslot:(i + 1) * Cslot] = t.permute(0, 2, 3, 1).half()
    return buf


def router_topk(x, pack, topk, pool=4):
    B, H, W, C = x.shape
    Cr, E = pack["Cr"], pack["E"]
    xin = x.float().permute(0, 3, 1, 2)
    if H > pool and W > pool:
        xin = F.avg_pool2d(xin, pool, pool)
    w1 = pack["w1"].permute(0, 1, 3, 2).reshape(3, 3, C, Cr).permute(3, 2, 0, 1)   # [tap][c/4][r][4] -> [r][c][ky][kx]
    h = F.silu(F.conv2d(xin, w1, None, 1, 1) * pack["scale1"].view(1, -1, 1, 1) + pack["shift1"].view(1, -1, 1, 1))
    o = F.conv2d(h, pack["w2"].view(E, Cr, 1, 1)) * pack["scale2"].view(1, -1, 1, 1) + pack["shift2"].view(1, -1, 1, 1)
    probs = torch.softmax(o.mean((2, 3)), 1)
    w, idx = torch.topk(probs, topk, 1)
    return idx.int(), (w / w.sum(1, keepdim=True).clamp_min(1e-6)).contiguous(), probs


def moe_combine(x, ws_packed, bias_s, o, o_scale, o_shift, topk, add_residual=True, out=None):
    B, H, W, C = x.shape
    shared = F.silu(x.float() @ ws_packed.float()[:, :C].t() + bias_s.view(1, 1, 1, -1))
    e = (o.float() * o_scale.view(B * topk, 1, -1) + o_shift.view(B * topk, 1, -1)).reshape(B, topk, H, W, C).sum(1)
    y = shared + e + (x.float() if add_residual else 0)
    return _out(y, out)


def detect_dense(boxes, logits, strides, nc, xyxy, reg_max=1):
    lv = [b.shape[1:3] for b in boxes]
    bx = torch.cat([b.float().reshape(b.shape[0], -1, b.shape[3]).transpose(1, 2) for b in boxes], 2)
    sc = torch.cat([c.float().reshape(c.shape[0], -1, c.shape[3]).transpose(1, 2) for c in logits], 2)
    return O.detect_decode(bx, sc, lv, [float(s) for s in strides], xyxy, reg_max)


def detect_topk(boxes, logits, strides, nc, max_det=300, return_anchor=False):
    r, idx = O.detect_postprocess(detect_dense(boxes, logits, strides, nc, True, 1), nc, max_det)
    return (r, idx.int()) if return_anchor else r


def kpts_decode(kpts, strides, ndim):
    B, nk = kpts[0].shape[0], kpts[0].shape[3]
    outs = []
    for t, s in zip(kpts, strides):
        h, w = t.shape[1:3]
        y = t.float().reshape(B, h * w, nk).transpose(1, 2).clone()                 # (B, nk, h*w)
        if ndim > 1:                                                                # ndim 1: plain gather (mask coefficients)
            gx = torch.arange(w, dtype=torch.float32).repeat(h)
            gy = torch.arange(h, dtype=torch.float32).repeat_interleave(w)
            y[:, 0::ndim] = (y[:, 0::ndim] * 2.0 + gx) * float(s)
            y[:, 1::ndim] = (y[:, 1::ndim] * 2.0 + gy) * float(s)
        if ndim == 3:
            y[:, 2::ndim] = torch.sigmoid(y[:, 2::ndim])
        outs.append(y)
    return torch.cat(outs, 2)


def obb_finish(y, angles, strides, nc):
    B, rows, A = y.shape
    raw = torch.cat([t.float().reshape(B, -1, 1).transpose(1, 2) for t in angles], 2)      # (B, 1, A)
    ang = (torch.sigmoid(raw) - 0.25) * math.pi
    ax = torch.cat([(torch.arange(t.shape[2], dtype=torch.float32) + 0.5).repeat(t.shape[1]) for t in angles])
    ay = torch.cat([(torch.arange(t.shape[1], dtype=torch.float32) + 0.5).repeat_interleave(t.shape[2]) for t in angles])
    st = torch.cat([torch.full((t.shape[1] * t.shape[2],), float(s)) for t, s in zip(angles, strides)])
    xf, yf = y[:, 0] / st - ax, y[:, 1] / st - ay
    c, s_ = torch.cos(ang[:, 0]), torch.sin(ang[:, 0])
    out = torch.cat([y, ang], 1).clone()
    out[:, 0] = (xf * c - yf * s_ + ax) * st
    out[:, 1] = (xf * s_ + yf * c + ay) * st
    return out


def latent_router(tokens, pk):
    x = torch.stack([t.float().reshape(t.shape[0], -1) for t in tokens], 1)
    if pk.get("emb") is not None:
        x = x + pk["emb"].unsqueeze(0)
    r = F.layer_norm(x.mean(1), (x.shape[-1],), pk["ln_w"], pk["ln_b"], pk["ln_eps"])
    h = F.silu(F.linear(F.silu(F.linear(r, pk["w1"], pk["b1"])), pk["w2"], pk["b2"]))
    logits = torch.nan_to_num(F.linear(h, pk["wh"], pk["bh"]), nan=0.0, posinf=30.0, neginf=-30.0).clamp(-30.0, 30.0)
    return torch.softmax(logits / max(pk["temperature"], 0.1), -1), logits


def install_model():
    """Everything `install()` covers plus the whole-model ops above."""
    install()
    for name, fn in dict(stem_conv=stem_conv, attention=attention, concat2=concat2, sppf_pool=sppf_pool, router_topk=router_topk,
                         moe_combine=moe_combine, moe_expert_gemm=moe_expert_gemm, gn_finalize=gn_finalize, detect_dense=detect_dense,
                         moe_ffn_supported=moe_ffn_supported, moe_ffn_stats=moe_ffn_stats, moe_ffn_fused=moe_ffn_fused, gn_finalize_tiles=gn_finalize_tiles,
                         moe_ffn_fused_gn=moe_ffn_fused_gn, router_partial=router_partial, moe_ffn_stats_routed=moe_ffn_stats_routed, moe_combine_gn_supported=moe_combine_gn_supported, moe_combine_gn=moe_combine_gn,
                         detect_topk=detect_topk, kpts_decode=kpts_decode, obb_finish=obb_finish, latent_router=latent_router).items():
        setattr(ops, name, fn)
    ops.pitch = lambda t, dtype=torch.float16: (t.stride(2) if t.shape[2] > 1 else (t.stride(1) if t.shape[1] > 1 else (t.stride(0) if t.shape[0] > 1 else t.shape[3])))
    from yolo_master_b200.nn.modules import gated, head, latent, moe
    for mod in (gated, head, latent, moe):
        if hasattr(mod, "to_nhwc"):
            mod.to_nhwc = _base.to_nhwc

    def conv_forward(self, x):                   # Conv.forward without its "CUDA tensors only" guard on the stem
        pk = self.get_pack()
        if pk["kind"] == "stem":
            return _base.to_nchw(ops.stem_conv(x, pk["w"], pk["bias"], self.conv.out_channels))
        return _base.to_nchw(self.fwd_nhwc(conv.to_nhwc(x)))
    conv.Conv.forward = conv.Conv.forward_fuse = conv_forward


def dwconv3_routed(x, w_taps, route, dil):
    """Torch restatement of ym_dwconv3_routed_nhwc (checked against the kernel itself in tests/test_cuda_host_emu.py)."""
    B, H, W, C = x.shape
    y = torch.empty((B, C, H, W))
    for b in range(B):
        e = int(route[b])
        d = int(dil[e])
        y[b] = F.conv2d(x[b:b + 1].float().permute(0, 3, 1, 2), w_taps[e].float().t().reshape(C, 1, 3, 3), None, 1, d, d, C)[0]
    return _out(y.permute(0, 2, 3, 1), None)


def route_affine(scale, shift, gamma, beta, route, route_w=None):
    g, bt = gamma[route.long()], beta[route.long()]
    scale.mul_(g)
    shift.mul_(g).add_(bt)
    if route_w is not None:
        scale.mul_(route_w.view(-1, 1))
        shift.mul_(route_w.view(-1, 1))
    return scale, shift


def install():
    for name, fn in dict(dwconv3_routed=dwconv3_routed, route_affine=route_affine).items():
        setattr(ops, name, fn)
    for name, fn in dict(conv2d=conv2d, dwconv=dwconv, ew=ew, groupnorm_stats=groupnorm_stats, layernorm=layernorm, attn_small=attn_small,
                         attn_window=attn_window, deform_sample=deform_sample, token_router=token_router, linear_attn=linear_attn,
                         adaptive_avgpool=adaptive_avgpool, gap=gap).items():
        setattr(ops, name, fn)
    ops.new_act = lambda B, H, W, C, device: torch.empty((B, H, W, C), dtype=torch.float16)

    def to_nhwc(x):
        v = x.half().permute(0, 2, 3, 1)
        return v if v.is_contiguous() else v.contiguous()

    for mod in (_base, block, conv, moa, mot):
        if hasattr(mod, "to_nhwc"):
            mod.to_nhwc = to_nhwc


def seeded(module, seed):
    sd = module.state_dict()
    fill_state_dict_(sd, seed)
    for k in sd:
        if k.endswith("router.3.weight"):
            sd[k] *= 3
    module.load_state_dict(sd)
    return module.eval(), {"m." + k: v.clone().float() for k, v in sd.items()}


def report(name, y, ref, sim):
    e, n = (y.float() - ref).abs(), (sim - ref).abs()
    flag = "OK " if float(e.mean()) <= 2 * float(n.mean()) + 2e-4 * float(ref.pow(2).mean().sqrt()) and float(e.max()) <= 3 * float(n.max()) + 2e-3 else "BAD"
    print(f"{flag} {name:44s} mean err {float(e.mean()):.2e} (noise {float(n.mean()):.2e})  max {float(e.max()):.2e} (noise {float(n.max()):.2e})")
    return flag == "OK "


def main():
    install()
    ok = True
    g = torch.Generator().manual_seed(0)

    def run(mod, fn, x, name):
        nonlocal ok
        with torch.no_grad():
            y = mod.fwd_nhwc(x.half().permute(0, 2, 3, 1).contiguous()).float().permute(0, 3, 1, 2)
        ref = fn(x)
        with O.fp16_storage(), O.fp16_weights():
            sim = fn(x)
        ok &= report(name, y, ref, sim)

    for dim, nh, H, W in [(64, 8, 20, 20), (128, 8, 10, 12), (64, 8, 7, 5)]:
        x = torch.randn((2, dim, H, W), generator=g).half().float()
        m, sd = seeded(mot._LocalConvTransformerExpert(dim, nh), 1)
        run(m, lambda t: O.mot_local_expert(sd, "m", t, nh), x, f"LocalConv {dim} {H}x{W}")
        m, sd = seeded(mot._LocalConvTransformerExpert(dim, nh, local_window_size=4), 2)
        run(m, lambda t: O.mot_local_expert(sd, "m", t, nh, 4), x, f"LocalConv windowed {dim} {H}x{W}")
        for shift in (0, 1):
            m, sd = seeded(mot._WindowTransformerExpert(dim, nh, 7, shift_size=shift), 3 + shift)
            run(m, lambda t: O.mot_window_expert(sd, "m", t, nh, 7, 3 if shift else 0), x, f"Window shift={shift} {dim} {H}x{W}")
        m, sd = seeded(mot._DeformableTransformerExpert(dim, nh), 5)
        run(m, lambda t: O.mot_deform_expert(sd, "m", t, nh), x, f"Deformable {dim} {H}x{W}")
        m, sd = seeded(mot.MoTBlock(dim, nh, 2, temperature=0.8), 7)
        run(m, lambda t: O.mot_block(sd, "m", t, nh, 2), x, f"MoTBlock {dim} {H}x{W}")
    m, sd = seeded(mot.C2fMoT(64, 128, 2, 8, 2), 8)
    x = torch.randn((2, 64, 14, 14), generator=g).half().float()
    run(m, lambda t: O.layer_c2f_mot(sd, "m", t, 64, 128, 2, 8, 2), x, "C2fMoT n=2")
    for dim, heads, H, W in [(32, 3, 40, 40), (32, 3, 12, 12), (32, 3, 22, 22), (64, 3, 24, 20), (96, 6, 9, 9)]:
        x = torch.randn((2, dim, H, W), generator=g).half().float()
        hd, hpg = max(dim // heads, 16), heads // 3
        for cls, fn, nm in ((moa._LocalAttnHead, O.moa_local_head, "local"), (moa._RegionalAttnHead, O.moa_region_head, "regional"),
                            (moa._GlobalAttnHead, O.moa_global_head, "global")):
            m, sd = seeded(cls(dim, hpg, hd), 1)
            m.fwd_nhwc = (lambda mm: lambda t: ops.groupnorm(mm.head_raw(t), mm.norm.num_groups, *mm.get_pack()["norm"]))(m)
            run(m, lambda t: fn(sd, "m", t, hpg, hd), x, f"MoA {nm} head {dim} {H}x{W}")
        m, sd = seeded(moa.MoABlock(dim, heads, temperature=0.8), 5)
        run(m, lambda t: O.moa_block(sd, "m", t, heads, 0.8), x, f"MoABlock {dim} {H}x{W}")
    m, sd = seeded(moa.C2fMoA(64, 64, 1, 3, 2.0, 0.8, True), 9)
    x = torch.randn((2, 64, 24, 24), generator=g).half().float()
    run(m, lambda t: O.layer_c2f_moa(sd, "m", t, 64, 64, 1, 3, 2.0, 0.8, True), x, "C2fMoA")
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
