"""GPU parity of the Mixture-of-Transformers / Mixture-of-Attention kernels and modules against the
CPU oracle: every new C-ABI kernel alone on seeded inputs, then each expert / head / router / block module against the
oracle function of the same name with identical (key-seeded) weights."""
import math

import pytest
import torch
import torch.nn.functional as F

from _util import assert_close, assert_within_noise
from oracle import yolo_master_oracle as O
from yolo_master_b200 import ops
from yolo_master_b200.nn.modules import moa as MA
from yolo_master_b200.nn.modules import mot as MT
from yolo_master_b200.utils.synth import fill_state_dict_

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def dev_nhwc(x_nchw):
    """fp32 NCHW (already fp16-representable) -> fp16 NHWC contiguous on the GPU."""
    return x_nchw.permute(0, 2, 3, 1).contiguous().half().to(DEV)


def back(y_nhwc):
    return y_nhwc.float().cpu().permute(0, 3, 1, 2)


def h16(x):
    return x.half().float()


# ---------------------------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------------------------
def test_ew_ops():
    a, b = h16(rnd(2, 24, 5, 7, seed=1)), h16(rnd(2, 24, 5, 7, seed=2))
    chan = torch.rand(24, generator=torch.Generator().manual_seed(3))
    tok = torch.rand((2, 5, 7, 3), generator=torch.Generator().manual_seed(4))
    da, db = dev_nhwc(a), dev_nhwc(b)
    y = ops.ew(ops.EW_SCALE_RES, a=da, b=db, p0=chan.to(DEV))
    assert_close(back(y), a + chan.view(1, -1, 1, 1) * b, what="scale_res")
    y = ops.ew(ops.EW_TOKEN_ACC, a=da, b=db, tok=tok.to(DEV).view(-1), ldt=3, toff=1)
    assert_close(back(y), a + tok[..., 1].unsqueeze(1) * b, what="token_acc")
    y = ops.ew(ops.EW_TOKEN_ACC, a=None, b=db, tok=tok.to(DEV).view(-1), ldt=3, toff=2)
    assert_close(back(y), tok[..., 2].unsqueeze(1) * b, what="token_acc (no a)")
    assert_close(back(ops.ew(ops.EW_GLU, a=da, b=db)), torch.sigmoid(a) * b, what="glu")
    assert_close(back(ops.ew(ops.EW_GELU, a=da)), F.gelu(a), what="gelu")
    t = torch.tensor([0.3])
    assert_close(back(ops.ew(ops.EW_LERP, a=da, b=db, p0=t.to(DEV))), 0.3 * a + 0.7 * b, what="lerp")
    # channel-slice views (pitch != C)
    wide = dev_nhwc(h16(rnd(2, 48, 5, 7, seed=5)))
    y = ops.ew(ops.EW_GLU, a=wide[..., :24], b=wide[..., 24:])
    w32 = back(wide)
    assert_close(back(y), torch.sigmoid(w32[:, :24]) * w32[:, 24:], what="glu on slices")


@pytest.mark.parametrize("C,G,H,W", [(64, 8, 9, 11), (32, 8, 40, 40), (24, 4, 3, 5)])
def test_groupnorm(C, G, H, W):
    x = h16(rnd(3, C, H, W, seed=C) * 2 + 0.5)
    g, b = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    ref = F.group_norm(x, G, g, b, 1e-5)
    y = ops.groupnorm(dev_nhwc(x), G, g.to(DEV), b.to(DEV))
    assert_close(back(y), ref, atol=2e-3, what="groupnorm")
    add = h16(rnd(3, C, H, W, seed=7))
    tok = torch.rand((3, H, W, 3))
    y = ops.groupnorm(dev_nhwc(x), G, g.to(DEV), b.to(DEV), act=True, tok=tok.to(DEV).view(-1), ldt=3, toff=1, add=dev_nhwc(add))
    assert_close(back(y), add + tok[..., 1].unsqueeze(1) * F.silu(ref), atol=2e-3, what="groupnorm+silu+tok+add")


@pytest.mark.parametrize("C", [64, 128, 256])
def test_layernorm(C):
    x = h16(rnd(2, C, 6, 5, seed=C) * 1.5 + 0.3)
    g, b = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    ref = F.layer_norm(x.permute(0, 2, 3, 1), (C,), g, b, 1e-5).permute(0, 3, 1, 2)
    assert_close(back(ops.layernorm(dev_nhwc(x), g.to(DEV), b.to(DEV))), ref, atol=2e-3, what="layernorm")


def _heads(t, nh, hd):   # NCHW (C = nh*hd) -> (B, nh, N, hd)
    B, C, H, W = t.shape
    return t.reshape(B, nh, hd, H * W).transpose(2, 3)


@pytest.mark.parametrize("hd,nh,H,W,h2,w2", [(8, 8, 13, 10, 13, 10), (16, 2, 20, 20, 10, 10), (24, 1, 9, 7, 4, 3), (32, 4, 8, 9, 8, 9)])
def test_attn_small(hd, nh, H, W, h2, w2):
    q, k, v = h16(rnd(2, nh * hd, H, W, seed=1)), h16(rnd(2, nh * hd, h2, w2, seed=2)), h16(rnd(2, nh * hd, h2, w2, seed=3))
    ref = O._sdpa(_heads(q, nh, hd), _heads(k, nh, hd), _heads(v, nh, hd), hd ** -0.5).transpose(2, 3).reshape(2, nh * hd, H, W)
    y = ops.attn_small(dev_nhwc(q), dev_nhwc(k), dev_nhwc(v), nh, hd, hd ** -0.5)
    assert_close(back(y), ref, what=f"attn_small hd={hd}")


@pytest.mark.parametrize("hd,nh,H,W,win,shift,padded", [(8, 8, 10, 9, 7, 0, False), (16, 2, 16, 20, 7, 3, True), (8, 4, 14, 14, 7, 3, True),
                                                      (24, 1, 5, 6, 5, 0, False), (32, 2, 20, 20, 7, 0, True), (16, 1, 3, 4, 3, 0, False)])
def test_attn_window(hd, nh, H, W, win, shift, padded):
    """Reference token mapping: pad (with the pad vectors) -> roll(-shift) -> partition -> SDPA -> reverse -> roll(+shift) -> crop."""
    C = nh * hd
    q, k, v = (h16(rnd(2, C, H, W, seed=s)) for s in (1, 2, 3))
    padk = h16(rnd(C, seed=4)) if padded else torch.zeros(C)
    padv = h16(rnd(C, seed=5)) if padded else torch.zeros(C)

    def prep(t, padvec):
        t = t.permute(0, 2, 3, 1)
        Hp, Wp = math.ceil(H / win) * win, math.ceil(W / win) * win
        full = padvec.view(1, 1, 1, C).expand(2, Hp, Wp, C).clone()
        full[:, :H, :W] = t
        if shift:
            full = torch.roll(full, (-shift, -shift), (1, 2))
        return O._win_part(full, win).reshape(-1, win * win, nh, hd).permute(0, 2, 1, 3), Hp, Wp

    qw, Hp, Wp = prep(q, torch.zeros(C))
    kw, _, _ = prep(k, padk)
    vw, _, _ = prep(v, padv)
    o = O._sdpa(qw, kw, vw, hd ** -0.5).transpose(1, 2).reshape(-1, win * win, C)
    o = O._win_rev(o, win, Hp, Wp)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    ref = o[:, :H, :W].permute(0, 3, 1, 2)
    y = ops.attn_window(dev_nhwc(q), dev_nhwc(k), dev_nhwc(v), nh, hd, win, shift, hd ** -0.5,
                        padk=padk.half().to(DEV) if padded else None, padv=padv.half().to(DEV) if padded else None)
    assert_close(back(y), ref, what=f"attn_window hd={hd} win={win} shift={shift}")


@pytest.mark.parametrize("nh,hd,H,W", [(8, 8, 12, 9), (8, 16, 7, 7), (4, 32, 5, 6)])
def test_deform_sample(nh, hd, H, W):
    B, npt, C = 2, 4, nh * hd
    N = H * W
    v = h16(rnd(B, C, H, W, seed=1))
    oa = rnd(B, H, W, nh * npt * 3, seed=2) * 1.5
    off = oa[..., :nh * npt * 2].reshape(B, N, nh, npt, 2).tanh()
    aw = oa[..., nh * npt * 2:].reshape(B, N, nh, npt).softmax(-1)
    idx = torch.arange(N)
    ref_pt = torch.stack([(idx % W).float() / max(W - 1, 1) * 2 - 1, (idx // W).float() / max(H - 1, 1) * 2 - 1], -1)[None, :, None, None, :]
    locs = (ref_pt + off * 0.25).clamp(-1, 1)
    samp = F.grid_sample(v.reshape(B * nh, hd, H, W), locs.permute(0, 2, 1, 3, 4).reshape(B * nh, N, npt, 2), mode="bilinear",
                         padding_mode="zeros", align_corners=True)
    ref = (aw.unsqueeze(-1) * samp.reshape(B, nh, hd, N, npt).permute(0, 3, 1, 4, 2)).sum(3).reshape(B, H, W, C).permute(0, 3, 1, 2)
    y = ops.deform_sample(oa.to(DEV).contiguous(), dev_nhwc(v), nh, hd, npt, True)
    assert_close(back(y), ref, atol=2e-3, what="deform_sample")


@pytest.mark.parametrize("H,W,h,w", [(40, 40, 20, 20), (13, 9, 6, 4), (7, 7, 3, 3)])
def test_adaptive_avgpool(H, W, h, w):
    x = h16(rnd(2, 32, H, W, seed=3))
    assert_close(back(ops.adaptive_avgpool(dev_nhwc(x), h, w)), F.adaptive_avg_pool2d(x, (h, w)), what="adaptive_avgpool")


@pytest.mark.parametrize("hd,nh,H,W", [(16, 1, 30, 25), (21, 1, 12, 11), (8, 2, 40, 40)])
def test_linear_attn(hd, nh, H, W):
    hdp = (hd + 7) // 8 * 8
    q, k, v = (h16(rnd(2, nh, H * W, hd, seed=s)) for s in (1, 2, 3))
    rf = torch.linalg.qr(rnd(hd, hd, seed=4))[0].contiguous()
    ref = O._linear_attn(q, k, v, rf)                                     # (B, nh, N, hd)

    def lay(t):   # -> (B,H,W,nh*hdp) fp16 with zero-padded heads
        z = torch.zeros(2, nh, H * W, hdp)
        z[..., :hd] = t
        return z.permute(0, 2, 1, 3).reshape(2, H, W, nh * hdp).half().to(DEV).contiguous()

    y = ops.linear_attn(lay(q), lay(k), lay(v), nh, hdp, hd, rf.to(DEV)).float().cpu().reshape(2, H * W, nh, hdp)[..., :hd]
    assert_close(y.permute(0, 2, 1, 3), ref, atol=2e-3, what="linear_attn")


# ---------------------------------------------------------------------------------------------------------------------
# modules vs oracle functions (identical key-seeded weights)
# ---------------------------------------------------------------------------------------------------------------------
def seeded(module, seed):
    sd = module.state_dict()
    fill_state_dict_(sd, seed)
    for k in sd:   # spread router logits so that top-k margins are not all tiny / all huge
        if k.endswith("router.3.weight"):
            sd[k] *= 3
    module.load_state_dict(sd)
    return module.to(DEV).eval(), {"m." + k: v.clone().float() for k, v in sd.items()}


def run_mod(mod, x):
    with torch.no_grad():
        return back(mod.fwd_nhwc(dev_nhwc(x)))


def check_mod(y, fn, x, what, outlier_frac=0.0):
    ref = fn(x)
    with O.fp16_storage(), O.fp16_weights():
        sim = fn(x)
    assert_within_noise(y, ref, sim, what=what, outlier_frac=outlier_frac)


@pytest.mark.parametrize("dim,nh,H,W", [(64, 8, 20, 20), (128, 8, 10, 12), (64, 8, 7, 5)])
def test_mot_experts(dim, nh, H, W):
    x = h16(rnd(2, dim, H, W, seed=dim + H))
    m, sd = seeded(MT._LocalConvTransformerExpert(dim, nh), 1)
    check_mod(run_mod(m, x), lambda t: O.mot_local_expert(sd, "m", t, nh), x, "LocalConv expert")
    m, sd = seeded(MT._LocalConvTransformerExpert(dim, nh, local_window_size=4), 2)
    check_mod(run_mod(m, x), lambda t: O.mot_local_expert(sd, "m", t, nh, 4), x, "LocalConv expert (windowed)")
    for shift in (0, 1):
        m, sd = seeded(MT._WindowTransformerExpert(dim, nh, 7, shift_size=shift), 3 + shift)
        check_mod(run_mod(m, x), lambda t: O.mot_window_expert(sd, "m", t, nh, 7, 3 if shift else 0), x, f"Window expert shift={shift}")
    m, sd = seeded(MT._DeformableTransformerExpert(dim, nh), 5)
    check_mod(run_mod(m, x), lambda t: O.mot_deform_expert(sd, "m", t, nh), x, "Deformable expert")


@pytest.mark.parametrize("dim,topk,H,W", [(64, 2, 40, 40), (128, 1, 20, 20), (256, 2, 9, 7)])
def test_mot_router_indices_exact(dim, topk, H, W):
    m, sd = seeded(MT._MoTRouter(dim, 3, topk, temperature=0.8), 11)
    x = h16(rnd(3, dim, H, W, seed=dim))
    with torch.no_grad():
        w, idx = m.route(dev_nhwc(x))
    wr, ir, logits = O.mot_router(sd, "m", x, topk)
    probs = F.softmax(logits / sd["m.temperature"], dim=1).topk(min(topk + 1, 3), dim=1)[0]
    margin = (probs[:, topk - 1] - probs[:, topk]) if topk < 3 else torch.ones_like(probs[:, 0])
    if topk == 2:
        margin = torch.minimum(margin, probs[:, 0] - probs[:, 1])        # the ORDER of the two winners must be stable too
    safe = margin > 1e-4
    assert safe.float().mean() > 0.97
    got = idx.cpu().permute(0, 3, 1, 2).long()
    assert torch.equal(got[safe.unsqueeze(1).expand_as(got)], ir[safe.unsqueeze(1).expand_as(ir)])   # bit-exact top-k indices
    gw = w.cpu().permute(0, 3, 1, 2)
    sw = safe.unsqueeze(1).expand_as(gw)
    torch.testing.assert_close(gw[sw], wr[sw], atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("dim,nh,topk,H,W", [(64, 8, 2, 20, 20), (128, 8, 1, 10, 10)])
def test_mot_block_and_c2f(dim, nh, topk, H, W):
    x = h16(rnd(2, dim, H, W, seed=21))
    m, sd = seeded(MT.MoTBlock(dim, nh, topk, temperature=0.8), 7)
    check_mod(run_mod(m, x), lambda t: O.mot_block(sd, "m", t, nh, topk), x, "MoTBlock")
    m, sd = seeded(MT.C2fMoT(dim, 2 * dim, 2, 8, topk), 8)
    xin = h16(rnd(2, dim, H, W, seed=22))
    check_mod(run_mod(m, xin), lambda t: O.layer_c2f_mot(sd, "m", t, dim, 2 * dim, 2, 8, topk), xin, "C2fMoT (n=2: shifted window block)",
              outlier_frac=0.02)   # the 2nd block's router sees the 1st block's fp16 output: near-tie tokens may flip


@pytest.mark.parametrize("dim,heads,H,W", [(32, 3, 40, 40), (32, 3, 12, 12), (32, 3, 22, 22), (64, 3, 24, 20), (96, 6, 9, 9)])
def test_moa_heads_and_block(dim, heads, H, W):
    """N = 1600 -> linear attention, 144 / 81 -> exact, 484 -> blend window; dim 64 with 3 heads -> head_dim 21 (padded to 24)."""
    x = h16(rnd(2, dim, H, W, seed=dim + H))
    hd, hpg = max(dim // heads, 16), heads // 3
    m, sd = seeded(MA._LocalAttnHead(dim, hpg, hd), 1)
    check_mod(back(ops.groupnorm(m.head_raw(dev_nhwc(x)), m.norm.num_groups, *m.get_pack()["norm"])),
              lambda t: O.moa_local_head(sd, "m", t, hpg, hd), x, "MoA local head")
    m, sd = seeded(MA._RegionalAttnHead(dim, hpg, hd), 2)
    check_mod(back(ops.groupnorm(m.head_raw(dev_nhwc(x)), m.norm.num_groups, *m.get_pack()["norm"])),
              lambda t: O.moa_region_head(sd, "m", t, hpg, hd), x, "MoA regional head")
    m, sd = seeded(MA._GlobalAttnHead(dim, hpg, hd), 3)
    check_mod(back(ops.groupnorm(m.head_raw(dev_nhwc(x)), m.norm.num_groups, *m.get_pack()["norm"])),
              lambda t: O.moa_global_head(sd, "m", t, hpg, hd), x, "MoA global head")
    m, sd = seeded(MA._MoARouter(dim, 3, temperature=0.8), 4)
    with torch.no_grad():
        w = m.route(dev_nhwc(x)).cpu().permute(0, 3, 1, 2)
    torch.testing.assert_close(w, O.moa_router(sd, "m", x, 0.8), atol=2e-5, rtol=1e-4)
    m, sd = seeded(MA.MoABlock(dim, heads, temperature=0.8), 5)
    check_mod(run_mod(m, x), lambda t: O.moa_block(sd, "m", t, heads, 0.8), x, "MoABlock")


def test_c2f_moa():
    m, sd = seeded(MA.C2fMoA(64, 64, 1, 3, 2.0, 0.8, True), 9)
    x = h16(rnd(2, 64, 24, 24, seed=31))
    check_mod(run_mod(m, x), lambda t: O.layer_c2f_moa(sd, "m", t, 64, 64, 1, 3, 2.0, 0.8, True), x, "C2fMoA")
