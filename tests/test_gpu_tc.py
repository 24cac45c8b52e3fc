"""tcgen05 kernels (TMEM accumulators, swizzled smem operands) through the C ABI: plain GEMM and the ES-MoE dispatch."""
import pytest
import torch

from _util import assert_close
from oracle.moe_dispatch_oracle import compute_sparse_experts_batched, conv1x1_experts

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 64, 128), (1000, 128, 64), (6400, 64, 192), (777, 256, 256),
                                   (512, 192, 72), (128, 384, 384), (300, 80, 256)])
def test_tc_gemm_matches_fp32(M, N, K):
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((M, K), generator=g).half()
    b = (torch.randn((N, K), generator=g) / K ** 0.5).half()
    bias = torch.randn((N,), generator=g)
    out = ops.tc_gemm_nt(a.to(DEV), b.to(DEV), bias.to(DEV))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    assert_close(out, ref, what=f"tc_gemm {M}x{N}x{K}")


def test_tc_gemm_epilogue_and_pitched_views():
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(1)
    M, N, K = 640, 64, 128
    abuf = torch.randn((M, K + 64), generator=g).half().to(DEV)
    a = abuf[:, 32:32 + K]                      # channel slice of a wider NHWC buffer (pitch 192)
    b = (torch.randn((N, K), generator=g) / K ** 0.5).half().to(DEV)
    res = torch.randn((M, N), generator=g).half().to(DEV)
    obuf = torch.zeros((M, 2 * N), dtype=torch.float16, device=DEV)
    out = ops.tc_gemm_nt(a, b, None, res, act=True, out=obuf[:, N:])
    ref = torch.nn.functional.silu(a.float().cpu() @ b.float().cpu().t()) + res.float().cpu()
    assert_close(out, ref, what="tc_gemm silu+res")
    assert float(obuf[:, :N].abs().max()) == 0.0


@pytest.mark.parametrize("impl", ["v3", "v2", "v1"])
@pytest.mark.parametrize("B,hw,C,E,k", [(8, (16, 16), 64, 4, 2), (64, (32, 32), 256, 8, 2), (5, (9, 13), 128, 8, 2), (6, (20, 20), 256, 16, 1),
                                        (37, (16, 8), 128, 8, 2), (3, (64, 64), 256, 8, 2), (10, (16, 16), 128, 8, 2), (5, (32, 16), 256, 8, 1),
                                        (151, (16, 16), 256, 8, 2)])
def test_moe_dispatch_vs_oracle(B, hw, C, E, k, impl):
    """C5 configuration (x=(64,256,32,32), 8 experts, top-2) and ragged variants against the restated dispatcher."""
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(B * 7 + E)
    x = torch.randn((B, C, *hw), generator=g).half()
    W = (torch.randn((E, C, C), generator=g) / C ** 0.5).half()
    idx = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(B)]).int()
    w = torch.rand((B, k), generator=g)
    w = w / w.sum(1, keepdim=True)
    w[0, -1] = 0.005                                   # below the 0.01 eval threshold: that route must be dropped
    if B > 2:
        w[2, :] = 0.004                                # an image with no live route at all: its output is exactly zero
    ops.DISPATCH_IMPL = impl
    try:
        out = ops.moe_dispatch(x.to(DEV).permute(0, 2, 3, 1).contiguous(), W.to(DEV), idx.to(DEV), w.to(DEV))
        torch.cuda.synchronize()
    finally:
        ops.DISPATCH_IMPL = "v3"
    ref = compute_sparse_experts_batched(x.float(), conv1x1_experts(W.float()), w, idx.long(), C)
    # the reference rounds every expert output to fp16 before weighting (utils.py:200-203); with two experts of opposite
    # sign that rounding is visible on the (small) sum, so a handful of elements may exceed the per-element bound vs fp32
    assert_close(out.permute(0, 3, 1, 2), ref, max_bad_frac=2e-5, what="moe_dispatch vs fp32 oracle")
    experts16 = [lambda t, f=f: f(t).half().float() for f in conv1x1_experts(W.float())]   # fp16 expert outputs, as the reference
    ref16 = compute_sparse_experts_batched(x.float(), experts16, w, idx.long(), C)
    # a single accumulation-order flip of an fp16 rounding on a large expert output survives cancellation: allow ~1e-6
    assert_close(out.permute(0, 3, 1, 2), ref16, max_bad_frac=2e-6, what="moe_dispatch vs fp16-expert-output oracle")


@pytest.fixture(params=[0, 1, 2, 3], ids=["rowwise", "chunked", "tmemP", "chunked+tmemP"])
def softmax_mode(request):
    """Both softmax schedules of the tcgen05 attention kernel (ym_set_attention_chunked)."""
    from yolo_master_b200 import _lib
    prev = _lib.load().ym_set_attention_chunked(request.param)
    yield request.param
    assert _lib.load().ym_set_attention_chunked(prev) == request.param


@pytest.mark.parametrize("N,heads,dv,batch", [(64, 1, 32, 1), (128, 2, 32, 2), (400, 2, 32, 3), (1600, 2, 32, 2), (221, 4, 32, 2),
                                              (400, 2, 64, 2), (100, 2, 64, 1), (6400, 2, 32, 1), (40, 1, 32, 1), (17, 2, 32, 2)])
def test_tc_attention_strict(N, heads, dv, batch, softmax_mode):
    """tcgen05 attention kernel alone vs fp32 softmax attention (strict tolerance), incl. ragged N and d_v = 64."""
    from yolo_master_b200 import _lib, ops
    hs = 64 + dv
    g = torch.Generator().manual_seed(N + dv)
    qkv = torch.randn((batch, N, 1, heads * hs), generator=g).half()
    out = torch.empty((batch, N, 1, heads * dv), dtype=torch.float16, device=DEV)
    _lib.check(_lib.load().ym_attention_fwd_tc(qkv.to(DEV).data_ptr(), heads * hs, batch, N, heads, hs, 0, 32, 64, 32, dv,
                                               32 ** -0.5, out.data_ptr(), heads * dv, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    t = qkv.float().view(batch, N, heads, hs).permute(0, 2, 1, 3)
    q, k, v = t[..., :32], t[..., 32:64], t[..., 64:]
    ref = (torch.softmax((q * 32 ** -0.5) @ k.transpose(-1, -2), -1) @ v).permute(0, 2, 1, 3).reshape(batch, N, 1, heads * dv)
    assert_close(out, ref, what=f"tc attention N={N} dv={dv}")


def test_tc_attention_large_logits_lazy_rescale(softmax_mode):
    """Rows whose running max keeps growing exercise the lazy O rescaling path (and, chunked, the in-row P rescale)."""
    from yolo_master_b200 import _lib
    N, heads, dv, hs = 512, 1, 32, 96
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn((1, N, 1, hs), generator=g)
    qkv[..., 32:64] *= torch.linspace(0.5, 6.0, N).view(1, N, 1, 1)     # key norms increase along the sequence
    qkv = qkv.half()
    out = torch.empty((1, N, 1, dv), dtype=torch.float16, device=DEV)
    _lib.check(_lib.load().ym_attention_fwd_tc(qkv.to(DEV).data_ptr(), hs, 1, N, heads, hs, 0, 32, 64, 32, dv, 32 ** -0.5,
                                               out.data_ptr(), dv, torch.cuda.current_stream().cuda_stream))
    t = qkv.float().view(1, N, 1, hs).permute(0, 2, 1, 3)
    q, k, v = t[..., :32], t[..., 32:64], t[..., 64:]
    ref = (torch.softmax((q * 32 ** -0.5) @ k.transpose(-1, -2), -1) @ v).permute(0, 2, 1, 3).reshape(1, N, 1, dv)
    assert_close(out, ref, what="tc attention lazy rescale")


def _attn_ref(qkv, batch, N, heads, hs, dv):
    t = qkv.float().view(batch, N, heads, hs).permute(0, 2, 1, 3)
    q, k, v = t[..., :32], t[..., 32:64], t[..., 64:]
    return (torch.softmax((q * 32 ** -0.5) @ k.transpose(-1, -2), -1) @ v).permute(0, 2, 1, 3).reshape(batch, N, 1, heads * dv)


@pytest.mark.parametrize("N,heads,dv,batch", [(64, 1, 32, 1), (128, 2, 32, 2), (256, 2, 32, 1), (257, 2, 32, 2), (400, 2, 32, 3), (1600, 2, 32, 2),
                                              (221, 4, 32, 2), (400, 2, 64, 2), (100, 2, 64, 1), (6400, 2, 32, 2), (40, 1, 32, 1), (17, 2, 32, 2),
                                              (129, 1, 32, 1), (385, 3, 32, 1), (449, 2, 64, 2), (3200, 1, 32, 1)])
def test_tc_attention2_strict(N, heads, dv, batch):
    """Warp-specialised tcgen05 / TMA attention kernel (ym_attention_fwd_tc2) alone vs fp32 softmax attention, strict tolerance:
    ragged N (query-tile and key-tile tails, an absent second query tile), one and several heads, d_v = 64, the P3 length - with one
    and with two query tiles per CTA (bit-identical to each other: a row sees the same keys in the same order)."""
    from yolo_master_b200 import _lib
    hs = 64 + dv
    g = torch.Generator().manual_seed(N + dv)
    qkv = torch.randn((batch, N, 1, heads * hs), generator=g).half()
    outs = []
    for qt, variant in ((2, None), (1, None), (2, 0), (2, 7)):
        prev = _lib.load().ym_set_attention2_qtiles(qt)
        prev_var = _lib.load().ym_set_attention2_variant(variant) if variant is not None else None
        try:
            out = torch.full((batch, N, 1, heads * dv), float("nan"), dtype=torch.float16, device=DEV)
            _lib.check(_lib.load().ym_attention_fwd_tc2(qkv.to(DEV).data_ptr(), heads * hs, batch, N, heads, hs, 0, 32, 64, 32, dv,
                                                        32 ** -0.5, out.data_ptr(), heads * dv, torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
        finally:
            _lib.load().ym_set_attention2_qtiles(prev)
            if prev_var is not None:
                _lib.load().ym_set_attention2_variant(prev_var)
        assert_close(out, _attn_ref(qkv, batch, N, heads, hs, dv), what=f"tc attention2 N={N} dv={dv} q_tiles={qt} variant={variant}")
        outs.append(out)
    assert all(torch.equal(outs[0], o) for o in outs[1:]), "query tiles per CTA / scheduling variants must agree bit for bit"


def test_tc_attention2_large_logits_lazy_rescale_and_impl_agreement():
    """Growing row maxima exercise the lazy O rescale; the three kernels behind ym_attention_fwd agree within the strict tolerance."""
    from yolo_master_b200 import _lib, ops
    N, heads, dv, hs = 512, 1, 32, 96
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn((1, N, 1, hs), generator=g)
    qkv[..., 32:64] *= torch.linspace(0.5, 6.0, N).view(1, N, 1, 1)     # key norms increase along the sequence
    qkv = qkv.half()
    ref = _attn_ref(qkv, 1, N, heads, hs, dv)
    outs = []
    for impl in (2, 1, 0):
        prev = _lib.load().ym_set_attention_impl(impl)
        try:
            outs.append(ops.attention(qkv.to(DEV), 1, N, heads, hs, 0, 32, 64, 32, dv, 32 ** -0.5))
            torch.cuda.synchronize()
        finally:
            _lib.load().ym_set_attention_impl(prev)
        assert_close(outs[-1].reshape(ref.shape), ref, what=f"attention impl {impl} lazy rescale")
    assert _lib.load().ym_set_attention_impl(2) == 2, "the warp-specialised kernel is the default behind ym_attention_fwd"


def test_tc_attention2_area_layout_interleaved_qkv():
    """The layout AAttn really uses (block.py:1708-1722): per head [q | k | v] interleaved in one (B, N, heads*3*32) buffer, output written
    into a channel slice of a wider buffer (pitch > heads*d_v)."""
    from yolo_master_b200 import _lib
    B, N, heads, hd = 3, 1600, 2, 32
    g = torch.Generator().manual_seed(11)
    qkv = torch.randn((B, N, 1, heads * 3 * hd), generator=g).half()
    wide = torch.zeros((B, N, 1, 256), dtype=torch.float16, device=DEV)
    out = wide[..., 64:64 + heads * hd]
    _lib.check(_lib.load().ym_attention_fwd_tc2(qkv.to(DEV).data_ptr(), heads * 3 * hd, B, N, heads, 3 * hd, 0, hd, 2 * hd, hd, hd,
                                                hd ** -0.5, out.data_ptr(), 256, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert_close(out, _attn_ref(qkv, B, N, heads, 3 * hd, hd), what="tc attention2 AAttn layout")
    assert float(wide[..., :64].abs().max()) == 0 and float(wide[..., 128:].abs().max()) == 0, "wrote outside its channel slice"


@pytest.mark.parametrize("B,HW,C,HID,E,topk", [(4, 6400, 64, 128, 4, 2), (3, 1600, 64, 128, 8, 2), (5, 400, 128, 256, 16, 2), (2, 100, 64, 128, 4, 1),
                                               (1, 129, 128, 256, 4, 2), (7, 37, 64, 128, 4, 2)])
def test_moe_ffn_tc_matches_fp32_chain(B, HW, C, HID, E, topk):
    """ym_moe_ffn (tcgen05, hidden kept in tensor memory) stage by stage against fp32 torch on the same fp16 inputs: GroupNorm-1 affine
    from the stage-1 statistics, o = SiLU(GN1(h)) W2^T and the GroupNorm-2 affine from the stage-2 statistics (strict tolerance), ragged
    HW (row-tile tails, empty trailing strips), both width pairs, a channel-slice input view."""
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(HW + C)
    wide = torch.randn((B, HW, 1, C + 64), generator=g).half().to(DEV)
    x = wide[..., 32:32 + C]                                                  # pitch C + 64, 64-byte offset
    w1 = (torch.randn((E, HID, C), generator=g) / C ** 0.5).half().to(DEV)
    w2 = (torch.randn((E, C, HID), generator=g) / HID ** 0.5).half().to(DEV)
    ridx = torch.stack([torch.randperm(E, generator=g)[:topk] for _ in range(B)]).int().view(-1).to(DEV)
    rw = torch.rand((B * topk,), generator=g).to(DEV)
    G1, G2, eps = 8, 8, 1e-5
    ga1, be1 = (0.6 + 0.8 * torch.rand((E, HID), generator=g)).to(DEV), (0.2 * torch.randn((E, HID), generator=g)).to(DEV)
    ga2, be2 = (0.6 + 0.8 * torch.rand((E, C), generator=g)).to(DEV), (0.2 * torch.randn((E, C), generator=g)).to(DEV)
    P = B * topk
    st1, strips = ops.moe_ffn_stats(x, topk, w1, ridx)
    sc1, sh1 = ops.gn_finalize_tiles(st1, P, strips, G1, HID, HW * (HID // G1), eps, ga1, be1, ridx)
    o, st2 = ops.moe_ffn_fused(x, topk, w1, w2, ridx, sc1, sh1, strips)
    sc2, sh2 = ops.gn_finalize_tiles(st2, P, strips, G2, C, HW * (C // G2), eps, ga2, be2, ridx, route_w=rw)
    torch.cuda.synchronize()
    xf = x.float().reshape(B, HW, C)
    for p in range(P):
        e = int(ridx[p])
        h = (xf[p // topk] @ w1[e].float().t()).half().float()                 # the fp16 hidden of the reference's fp16 execution
        hg = h.view(HW, G1, HID // G1)
        mean, var = hg.mean((0, 2)), hg.var((0, 2), unbiased=False)
        rstd = 1.0 / torch.sqrt(var + eps)
        s1 = rstd.repeat_interleave(HID // G1) * ga1[e]
        t1 = be1[e] - mean.repeat_interleave(HID // G1) * s1
        assert_close(sc1[p], s1, atol=1e-4, rtol=1e-3, what=f"GN1 scale p={p}")
        assert_close(sh1[p], t1, atol=1e-3, rtol=1e-3, what=f"GN1 shift p={p}")
        a = torch.nn.functional.silu(h * s1 + t1).half().float()
        oref = a @ w2[e].float().t()
        assert_close(o[p], oref, what=f"o p={p}")
        og = o[p].float().view(HW, G2, C // G2)
        m2, v2 = og.mean((0, 2)), og.var((0, 2), unbiased=False)
        r2 = 1.0 / torch.sqrt(v2 + eps)
        s2 = rw[p] * r2.repeat_interleave(C // G2) * ga2[e]
        t2 = rw[p] * be2[e] - m2.repeat_interleave(C // G2) * s2
        assert_close(sc2[p], s2, atol=1e-4, rtol=1e-3, what=f"GN2 scale p={p}")
        assert_close(sh2[p], t2, atol=1e-3, rtol=1e-3, what=f"GN2 shift p={p}")
    assert float(wide[..., :32].float().abs().max()) > 0      # the view's neighbours are inputs only - nothing was written there


@pytest.mark.parametrize("B,HW,C,topk,res", [(3, 6400, 64, 2, True), (4, 1600, 64, 2, True), (5, 400, 128, 2, True), (2, 130, 128, 1, False), (7, 37, 64, 2, True)])
def test_moe_combine_tc_matches_fp32(B, HW, C, topk, res):
    """ym_moe_combine_tc (tcgen05 shared expert + routed sum + residual) against fp32 torch, and bit-for-bit layout checks: channel-slice
    input and output views, ragged HW."""
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(HW + C + 1)
    wide = torch.randn((B, HW, 1, C + 32), generator=g).half().to(DEV)
    x = wide[..., 16:16 + C]
    ws = (torch.randn((C, C), generator=g) / C ** 0.5).half().to(DEV)
    bs = (0.2 * torch.randn((C,), generator=g)).to(DEV)
    o = torch.randn((B * topk, HW, C), generator=g).half().to(DEV)
    sc = (0.5 * torch.rand((B * topk, C), generator=g)).to(DEV)
    sh = (0.2 * torch.randn((B * topk, C), generator=g)).to(DEV)
    owide = torch.zeros((B, HW, 1, C + 64), dtype=torch.float16, device=DEV)
    out = owide[..., 64:]
    for impl in ("tc", "mma"):
        ops.MOE_COMBINE_IMPL = impl
        try:
            owide.zero_()
            ops.moe_combine(x, ws, bs, o, sc, sh, topk, add_residual=res, out=out)
            torch.cuda.synchronize()
        finally:
            ops.MOE_COMBINE_IMPL = "tc"
        xf = x.float().reshape(B, HW, C)
        ref = torch.nn.functional.silu(xf @ ws.float().t() + bs)
        of = o.float().view(B, topk, HW, C)
        for j in range(topk):
            ref = ref + of[:, j] * sc.view(B, topk, 1, C)[:, j] + sh.view(B, topk, 1, C)[:, j]
        if res:
            ref = ref + xf
        assert_close(out.reshape(B, HW, C), ref, what=f"moe_combine {impl}")
        assert float(owide[..., :64].abs().max()) == 0, "wrote outside its channel slice"


def test_moe_ffn_tc_and_mma_chains_agree_in_the_block():
    """OptimizedMOEImproved through both expert-FFN implementations (MOE_FFN_IMPL): same routing, block outputs within the strict tolerance
    of each other (the fp32 accumulation order of the two GEMM engines differs, nothing else)."""
    from yolo_master_b200.nn.modules import moe as moe_mod
    from yolo_master_b200.utils.synth import fill_state_dict_
    m = moe_mod.OptimizedMOEImproved(64, 64, num_experts=4, top_k=2).to(DEV).eval()
    sd = m.state_dict()
    fill_state_dict_(sd, 3)
    m.load_state_dict(sd)
    x = (torch.randn((4, 40, 40, 64), generator=torch.Generator().manual_seed(1)) * 0.7).half().to(DEV)
    outs = {}
    for impl in ("tc", "mma"):
        moe_mod.MOE_FFN_IMPL = impl
        try:
            with torch.no_grad():
                outs[impl] = m.fwd_nhwc(x).clone()
                snap = {k: v.clone() for k, v in m.last_routing_snapshot.items()}
            outs[impl + "_idx"] = snap["topk_indices"]
        finally:
            moe_mod.MOE_FFN_IMPL = "tc"
    assert torch.equal(outs["tc_idx"], outs["mma_idx"])
    assert_close(outs["tc"], outs["mma"].float(), what="tc vs mma expert FFN chain")


@pytest.mark.parametrize("c1,c2,k,s,act", [
    (64, 64, 1, 1, True), (64, 192, 1, 1, False), (384, 128, 1, 1, True), (48, 64, 1, 1, True), (96, 64, 1, 1, True),
    (16, 32, 3, 2, True), (64, 64, 3, 2, True), (32, 32, 3, 1, True), (16, 8, 3, 1, True), (128, 256, 3, 2, True),
    (128, 64, 3, 1, True), (80, 80, 1, 1, True), (256, 16, 3, 1, True), (32, 16, 1, 1, True), (24, 32, 3, 1, True),
    (96, 64, 3, 1, True), (40, 64, 1, 1, True),
])
@pytest.mark.parametrize("hw", [(40, 40), (37, 23), (20, 20)])
@pytest.mark.parametrize("version", [2, 1])
def test_tc_conv_matches_oracle(c1, c2, k, s, act, hw, version):
    """TMA + tcgen05 convolution (v2: persistent warp-specialised + TMA store; v1: one tile per CTA) vs the CPU oracle
    (strict tolerance) and vs the mma.sync kernel."""
    from oracle import yolo_master_oracle as O
    from yolo_master_b200 import ops
    from yolo_master_b200.nn import modules as M
    from yolo_master_b200.utils.synth import fill_state_dict_
    m = M.Conv(c1, c2, k, s, act=act)
    sd = m.state_dict()
    fill_state_dict_(sd, c1 * 7 + c2)
    m.load_state_dict(sd)
    m.eval().to(DEV)
    sdo = {"m." + kk: v.clone().float().cpu() for kk, v in sd.items()}
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, c1, *hw), generator=g).half()
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
    assert ops.lib().ym_conv2d_tc_supported(c1, c2, k, k, s, k // 2, c1) == 1
    old = ops.CONV_IMPL
    oldv = ops.lib().ym_set_tc_conv_version(version)
    try:
        ops.CONV_IMPL = "tc"
        with torch.no_grad():
            y_tc = m(xd)
        ops.CONV_IMPL = "legacy"
        with torch.no_grad():
            y_lg = m(xd)
    finally:
        ops.CONV_IMPL = old
        ops.lib().ym_set_tc_conv_version(oldv)
    torch.cuda.synchronize()
    with O.fp16_weights():   # the reference's deployed weights: BN folded, rounded to fp16
        ref = O.conv_block(sdo, "m", x.float(), s, 1, act)
    assert_close(y_tc, ref, what=f"tc_conv({c1},{c2},{k},{s}) {hw}")
    assert_close(y_tc, O.conv_block(sdo, "m", x.float(), s, 1, act), max_bad_frac=1e-4, what="vs fp32-weight oracle")
    assert_close(y_tc, y_lg.float().cpu(), atol=2e-3, rtol=2e-3, what="tc vs mma.sync")


def test_tc_conv_slices_residual_f32():
    """Channel-sliced input/output views, fused residual, fp32 output (Detect's last 1x1)."""
    from yolo_master_b200 import ops
    from yolo_master_b200.nn.modules._base import pack_gemm_weight
    g = torch.Generator().manual_seed(5)
    B, H, W, C1, C2 = 2, 24, 20, 64, 32
    buf = torch.randn((B, H, W, 3 * C1), generator=g).half().to(DEV)
    x = buf[..., C1:2 * C1]
    w = torch.randn((C2, C1, 3, 3), generator=g) / (9 * C1) ** 0.5
    bias = torch.randn((C2,), generator=g).to(DEV)
    res = torch.randn((B, H, W, C2), generator=g).half().to(DEV)
    obuf = torch.zeros((B, H, W, 2 * C2), dtype=torch.float16, device=DEV)
    wp = pack_gemm_weight(w.to(DEV))
    ops.CONV_IMPL = "tc"
    y = ops.conv2d(x, wp, bias, C2, 3, 3, 1, 1, True, out=obuf[..., C2:], res=res)
    ref = torch.nn.functional.silu(torch.nn.functional.conv2d(x.float().cpu().permute(0, 3, 1, 2), w.half().float(), bias.cpu(), 1, 1))
    ref = ref.permute(0, 2, 3, 1) + res.float().cpu()
    assert_close(y, ref, what="tc_conv slice+res")
    assert float(obuf[..., :C2].abs().max()) == 0.0
    y32 = ops.conv2d(x, wp, bias, C2, 3, 3, 1, 1, False, out_f32=True)
    ref32 = torch.nn.functional.conv2d(x.float().cpu().permute(0, 3, 1, 2), w.half().float(), bias.cpu(), 1, 1).permute(0, 2, 3, 1)
    assert y32.dtype == torch.float32
    assert_close(y32, ref32, what="tc_conv f32 out")


def test_moe_groupnorm_finalised_in_the_consumer_kernels_is_bit_identical():
    """ym_moe_ffn_gn / ym_moe_combine_tc_gn (GroupNorm-1 / -2 finalised inside pass 2 / the combine from the partial sums) against the
    three-launch form with ym_gn_finalize_tiles in between: the same bits, on both width pairs and a ragged map."""
    from yolo_master_b200.nn.modules import moe as moe_mod
    from yolo_master_b200.nn import modules as M
    for C, hw in ((64, (20, 20)), (128, (13, 9))):
        torch.manual_seed(C)
        m = M.OptimizedMOEImproved(C, C, num_experts=4, top_k=2).to(DEV).eval()
        for prm in m.parameters():
            prm.data.normal_(0, 0.2)
        x = torch.randn((3, C, *hw), device=DEV).half().contiguous(memory_format=torch.channels_last)
        outs = {}
        for fold in (True, False):
            prev = moe_mod.MOE_GN_FOLD, moe_mod.MOE_ROUTER_FOLD
            moe_mod.MOE_GN_FOLD = moe_mod.MOE_ROUTER_FOLD = fold
            try:
                with torch.no_grad():
                    outs[fold] = m(x).clone()
                    snap = m.last_routing_snapshot
                    outs[(fold, "route")] = (snap["topk_indices"].clone(), snap["topk_weights"].clone(), snap["router_probs"].clone())
            finally:
                moe_mod.MOE_GN_FOLD, moe_mod.MOE_ROUTER_FOLD = prev
        assert torch.equal(outs[True], outs[False]), f"C={C}: folded GroupNorm / router finalisation must not change a bit"
        for a, b in zip(outs[(True, "route")], outs[(False, "route")]):
            assert torch.equal(a, b), "routing table published by the statistics pass must equal ym_router_topk's"
