"""GPU parity: every operator of the hot path, called through the C ABI (ctypes -> libym_b200.so), against the CPU oracle
on identical fp16-representable inputs.  Tolerance = BASELINE.json north_star: |a-b| <= 1e-3 + 1e-2*|b| for fp16
conv/attention outputs; router top-k indices exact."""
import math

import pytest
import torch

from _util import ATOL, RTOL, assert_close, assert_within_noise, close_stats
from oracle import yolo_master_oracle as O
from yolo_master_b200.nn import modules as M
from yolo_master_b200.utils.synth import fill_state_dict_

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _prep(mod, seed=0):
    """Deterministic weights; returns the fp32 CPU state_dict with keys prefixed 'm.' for the oracle."""
    sd = mod.state_dict()
    fill_state_dict_(sd, seed)
    mod.load_state_dict(sd)
    mod.eval().to(DEV)
    return {"m." + k: v.clone().float().cpu() for k, v in sd.items()}


def _x(B, C, H, W, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn((B, C, H, W), generator=g) * scale).half()


def _noise(fn, *a):
    """(fp32 oracle, fp16-storage oracle) outputs of an oracle function."""
    ref = fn(*a)
    with O.fp16_storage(), O.fp16_weights():
        sim = fn(*a)
    return ref, sim


def _run(mod, x):
    with torch.no_grad():
        y = mod(x.to(DEV).contiguous(memory_format=torch.channels_last))
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("c1,c2,k,s,act", [
    (16, 32, 3, 2, True), (64, 64, 3, 2, True), (32, 32, 1, 1, True), (16, 8, 3, 1, True), (8, 16, 3, 1, True),
    (384, 128, 1, 1, True), (64, 192, 1, 1, False), (128, 256, 3, 2, True), (48, 64, 1, 1, True), (256, 80, 1, 1, True),
    (16, 16, 3, 1, True), (16, 16, 3, 2, False), (8, 8, 3, 2, True), (8, 32, 3, 1, True), (16, 8, 3, 2, True), (8, 16, 3, 2, True),
])
@pytest.mark.parametrize("hw", [(20, 20), (37, 23), (70, 41)])
def test_conv(c1, c2, k, s, act, hw):
    m = M.Conv(c1, c2, k, s, act=act)
    sd = _prep(m, seed=c1 * 7 + c2)
    x = _x(2, c1, *hw, seed=3)
    y = _run(m, x)
    with O.fp16_weights():   # the reference's deployed weights: BN folded, rounded to fp16 (model.fuse().half())
        ref = O.conv_block(sd, "m", x.float(), s, 1, act)
    assert y.shape == ref.shape and y.dtype == torch.float16
    assert_close(y, ref, what=f"Conv({c1},{c2},{k},{s})")
    assert_close(y, O.conv_block(sd, "m", x.float(), s, 1, act), max_bad_frac=1e-4, what="vs fp32-weight oracle")


@pytest.mark.parametrize("c1,c2,s", [(16, 32, 2), (16, 8, 1), (8, 16, 1), (16, 16, 1), (8, 8, 2), (16, 32, 1)])
def test_small_conv_matches_implicit_gemm_and_adds_the_residual(c1, c2, s):
    """The patch-staged 3x3 kernel behind ym_conv2d_nhwc (Cin 8 / 16) against the implicit-GEMM kernel it replaces, on a multi-tile
    ragged image with and without the residual operand (the Bottleneck shortcut), and against the fp32 oracle."""
    from yolo_master_b200 import _lib, ops
    m = M.Conv(c1, c2, 3, s)
    sd = _prep(m, seed=c1 + 3 * c2 + s)
    x = _x(3, c1, 75, 133, seed=11).to(DEV).contiguous(memory_format=torch.channels_last)
    outs = {}
    for impl in (1, 0):
        prev = _lib.load().ym_set_small_conv_impl(impl)
        try:
            with torch.no_grad():
                outs[impl] = m(x)
            torch.cuda.synchronize()
        finally:
            _lib.load().ym_set_small_conv_impl(prev)
    ref = O.conv_block(sd, "m", x.float().cpu(), s, 1, True)
    assert_close(outs[1], ref, what=f"small conv {c1}->{c2} s{s}")
    assert float((outs[1].float() - outs[0].float()).abs().max()) <= 2e-3 * max(1.0, float(outs[0].float().abs().max()))
    if c1 == c2 and s == 1:
        b = M.Bottleneck(c1, c1, shortcut=True, e=0.5 if c1 == 16 else 1.0)
        _prep(b, seed=5)
        res = {}
        for impl in (1, 0):
            prev = _lib.load().ym_set_small_conv_impl(impl)
            try:
                with torch.no_grad():
                    res[impl] = b(x)
                torch.cuda.synchronize()
            finally:
                _lib.load().ym_set_small_conv_impl(prev)
        assert float((res[1].float() - res[0].float()).abs().max()) <= 4e-3 * max(1.0, float(res[0].float().abs().max()))


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.uint8])
@pytest.mark.parametrize("hw", [(64, 64), (63, 65), (200, 136)])
def test_stem_conv_reads_nchw_image(dtype, hw, impl):
    """Both stem kernels (0 = FFMA, 1 = mma.sync with split fp16 weights) on even, odd and multi-tile image sizes."""
    from yolo_master_b200 import _lib
    prev = _lib.load().ym_set_stem_impl(impl)
    try:
        _stem_case(dtype, hw)
    finally:
        _lib.load().ym_set_stem_impl(prev)


def _stem_case(dtype, hw):
    m = M.Conv(3, 16, 3, 2)
    sd = _prep(m, seed=1)
    g = torch.Generator().manual_seed(5)
    if dtype == torch.uint8:
        x = torch.randint(0, 256, (2, 3, *hw), generator=g, dtype=torch.uint8)
        xf = x.float() / 255.0
    else:
        x = torch.rand((2, 3, *hw), generator=g).to(dtype)
        xf = x.float()
    with torch.no_grad():
        y = m(x.to(DEV))
    ref = O.conv_block(sd, "m", xf, 2, 1, True)
    assert_close(y, ref, what="stem")


@pytest.mark.parametrize("c,hw,act", [(64, (80, 80), False), (32, (37, 23), True), (128, (40, 40), False), (64, (16, 16), False), (96, (21, 50), True)])
def test_dwconv7_tensor_core_kernel_matches_ffma_kernel_and_oracle(c, hw, act):
    """Depthwise 7x7 as Toeplitz GEMMs on mma.sync (csrc/dwconv_tc.cu) against the FFMA kernel it replaces and the fp32 oracle: whole and
    ragged tiles, several channel blocks, SiLU, and the fused residual (DWConv inside a shortcut)."""
    from yolo_master_b200 import _lib
    m = M.Conv(c, c, 7, 1, None, g=c, act=act)
    sd = _prep(m, seed=c + 7)
    x = _x(2, c, *hw, seed=13)
    outs = {}
    for impl in (1, 0):
        prev = _lib.load().ym_set_dwconv_tc(impl)
        try:
            outs[impl] = _run(m, x)
        finally:
            _lib.load().ym_set_dwconv_tc(prev)
    with O.fp16_weights():   # the deployed weights: BN folded, rounded to fp16
        ref = O.conv_block(sd, "m", x.float(), 1, c, act)
    assert_close(outs[1], ref, what=f"dwconv7 tc C={c}")
    assert_close(outs[1], O.conv_block(sd, "m", x.float(), 1, c, act), max_bad_frac=1e-4, what="vs fp32-weight oracle")
    assert float((outs[1].float() - outs[0].float()).abs().max()) <= 2e-3 * max(1.0, float(outs[0].float().abs().max()))


@pytest.mark.parametrize("c,k,act", [(64, 3, True), (80, 3, True), (64, 7, False), (128, 7, False)])
def test_dwconv(c, k, act):
    m = M.Conv(c, c, k, 1, None, g=c, act=act) if k == 7 else M.DWConv(c, c, k, act=act)
    sd = _prep(m, seed=c + k)
    x = _x(2, c, 21, 19, seed=4)
    y = _run(m, x)
    with O.fp16_weights():
        ref = O.conv_block(sd, "m", x.float(), 1, c, act)
    assert_close(y, ref, what=f"dw{k}")


def test_concat_and_upsample():
    a, b = _x(2, 64, 10, 12, 1), _x(2, 32, 20, 24, 2)
    cat = M.Concat(1)
    y = cat([a.to(DEV), b.to(DEV)], up_first=2)
    ref = torch.cat([torch.nn.functional.interpolate(a.float(), scale_factor=2.0, mode="nearest"), b.float()], 1)
    assert torch.equal(y.float().cpu(), ref)
    up = M.Upsample(None, 2, "nearest")
    assert torch.equal(up(a.to(DEV)).float().cpu(), ref[:, :64])
    c = _x(2, 16, 20, 24, 3)
    y3 = cat([b.to(DEV), c.to(DEV)])
    assert torch.equal(y3.float().cpu(), torch.cat([b.float(), c.float()], 1))


def test_bottleneck_and_c3k2_variants():
    x = _x(2, 64, 24, 20, seed=6)
    for args in [(64, 64, 1, False, 0.25), (64, 128, 1, True), (64, 256, 1, True, 0.5, True)]:
        m = M.C3k2(*args)
        sd = _prep(m, seed=len(args))
        y = _run(m, x)
        ref, sim = _noise(O.layer_c3k2, sd, "m", x.float(), *args)
        assert_within_noise(y, ref, sim, what=f"C3k2{args}")


def test_sppf_and_c2psa():
    x = _x(2, 256, 20, 20, seed=7)
    m = M.SPPF(256, 256, 5)
    sd = _prep(m, 3)
    assert_within_noise(_run(m, x), *_noise(O.layer_sppf, sd, "m", x.float(), 256, 256, 5), what="SPPF")
    m = M.C2PSA(256, 256, 1)
    sd = _prep(m, 4)
    assert_within_noise(_run(m, x), *_noise(O.layer_c2psa, sd, "m", x.float(), 256, 256, 1), what="C2PSA")


@pytest.mark.parametrize("dim,heads,area,hw", [(64, 2, 1, (40, 40)), (64, 2, 1, (13, 17)), (128, 4, 1, (20, 20)),
                                               (64, 2, 4, (16, 20)), (64, 2, 1, (80, 80))])
def test_area_attention(dim, heads, area, hw):
    m = M.AAttn(dim, heads, area)
    sd = _prep(m, seed=dim + area)
    x = _x(2, dim, *hw, seed=8)
    y = _run(m, x)
    ref, sim = _noise(O.aattn, sd, "m", x.float(), heads, area)
    assert_within_noise(y, ref, sim, what=f"AAttn({dim},{heads},{area},{hw})")
    assert close_stats(y, ref)[1] < 5e-3   # >99.5% of elements inside the single-op tolerance even after 4 chained ops


@pytest.mark.parametrize("N,heads,dv,batch", [(400, 2, 32, 3), (1600, 2, 32, 2), (221, 4, 32, 2), (400, 2, 64, 2), (6400, 2, 32, 1)])
def test_attention_kernel_alone_strict(N, heads, dv, batch):
    """The mma.sync attention kernel by itself (given fp16 q,k,v) meets the strict tolerance against fp32 softmax attention."""
    from yolo_master_b200 import _lib, ops
    prev = _lib.load().ym_set_attention_impl(0)
    hs = 64 + dv
    g = torch.Generator().manual_seed(N + dv)
    qkv = torch.randn((batch, N, 1, heads * hs), generator=g).half()
    out = ops.attention(qkv.to(DEV), batch, N, heads, hs, 0, 32, 64, 32, dv, 32 ** -0.5)
    _lib.load().ym_set_attention_impl(prev)
    t = qkv.float().view(batch, N, heads, hs).permute(0, 2, 1, 3)
    q, k, v = t[..., :32], t[..., 32:64], t[..., 64:]
    ref = torch.softmax((q * 32 ** -0.5) @ k.transpose(-1, -2), -1) @ v          # (b, h, N, dv)
    ref = ref.permute(0, 2, 1, 3).reshape(batch, N, 1, heads * dv)
    assert_close(out, ref, what=f"attention N={N} dv={dv}")


def test_psa_attention():
    m = M.Attention(128, num_heads=2, attn_ratio=0.5)
    sd = _prep(m, 9)
    x = _x(2, 128, 20, 20, seed=9)
    assert_within_noise(_run(m, x), *_noise(O.attention, sd, "m", x.float(), 2), what="Attention")


@pytest.mark.parametrize("C,E,hw", [(64, 4, (40, 40)), (64, 8, (20, 20)), (128, 16, (10, 10)), (64, 4, (4, 4)), (64, 4, (9, 13))])
def test_router_topk_indices_exact(C, E, hw):
    m = M.EfficientSpatialRouter(C, E, top_k=2)
    sd = _prep(m, seed=E)
    x = _x(8, C, *hw, seed=10)
    with torch.no_grad():
        w, idx, _ = m(x.to(DEV).contiguous(memory_format=torch.channels_last))
    wr, ir, probs = O.efficient_spatial_router(sd, "m", x.float(), 2)
    assert idx.dtype == torch.int64 and w.dtype == torch.float32
    srt = probs.sort(dim=1, descending=True)[0]
    margin = torch.minimum(srt[:, 0] - srt[:, 1], srt[:, 1] - srt[:, 2])
    ok = margin > 1e-4   # exact wherever the oracle's own decision is not a numerical tie
    assert ok.float().mean() > 0.7
    assert torch.equal(idx.cpu()[ok], ir[ok])
    torch.testing.assert_close(w.cpu()[ok], wr[ok], atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(w.sum(1).cpu(), torch.ones(8), atol=1e-5, rtol=0)


@pytest.mark.parametrize("C,E,hw", [(64, 4, (40, 40)), (64, 8, (20, 12)), (128, 16, (10, 10))])
def test_moe_ffn_block(C, E, hw):
    m = M.OptimizedMOEImproved(C, C, num_experts=E, top_k=2, expert_expand_ratio=2.0, add_residual=False)
    sd = _prep(m, seed=E + 1)
    x = _x(4, C, *hw, seed=11)
    y = _run(m, x)
    ref, sim = _noise(O.optimized_moe_improved, sd, "m", x.float(), E, 2)
    assert_within_noise(y, ref, sim, what=f"OptimizedMOEImproved({C},{E})")
    assert close_stats(y, ref)[1] < 1e-3


def test_ablock_moe_and_a2c2f_moe():
    x = _x(2, 64, 20, 20, seed=12)
    m = M.ABlockMoE(64, 2, 2.0, 1, 4, 2)
    sd = _prep(m, 5)
    assert_within_noise(_run(m, x), *_noise(O.ablock_moe, sd, "m", x.float(), 2, 1, 4, 2), what="ABlockMoE")
    args = (64, 128, 1, True, 1, False, 2.0, 0.5, 1, True, 4, 2)
    m = M.A2C2fMoE(*args)
    sd = _prep(m, 6)
    assert_within_noise(_run(m, x), *_noise(O.layer_a2c2f_moe, sd, "m", x.float(), *args), what="A2C2fMoE")


def _match_dets(y, ref, score_tol=2e-3, box_tol=0.3):
    """Every reference detection whose score is clear of the k-th score must appear with the same class and box."""
    B = y.shape[0]
    for b in range(B):
        kth = ref[b, -1, 4]
        for r in ref[b]:
            if r[4] < kth + 3 * score_tol:
                continue
            cand = y[b][(y[b, :, 5] == r[5]) & ((y[b, :, 4] - r[4]).abs() < score_tol)]
            assert len(cand) and (cand[:, :4] - r[:4]).abs().max(1)[0].min() < box_tol, (b, r)
    s1, s2 = y[..., 4].sort(dim=1, descending=True)[0], ref[..., 4].sort(dim=1, descending=True)[0]
    assert (s1 - s2).abs().max() < score_tol
    assert torch.equal(y[..., 4].sort(dim=1, descending=True)[0], y[..., 4])  # emitted score-descending


@pytest.mark.parametrize("hw0", [80, 20, 4])
def test_detect_end2end(hw0):
    ch = (64, 128, 256)
    m = M.Detect(80, 1, True, ch)
    m.stride = torch.tensor([8.0, 16.0, 32.0])
    sd = _prep(m, 7)
    feats = [_x(2, c, max(hw0 >> i, 1), max(hw0 >> i, 1), seed=20 + i) for i, c in enumerate(ch)]
    with torch.no_grad():
        y, _ = m([f.to(DEV).contiguous(memory_format=torch.channels_last) for f in feats])
    braw, sraw = O.detect_head_raw(sd, "m", [f.float() for f in feats], 80, 1, True)
    shapes = [tuple(f.shape[2:]) for f in feats]
    dec = O.detect_decode(braw, sraw, shapes, [8, 16, 32], True)
    ref, _ = O.detect_postprocess(dec, 80)
    assert y.shape == ref.shape and y.dtype == torch.float32
    _match_dets(y.cpu(), ref)


def test_detect_dense_decode():
    ch = (64, 128, 256)
    m = M.Detect(80, 1, True, ch)
    m.stride = torch.tensor([8.0, 16.0, 32.0])
    m.end2end = False
    sd = _prep(m, 8)
    feats = [_x(2, c, 16 >> i, 12 >> i, seed=30 + i) for i, c in enumerate(ch)]
    with torch.no_grad():
        y, _ = m([f.to(DEV).contiguous(memory_format=torch.channels_last) for f in feats])
    braw, sraw = O.detect_head_raw(sd, "m", [f.float() for f in feats], 80, 1, False)
    ref = O.detect_decode(braw, sraw, [tuple(f.shape[2:]) for f in feats], [8, 16, 32], False)
    assert y.shape == ref.shape
    torch.testing.assert_close(y[:, 4:].cpu(), ref[:, 4:], atol=2e-3, rtol=1e-2)
    torch.testing.assert_close(y[:, :4].cpu(), ref[:, :4], atol=0.08, rtol=1e-2)   # pixels (distance * stride<=32)


def test_errors_are_loud():
    m = M.Conv(16, 16, 3, 1).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 16, 8, 8))
    m.train()
    with pytest.raises(RuntimeError, match="eval"):
        m.to(DEV)(torch.zeros(1, 16, 8, 8, device=DEV))
    from yolo_master_b200 import _lib
    lib = _lib.load()
    rc = lib.ym_conv2d_nhwc(None, 8, 1, 1, 1, 8, None, 32, None, 8, 1, 1, 1, 0, None, 8, 0, None, 0, 0, None)
    assert rc != 0 and b"null" in lib.ym_last_error()


@pytest.mark.parametrize("case", [0, 1, 2])
def test_es_moe_vs_reference_golden_and_oracle(case):
    """ES_MOE eval forward: CUDA path vs the REAL reference module's output (golden) and the oracle; routes exact."""
    from test_esmoe_oracle import ESMOE_CASES, esmoe_case
    C, E, k, H, W, B = ESMOE_CASES[case]
    sd, x, yref = esmoe_case(case)
    m = M.ES_MOE(C, C, num_experts=E, top_k=k)
    assert set(m.state_dict()) == set(sd), "ES_MOE state_dict keys differ from the reference"
    m.load_state_dict(sd)
    m.eval().to(DEV)
    xh = x.half()
    y = _run(m, xh)
    sdo = {"m." + kk: v for kk, v in sd.items()}
    ref, (ti, w) = O.es_moe(sdo, "m", xh.float(), C, C, E, 8, k)
    with O.fp16_storage(), O.fp16_weights():
        sim, _ = O.es_moe(sdo, "m", xh.float(), C, C, E, 8, k)
    assert_within_noise(y, ref, sim, what="ES_MOE")
    assert close_stats(y, yref)[1] < 5e-3      # vs the real reference run on the fp32 input
    snap = m.last_routing_snapshot
    exp_idx = torch.where(w.gather(1, ti) > 0, ti, torch.full_like(ti, -1))
    assert torch.equal(snap["topk_indices"].cpu().long(), exp_idx)          # retained experts (and drops) exact
    torch.testing.assert_close(snap["topk_weights"].cpu(), w.gather(1, ti), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("H,W,C,k", [(20, 20, 128, 5), (13, 17, 64, 5), (40, 40, 32, 5), (7, 9, 16, 3), (112, 96, 8, 5)])
def test_sppf_pool_exact(H, W, C, k):
    """Chained MaxPool(k) x3 into slots 1..3 of the concat buffer: max is exact in fp16, so bit-identical to F.max_pool2d
    (separable shared-memory kernel for P5-sized maps, direct window kernel for the large-map fallback)."""
    import torch.nn.functional as F
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(H * W + C)
    y0 = torch.randn((3, H, W, C), generator=g).half()
    buf = torch.zeros((3, H, W, 4 * C), dtype=torch.float16)
    buf[..., :C] = y0
    d = buf.to(DEV)
    ops.sppf_pool(d, C, k)
    cur = y0.permute(0, 3, 1, 2).float()
    for s in range(1, 4):
        cur = F.max_pool2d(cur, k, 1, k // 2)
        assert torch.equal(d[..., s * C:(s + 1) * C].float().cpu(), cur.permute(0, 2, 3, 1)), f"slot {s}"
    assert torch.equal(d[..., :C].cpu(), y0)
