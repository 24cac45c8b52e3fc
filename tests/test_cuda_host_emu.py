"""The extern "C" entry points written after round 1's GPU budget, exercised END TO END without a GPU: the real .cu translation units
(preproc.cu, gated.cu, nms_large.cu, mix.cu, postproc.cu) are compiled with g++ against tests/native/cuda_host_emu.h - a CUDA execution model on
host threads (one OS thread per CUDA thread of a block, std::barrier for __syncthreads, thread_local blockIdx / threadIdx) - and
called through the real `yolo_master_b200.ops` wrappers with CPU tensors.  Unlike the phase / per-element harnesses this covers
the argument checks, the launch geometry, the shared-memory sizing, the grid-stride loops and the vectorised stores of the kernels
themselves, and the ctypes marshalling of `ops.py`.  What it cannot cover: timing, sm_100a code generation, fast-math rounding."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _util import GOLD, ROOT
from oracle import letterbox_oracle as L
from oracle import nms_oracle as N
from oracle import postproc_oracle as PP
from oracle import yolo_master_oracle as O
from yolo_master_b200 import _lib, ops

CSRC = os.path.join(ROOT, "yolo-master_b200", "csrc")
UNITS = ["preproc.cu", "gated.cu", "nms_large.cu", "mix.cu", "postproc.cu"]
SYMBOLS = ["ym_letterbox_u8", "ym_scale_boxes", "ym_scale_coords", "ym_kpts_decode", "ym_obb_finish", "ym_gate_router", "ym_gate_router_scratch_floats",
           "ym_zero_cost_router", "ym_zero_cost_router_scratch_floats", "ym_pixel_router", "ym_latent_router", "ym_fc_gate", "ym_classify_head", "ym_gated_select", "ym_gated_select_scratch_floats",
           "ym_ctx_mean3", "ym_gap_nhwc", "ym_nms_batched_large", "ym_nms_large_scratch_bytes", "ym_ew_nhwc", "ym_last_error",
           "ym_dwconv3_routed_nhwc", "ym_route_affine", "ym_process_mask", "ym_process_mask_scratch_bytes", "ym_nms_rotated", "ym_nms_rotated_scratch_bytes"]


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ym_emu") / "libym_emu.so")
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-DYM_HOST_EMU", "-I", os.path.join(ROOT, "tests", "native"),
           "-I", CSRC, "-x", "c++", *[os.path.join(CSRC, u) for u in UNITS], os.path.join(ROOT, "tests", "native", "emu_api.cpp"), "-o", so]
    subprocess.run(cmd, check=True)
    lib = C.CDLL(so)
    for name in SYMBOLS:
        res, args = _lib.SIGNATURES[name]
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


@pytest.fixture()
def emu(emu_lib, monkeypatch):
    monkeypatch.setattr(ops, "lib", lambda: emu_lib)
    monkeypatch.setattr(_lib, "load", lambda: emu_lib)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "REQUIRE_CUDA", False)
    return emu_lib


def test_argument_checks_report_through_ym_last_error(emu):
    with pytest.raises(RuntimeError, match="multiples of 8"):
        ops.ew(ops.EW_SIGMOID, a=torch.zeros((1, 2, 2, 12), dtype=torch.float16))
    with pytest.raises(RuntimeError, match="does not fit"):
        ops.letterbox(torch.zeros((1, 4, 4, 3), dtype=torch.uint8), torch.zeros((8, 2), dtype=torch.int32), torch.zeros((8, 2), dtype=torch.int32),
                      False, 8, 8, 2, 0, 8, 8)


def test_letterbox_kernel_geometry(emu):
    """Vector (W % 4 == 0) and scalar store paths, CHW / HWC, the three output types, B > 1, the 2x area path - whole kernels."""
    from yolo_master_b200.data.augment import LetterBox
    rng = np.random.default_rng(3)
    for (h, w, new_shape) in ((60, 80, (96, 96)), (33, 7, (70, 50)), (250, 250, (125, 125)), (90, 160, (101, 203)), (64, 64, (64, 64))):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = L.preprocess_frame(img, new_shape)
        lb = LetterBox(new_shape)
        src = torch.from_numpy(img)[None]
        assert np.array_equal(lb.apply_batch(src, swap_rb=True, chw=True)[0].numpy(), want), (h, w, new_shape)
        assert np.array_equal(lb.apply_batch(src)[0].numpy(), L.letterbox_frame(img, new_shape)), (h, w, new_shape)
        u8 = torch.from_numpy(want)
        assert torch.equal(lb.apply_batch(src, swap_rb=True, chw=True, dtype=torch.float16)[0], u8.half() / 255)
        assert torch.equal(lb.apply_batch(src, swap_rb=True, chw=True, dtype=torch.float32)[0], u8.float() / 255)
    frames = np.stack([rng.integers(0, 256, (45, 60, 3), dtype=np.uint8) for _ in range(3)])
    out = LetterBox((64, 64)).apply_batch(torch.from_numpy(frames), swap_rb=True, chw=True).numpy()
    for i in range(3):
        assert np.array_equal(out[i], L.preprocess_frame(frames[i], (64, 64))), i
    case = torch.load(os.path.join(GOLD, "letterbox.golden.pt"))["cases"][7]             # 100 x 37 -> 640 x 640, the reference's CRC
    img = np.random.default_rng(case["seed"]).integers(0, 256, (case["h"], case["w"], 3), dtype=np.uint8)
    full = LetterBox((640, 640)).apply_batch(torch.from_numpy(img)[None], swap_rb=True, chw=True)[0].numpy()
    assert zlib.crc32(full.tobytes()) == case["crc"]


def test_scale_boxes_kpts_obb_kernels(emu):
    from yolo_master_b200.utils import ops as box_ops
    g = torch.Generator().manual_seed(4)
    shapes = [(480, 640), (1080, 1920), (100, 37)] * 50                                   # 150 images: two parameter blocks
    b = torch.rand((len(shapes), 7, 6), generator=g) * 700 - 30
    dev = b.clone()
    box_ops.scale_boxes_batch((640, 640), dev, shapes)
    for i in (0, 1, 2, 127, 128, 149):
        assert np.array_equal(dev[i, :, :4].numpy(), L.scale_boxes((640, 640), b[i, :, :4].numpy(), shapes[i])), i
    nk, B, lv, strides = 51, 2, [(8, 6), (4, 3), (2, 2)], [8.0, 16.0, 32.0]
    levels = [torch.randn((B, h, w, nk), generator=g) for h, w in lv]
    y = ops.kpts_decode(levels, strides, 3)
    raw = torch.cat([t.reshape(B, -1, nk).transpose(1, 2) for t in levels], 2)
    anchors, st = O.make_anchors(lv, strides)
    want = raw.clone()
    want[:, 2::3] = want[:, 2::3].sigmoid()
    want[:, 0::3] = (raw[:, 0::3] * 2.0 + (anchors.t()[0] - 0.5)) * st.t()
    want[:, 1::3] = (raw[:, 1::3] * 2.0 + (anchors.t()[1] - 0.5)) * st.t()
    torch.testing.assert_close(y, want, atol=1e-6, rtol=1e-6)
    assert torch.equal(ops.kpts_decode(levels, strides, 1), raw)                          # ndim 1: the plain gather Segment uses
    A, nc = raw.shape[2], 3
    yin = torch.rand((B, 4 + nc, A), generator=g) * 50
    ang = [torch.randn((B, h, w, 1), generator=g) for h, w in lv]
    out = ops.obb_finish(yin, ang, strides, nc)
    assert out.shape == (B, 4 + nc + 1, A) and torch.equal(out[:, 2:4 + nc], yin[:, 2:])
    a = (torch.cat([t.reshape(B, -1, 1).transpose(1, 2) for t in ang], 2).sigmoid() - 0.25) * 3.14159265358979
    torch.testing.assert_close(out[:, -1:], a, atol=1e-6, rtol=1e-6)


def test_gated_kernels(emu):
    """Router (three kernels, dynamic shared memory), gates, select, context mean, classifier - against the oracle functions."""
    from yolo_master_b200.nn.modules.gated import OptimalHybridGateMoE, UltimateOptimizedMoE, VisualEnhancedAdaptiveGateMoE
    from yolo_master_b200.utils.synth import fill_state_dict_
    g = torch.Generator().manual_seed(5)
    for cls, c, E, k, H, W in ((VisualEnhancedAdaptiveGateMoE, 128, 4, 2, 20, 24), (OptimalHybridGateMoE, 64, 16, 2, 9, 7)):
        m = cls(c, c, E, k, 0.5)
        sd = m.state_dict()
        fill_state_dict_(sd, 11)
        m.load_state_dict(sd)
        sdm = {"m." + k_: v.float() for k_, v in sd.items()}
        xd = torch.randn((3, c // 2, H, W), generator=g).half()
        idx, w, probs = ops.gate_router(xd.permute(0, 2, 3, 1).contiguous(), m.eval().get_pack()["router"], k)
        cx = torch.sigmoid(F.conv2d(xd.float().mean((2, 3), keepdim=True), sdm["m.complexity_estimator.1.weight"], sdm["m.complexity_estimator.1.bias"])).mean()
        rw, ri, rp = O.dual_stream_gate_router(sdm, "m.routing", xd.float(), k, 1.2)
        assert torch.equal(idx.long(), ri)
        torch.testing.assert_close(probs, rp, atol=2e-6, rtol=1e-4)
        torch.testing.assert_close(w, O.complexity_gate(rw, cx.clamp(0.3, 1.5)), atol=2e-6, rtol=1e-4)
    from yolo_master_b200.nn.modules.moe import UltraOptimizedMoE
    for E, hw in ((4, 24), (16, 7)):                     # pooled (24 > 8) and un-pooled router maps; per-thread partials per expert
        mu = UltraOptimizedMoE(64, 64, E, 2)
        sdu = mu.state_dict()
        fill_state_dict_(sdu, 9)
        mu.load_state_dict(sdu)
        xu = torch.randn((2, 64, hw, hw), generator=g).half()
        idx, w, probs = ops.pixel_router(xu.permute(0, 2, 3, 1).contiguous(), mu.eval().get_pack()["router"], 2, 0.01)
        rw, ri, rp = O.ultra_efficient_router({"m." + k_: v.float() for k_, v in sdu.items()}, "m.routing", xu.float(), 2)
        assert torch.equal(idx.long(), ri)
        torch.testing.assert_close(probs, rp, atol=2e-6, rtol=1e-4)
        torch.testing.assert_close(w, torch.where(rw > 0.01, rw, torch.zeros_like(rw)), atol=2e-6, rtol=1e-4)
    m = UltimateOptimizedMoE(64, 64, 4, 2, 0.5)
    sd = m.state_dict()
    fill_state_dict_(sd, 3)
    m.load_state_dict(sd)
    pk = m.eval().get_pack()
    xd = torch.randn((2, 32, 6, 5), generator=g).half()
    idx, w, _ = ops.zero_cost_router(xd.permute(0, 2, 3, 1).contiguous(), pk["fc"], 2.0, pk["cx_w"], pk["cx_b"], 2)
    rw, ri, _ = O.zero_cost_router({"m." + k_: v.float() for k_, v in sd.items()}, "m.routing", xd.float(), 2, 2.0)
    assert torch.equal(idx.long(), ri) and bool((w > 0).all())
    # select + context mean + classifier + fc gate with a strided input view
    B, H, W, E, oc, G = 2, 7, 5, 4, 16, 8
    buf = torch.randn((B, H, W, E * oc + 8), generator=g).half()
    idx = torch.tensor([[2, 0], [1, 3]], dtype=torch.int32)
    ww = torch.tensor([[0.7, 0.3], [1.0, 0.0]])
    gamma, beta = torch.randn((E, oc), generator=g), torch.randn((E, oc), generator=g)
    got = ops.gated_select(buf[..., :E * oc], idx, ww, gamma, beta, E, oc, G).float()
    f5 = buf[..., :E * oc].float().permute(0, 3, 1, 2).reshape(B, E, oc, H, W)
    sel = torch.gather(f5, 1, idx.long().view(B, 2, 1, 1, 1).expand(B, 2, oc, H, W))
    nrm = F.group_norm(sel.reshape(B * 2, oc, H, W), G).view(B, 2, oc, H, W) * gamma[idx.long()].view(B, 2, oc, 1, 1) + beta[idx.long()].view(B, 2, oc, 1, 1)
    torch.testing.assert_close(got, (F.silu(nrm) * ww.view(B, 2, 1, 1, 1)).sum(1).permute(0, 2, 3, 1), atol=4e-3, rtol=2e-3)
    a = torch.randn((2, 5, 7, 16), generator=g).half()
    b, c_ = torch.randn((2, 2, 3, 16), generator=g).half(), torch.randn((2, 1, 1, 16), generator=g).half()
    up = lambda t: F.interpolate(t.float().permute(0, 3, 1, 2), size=(5, 7), mode="nearest").permute(0, 2, 3, 1)
    torch.testing.assert_close(ops.ctx_mean3(a, b, c_).float(), torch.stack([a.float(), up(b), up(c_)]).mean(0), atol=2e-3, rtol=1e-3)
    v = torch.randn((3, 1, 1, 96), generator=g).half()
    wl, bl = torch.randn((10, 96), generator=g) * 0.1, torch.randn((10,), generator=g)
    probs, logits = ops.classify_head(v, wl, bl)
    torch.testing.assert_close(logits, v.view(3, 96).float() @ wl.t() + bl, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(probs, torch.softmax(logits, 1), atol=1e-7, rtol=1e-5)
    w1, w2, b2 = torch.randn((8, 96), generator=g) * 0.1, torch.randn((12, 8), generator=g), torch.randn((12,), generator=g)
    got = ops.fc_gate(v, w1, w2, b2, scale=0.4, offset=0.5)
    torch.testing.assert_close(got, 0.5 + 0.4 * torch.sigmoid(F.linear(F.silu(F.linear(v.view(3, 96).float(), w1)), w2, b2)), atol=1e-6, rtol=1e-5)


def test_slab_statistics_merge(emu):
    """The slab-parallel statistics (gate_r0 + merge, select_s0 + merge) with several slabs per image, a ragged last slab and a mean far
    from zero (where a naive sum-of-squares would cancel): the zero-cost router sees mean | std directly, the select sees GroupNorm."""
    from yolo_master_b200.nn.modules.gated import UltimateOptimizedMoE
    from yolo_master_b200.utils.synth import fill_state_dict_
    g = torch.Generator().manual_seed(8)
    m = UltimateOptimizedMoE(64, 64, 4, 2, 0.5)
    sd = m.state_dict()
    fill_state_dict_(sd, 3)
    m.load_state_dict(sd)
    pk = m.eval().get_pack()
    xd = (torch.randn((2, 32, 45, 11), generator=g) * 0.25 + 6.0).half()          # 45 rows -> 23 slabs of 2 rows, the last of 1
    idx, w, probs = ops.zero_cost_router(xd.permute(0, 2, 3, 1).contiguous(), pk["fc"], 2.0, pk["cx_w"], pk["cx_b"], 2)
    rw, ri, rp = O.zero_cost_router({"m." + k_: v.float() for k_, v in sd.items()}, "m.routing", xd.float(), 2, 2.0)
    assert torch.equal(idx.long(), ri)
    torch.testing.assert_close(probs, rp, atol=5e-6, rtol=2e-4)
    from yolo_master_b200.nn.modules.gated import VisualEnhancedAdaptiveGateMoE
    mv = VisualEnhancedAdaptiveGateMoE(128, 128, 8, 2, 0.5)
    sdv = mv.state_dict()
    fill_state_dict_(sdv, 11)
    mv.load_state_dict(sdv)
    sdm = {"m." + k_: v.float() for k_, v in sdv.items()}
    xv = (torch.randn((2, 64, 40, 44), generator=g) + 1.5).half()                 # pooled 10 x 11: depthwise slabs 7 x 16 (ragged), 1x1 slabs 4 x 32
    idx, w, probs = ops.gate_router(xv.permute(0, 2, 3, 1).contiguous(), mv.eval().get_pack()["router"], 2)
    rw, ri, rp = O.dual_stream_gate_router(sdm, "m.routing", xv.float(), 2, 1.2)
    assert torch.equal(idx.long(), ri)
    torch.testing.assert_close(probs, rp, atol=5e-6, rtol=2e-4)
    B, H, W, E, oc, G = 2, 27, 23, 4, 32, 8                                        # 621 pixels -> 3 slabs of 207
    buf = (torch.randn((B, H, W, E * oc), generator=g) * 0.5 + 4.0).half()
    idx = torch.tensor([[2, 0], [1, 3]], dtype=torch.int32)
    ww = torch.tensor([[0.7, 0.3], [0.4, 0.6]])
    gamma, beta = torch.randn((E, oc), generator=g), torch.randn((E, oc), generator=g)
    got = ops.gated_select(buf, idx, ww, gamma, beta, E, oc, G).float()
    f5 = buf.double().permute(0, 3, 1, 2).reshape(B, E, oc, H, W)
    sel = torch.gather(f5, 1, idx.long().view(B, 2, 1, 1, 1).expand(B, 2, oc, H, W))
    nrm = F.group_norm(sel.reshape(B * 2, oc, H, W), G).view(B, 2, oc, H, W) * gamma[idx.long()].double().view(B, 2, oc, 1, 1) + beta[idx.long()].double().view(B, 2, oc, 1, 1)
    torch.testing.assert_close(got, (F.silu(nrm) * ww.double().view(B, 2, 1, 1, 1)).sum(1).permute(0, 2, 3, 1).float(), atol=4e-3, rtol=2e-3)


def test_elementwise_ops_including_the_verified_ones(emu):
    """mix.cu through the same launch macro as on the GPU: the three new ops and, as a sanity check of the emulation itself, ops
    whose GPU behaviour is already established."""
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn((2, 5, 7, 16), generator=g).half(), torch.randn((2, 5, 7, 16), generator=g).half()
    t, chan = torch.tensor([0.37]), torch.randn((16,), generator=g)
    torch.testing.assert_close(ops.ew(ops.EW_SIGMOID, a=a).float(), torch.sigmoid(a.float()), atol=1e-3, rtol=1e-3)
    torch.testing.assert_close(ops.ew(ops.EW_MUL_GATE, a=a, b=b, p0=t).float(), a.float() * (1 + 0.37 * b.float()), atol=2e-3, rtol=2e-3)
    torch.testing.assert_close(ops.ew(ops.EW_MUL, a=a, b=b).float(), a.float() * b.float(), atol=2e-3, rtol=2e-3)
    torch.testing.assert_close(ops.ew(ops.EW_SCALE_RES, a=a, b=b, p0=chan).float(), a.float() + chan * b.float(), atol=2e-3, rtol=2e-3)
    torch.testing.assert_close(ops.ew(ops.EW_GLU, a=a, b=b).float(), torch.sigmoid(a.float()) * b.float(), atol=2e-3, rtol=2e-3)
    sc, sh = torch.rand((2, 16), generator=g), torch.randn((2, 16), generator=g)
    want = F.silu(a.float() * sc.view(2, 1, 1, 16) + sh.view(2, 1, 1, 16)) + b.float()
    torch.testing.assert_close(ops.ew(ops.EW_AFFINE, a=a, b=b, p0=sc, p1=sh, rows_per_img=35, act=True).float(), want, atol=4e-3, rtol=2e-3)
    wide = torch.randn((2, 5, 7, 32), generator=g).half()                                # channel-slice views in and out
    out = torch.zeros((2, 5, 7, 32), dtype=torch.float16)
    ops.ew(ops.EW_MUL, a=wide[..., 16:], b=b, out=out[..., :16])
    torch.testing.assert_close(out[..., :16].float(), wide[..., 16:].float() * b.float(), atol=2e-3, rtol=2e-3)
    assert bool((out[..., 16:] == 0).all())


def test_large_nms_kernel(emu):
    """The whole ym_nms_batched_large call (best-class kernel + 1024-thread per-image kernel with its static shared block)."""
    g = torch.Generator().manual_seed(17)
    B, nc, A = 2, 3, 1500
    centres = torch.rand((B, 2, A // 20 + 1), generator=g) * 320
    cxy = centres.repeat_interleave(20, 2)[:, :, :A] + torch.randn((B, 2, A), generator=g) * 6
    pred = torch.cat([cxy, torch.exp(torch.randn((B, 2, A), generator=g) * 0.5 + 3.0), torch.rand((B, nc, A), generator=g) ** 4], 1).contiguous()
    out, cnt, idx = ops.nms_batched_large(pred, 0.05, 0.6, 50, 1000)
    ro, rk = N.non_max_suppression(pred, 0.05, 0.6, max_det=50, max_nms=1000)
    for b in range(B):
        n = int(cnt[b])
        assert n == len(rk[b]) and torch.equal(idx[b, :n].long(), rk[b]) and torch.equal(out[b, :n], ro[b])


def test_latent_router_kernel(emu):
    """ym_latent_router (LayerNorm + two-layer MLP + head + softmax in one CTA per image, dynamic shared memory sized from C / hidden /
    E) against the oracle's LatentRouter, for 2 and 3 tokens, default and explicit hidden width, T < 0.1 clamped."""
    from yolo_master_b200.nn.modules.latent import LatentMixture
    from yolo_master_b200.utils.synth import fill_state_dict_
    g = torch.Generator().manual_seed(12)
    for in_ch, oc, hid, temp in (((64, 32), 64, None, 1.0), ((128, 64, 128), 128, 48, 0.5), ((256, 256), 256, None, 0.05)):
        m = LatentMixture(list(in_ch), oc, 4, 0.25, hid, temp)
        sd = m.state_dict()
        fill_state_dict_({k: v for k, v in sd.items() if torch.is_tensor(v)}, 7)
        m.load_state_dict(sd)
        sdm = {"m." + k: v.float() for k, v in sd.items() if torch.is_tensor(v)}
        tokens = [torch.randn((3, 1, 1, oc), generator=g).half() for _ in in_ch]
        probs, logits = ops.latent_router(tokens, m.eval().get_pack()["router"])
        rl, rp = O.latent_router(sdm, "m.router", torch.stack([t.view(3, oc) for t in tokens], 1), temp)
        torch.testing.assert_close(logits, rl, atol=2e-5, rtol=1e-4)
        torch.testing.assert_close(probs, rp, atol=2e-6, rtol=1e-4)


def test_global_average_pool_kernel(emu):
    """ym_gap_nhwc (one CTA per image x 64-channel slab, 8 octets x 32 pixel lanes, fixed-order reduction) against a float64 mean: ragged
    last slab, a channel-slice view (pitch > C), HW smaller than the lane count, and a pre-allocated output slice."""
    g = torch.Generator().manual_seed(3)
    for B, H, W, Cc, ld in ((2, 5, 7, 64, 64), (3, 20, 20, 72, 96), (1, 2, 3, 8, 8), (2, 9, 4, 200, 200)):
        buf = torch.randn((B, H, W, ld), generator=g).half()
        x = buf[..., :Cc]
        out = ops.gap(x)
        ref = x.double().mean((1, 2), keepdim=True)
        assert out.shape == (B, 1, 1, Cc)
        torch.testing.assert_close(out.double(), ref, atol=1e-3, rtol=1e-3)
        assert torch.equal(out, x.float().mean((1, 2), keepdim=True).half()) or (out.float() - ref.float()).abs().max() < 1e-3
    wide = torch.zeros((2, 1, 1, 96), dtype=torch.float16)
    x = torch.randn((2, 6, 6, 32), generator=g).half()
    ops.gap(x, out=wide[..., 64:96])
    torch.testing.assert_close(wide[..., 64:96].float(), x.float().mean((1, 2), keepdim=True), atol=1e-3, rtol=1e-3)
    assert not wide[..., :64].any()
    with pytest.raises(RuntimeError, match="multiples of 8"):
        ops.gap(torch.zeros((1, 2, 2, 12), dtype=torch.float16))


def _unpack_bits(case, key):
    shape = tuple(case[key + "_shape"])
    return torch.from_numpy(np.unpackbits(case[key + "_bits"].numpy())[:int(np.prod(shape))].reshape(shape))


def assert_masks_match(got, want, field, keep, what, tol=1e-4):
    """Bit-exact except where the thresholded fp32 field is within `tol` of zero (another summation order may land on the other
    side); the crop itself (keep) is integer-exact."""
    assert got.shape == want.shape and got.dtype == torch.uint8, what
    bad = got != want
    assert not (bad & ~keep).any(), (what, "mask pixels outside the crop")
    assert not (bad & (field.abs() > tol)).any(), (what, int(bad.sum()), float(field[bad].abs().max()) if bad.any() else 0.0)
    assert bad.float().mean().item() < 1e-3, (what, "too many marginal pixels")


def test_process_mask_kernels_match_reference_goldens(emu):
    """ym_process_mask (both kernels, vector and scalar stores, fp16 and fp32 prototypes) against the REFERENCE's ops.process_mask."""
    G = torch.load(os.path.join(GOLD, "postproc.golden.pt"))
    for i, c in enumerate(G["masks"]):
        n, nm = c["coef"].shape
        dets = torch.cat([c["boxes"], torch.zeros((n, 2)), c["coef"]], 1).contiguous()
        for key, up in (("up", True), ("native", False)):
            want = _unpack_bits(c, key)
            for protos in (c["protos"], c["protos"].half()):                      # the goldens' prototypes are fp16-representable
                got = ops.process_mask(protos, dets, c["shape"], upsample=up)
                if n == 0:
                    assert got.shape == want.shape
                    continue
                field = PP.mask_field(c["protos"], c["coef"], c["shape"], up)
                mh, mw = c["protos"].shape[1:]
                boxes = c["boxes"] if up else c["boxes"] * torch.tensor([[mw / c["shape"][1], mh / c["shape"][0]] * 2])
                assert_masks_match(got, want, field, PP.crop_keep(boxes, *field.shape[1:]), (i, key, str(protos.dtype)))
    with pytest.raises(RuntimeError, match="row pitch"):
        ops.process_mask(torch.zeros((32, 4, 4)), torch.zeros((1, 10)), (16, 16))
    with pytest.raises(RuntimeError, match="nm <= 64"):
        ops.process_mask(torch.zeros((65, 4, 4)), torch.zeros((1, 80)), (16, 16))


def test_process_mask_more_than_one_detection_group_and_ragged_width(emu):
    g = torch.Generator().manual_seed(4)
    protos = torch.randn((5, 9, 7), generator=g)
    n = 19                                                                        # three groups of 8 detections, the last ragged
    coef = torch.randn((n, 5), generator=g)
    boxes = torch.tensor([[3.0, 2.0, 20.0, 30.0]]).repeat(n, 1) + torch.rand((n, 4), generator=g) * 3
    dets = torch.cat([boxes, torch.zeros((n, 3)), coef, torch.zeros((n, 2))], 1).contiguous()   # coefficients at column 7, pitch 14
    for up, shape in ((True, (37, 27)), (True, (36, 28)), (False, (36, 28))):      # ow % 4 != 0 -> scalar stores
        got = ops.process_mask(protos, dets, shape, upsample=up, coef_col=7)
        want = PP.process_mask(protos, coef, boxes, shape, upsample=up)
        field = PP.mask_field(protos, coef, shape, up)
        bx = boxes if up else boxes * torch.tensor([[7 / shape[1], 9 / shape[0]] * 2])
        assert_masks_match(got, want, field, PP.crop_keep(bx, *field.shape[1:]), (up, shape))


def test_rotated_nms_kernels_match_reference_goldens(emu):
    """ym_nms_rotated (best class, sort, tiled fast-NMS sweep, ordered emit) against the REFERENCE's non_max_suppression(rotated=True)."""
    G = torch.load(os.path.join(GOLD, "postproc.golden.pt"))
    for i, c in enumerate(G["nms"]):
        _, _, margin = PP.non_max_suppression_rotated(c["pred"], c["conf"], c["iou"], c["max_det"], c["max_nms"])
        assert margin > 1e-5, (i, margin)                                        # every suppression decision is clear of fp32 noise
        out, cnt, idx = ops.nms_rotated(c["pred"], c["conf"], c["iou"], c["max_det"], c["max_nms"])
        for b in range(c["pred"].shape[0]):
            k = int(cnt[b])
            assert k == len(c["keep"][b]), (i, b, k, len(c["keep"][b]))
            assert torch.equal(idx[b, :k].long(), c["keep"][b]), (i, b)
            assert torch.equal(out[b, :k], c["out"][b]), (i, b)
            assert (idx[b, k:] == -1).all() and (out[b, k:] == 0).all()


def test_rotated_nms_ties_caps_and_empty(emu):
    nc, A = 2, 40
    pred = torch.zeros((1, 4 + nc + 1, A))
    pred[0, 0] = torch.arange(A) * 100.0                                          # far apart: nothing suppresses
    pred[0, 1] = 50.0
    pred[0, 2], pred[0, 3] = 8.0, 4.0
    pred[0, 4] = 0.5                                                              # all scores tie: anchor order
    pred[0, 5] = 0.1
    out, cnt, idx = ops.nms_rotated(pred, 0.25, 0.45, max_det=7, max_nms=25)
    assert int(cnt[0]) == 7 and idx[0].tolist() == list(range(7))
    out, cnt, idx = ops.nms_rotated(pred, 0.9, 0.45)
    assert int(cnt[0]) == 0 and (idx == -1).all()
    pred[0, 0] = 50.0                                                             # identical boxes, same class: only the first survives
    out, cnt, idx = ops.nms_rotated(pred, 0.25, 0.45)
    assert int(cnt[0]) == 1 and int(idx[0, 0]) == 0
    pred[0, 5, 1::2] = 0.9                                                        # odd anchors switch class (and score): two survivors
    out, cnt, idx = ops.nms_rotated(pred, 0.25, 0.45)
    assert int(cnt[0]) == 2 and idx[0, :2].tolist() == [1, 0] and out[0, 0, 5] == 1 and out[0, 1, 5] == 0
    want, keep, _ = PP.non_max_suppression_rotated(pred, 0.25, 0.45)
    assert torch.equal(out[0, :2], want[0]) and torch.equal(idx[0, :2].long(), keep[0])


class _TaskStub(torch.nn.Module):
    def __init__(self, out, nc):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.out, self.stride, self.names = out, torch.tensor([8.0, 16.0, 32.0]), {i: str(i) for i in range(nc)}

    def forward(self, x):
        return self.out


def test_obb_predictor_postprocess_end_to_end(emu):
    """OBBPredictor.postprocess: ym_nms_rotated -> (x, y, w, h, angle, conf, cls) rows with ym_scale_boxes(xywh) - against the
    reference's own NMS output (golden) followed by the oracle rescale, exactly."""
    from yolo_master_b200.engine import OBBPredictor
    c = torch.load(os.path.join(GOLD, "postproc.golden.pt"))["nms"][0]
    B = c["pred"].shape[0]
    pred = OBBPredictor(_TaskStub((c["pred"], {}), c["nc"]), imgsz=640, conf=c["conf"], iou=c["iou"], max_det=c["max_det"], device="cpu")
    frames = [np.zeros((480, 640, 3), np.uint8), np.zeros((360, 500, 3), np.uint8)][:B]
    res = pred.postprocess((c["pred"], {}), torch.zeros((B, 3, 640, 640)), frames)
    for b, r in enumerate(res):
        ref = c["out"][b]
        want = torch.cat([torch.from_numpy(L.scale_boxes((640, 640), ref[:, :4].numpy(), frames[b].shape, xywh=True)), ref[:, 6:7], ref[:, 4:6]], 1)
        assert r.boxes is None and r.masks is None and len(r) == len(ref)
        assert torch.equal(r.obb.data, want), b
        assert torch.equal(r.obb.xywhr, want[:, :5]) and torch.equal(r.obb.conf, ref[:, 4]) and torch.equal(r.obb.cls, ref[:, 5])
        assert r.obb.xyxyxyxy.shape == (len(ref), 4, 2) and torch.allclose(r.obb.xyxyxyxy.mean(1), want[:, :2], atol=1e-2)
        assert len(r[:3].obb) == 3 and r.cpu().obb.data.shape == want.shape


def test_segmentation_predictor_postprocess_end_to_end(emu, monkeypatch):
    """SegmentationPredictor.postprocess: NMS rows carrying the coefficients -> ym_process_mask (boxes / coefficients read in place
    from the (n, 6 + nm) rows) -> ym_scale_boxes -> mask-less rows dropped.  The shared-memory NMS kernel (nms.cu, warp intrinsics) is
    not part of the emulated units: its place is taken by the NMS oracle here; everything after it is the product path."""
    from yolo_master_b200.engine import SegmentationPredictor
    from yolo_master_b200.utils import nms as host_nms
    g = torch.Generator().manual_seed(21)
    B, nc, nm, A = 2, 5, 8, 300
    y = torch.zeros((B, 4 + nc + nm, A))
    y[:, 0:2] = torch.rand((B, 2, A), generator=g) * 140 + 10
    y[:, 2:4] = torch.rand((B, 2, A), generator=g) * 60 + 6
    y[:, 4:4 + nc] = torch.rand((B, nc, A), generator=g) ** 3
    y[:, 4 + nc:] = torch.randn((B, nm, A), generator=g)
    y[0, 4 + nc:, :] -= 100.0 * (torch.arange(A) % 3 == 0)                      # some detections with an all-negative field: dropped
    protos = torch.nn.functional.avg_pool2d(torch.randn((B, nm, 40, 40), generator=g), 3, 1, 1).abs().half()

    def oracle_nms(prediction, conf_thres, iou_thres, classes=None, agnostic=False, max_det=300, nc=0, **kw):
        outs, idxs = N.non_max_suppression(prediction[:, :4 + nc].float(), conf_thres, iou_thres, max_det)
        return [torch.cat([o, prediction[b, 4 + nc:, i].t().float()], 1) for b, (o, i) in enumerate(zip(outs, idxs))]

    monkeypatch.setattr(host_nms, "non_max_suppression", oracle_nms)
    pred = SegmentationPredictor(_TaskStub(((y, protos), {}), nc), imgsz=160, conf=0.3, iou=0.5, device="cpu")
    frames = [np.zeros((120, 160, 3), np.uint8), np.zeros((300, 200, 3), np.uint8)]
    res = pred.postprocess(((y, protos), {}), torch.zeros((B, 3, 160, 160)), frames)
    dropped = 0
    for b, r in enumerate(res):
        rows = oracle_nms(y, 0.3, 0.5, nc=nc)[b]
        masks = PP.process_mask(protos[b].float(), rows[:, 6:], rows[:, :4], (160, 160), upsample=True)
        keep = masks.amax((-2, -1)) > 0
        dropped += int((~keep).sum())
        want = rows[keep][:, :6].clone()
        want[:, :4] = torch.from_numpy(L.scale_boxes((160, 160), want[:, :4].numpy(), frames[b].shape))
        assert torch.equal(r.boxes.data, want), b
        field = PP.mask_field(protos[b].float(), rows[:, 6:], (160, 160), True)[keep]
        assert_masks_match(r.masks.data, masks[keep], field, PP.crop_keep(rows[keep][:, :4], 160, 160), b)
        assert r.masks.shape == (int(keep.sum()), 160, 160) and len(r) == int(keep.sum())
    assert dropped > 0
    with pytest.raises(NotImplementedError):
        SegmentationPredictor(_TaskStub(None, nc), retina_masks=True, device="cpu")


def test_routed_dilated_depthwise_and_route_affine_kernels(emu):
    """ym_dwconv3_routed_nhwc (taps and dilation chosen per image on the device, dilations 1 ... 8 as the 16-expert v0_14 block uses,
    on a sliced view with a pitch) and ym_route_affine against torch: F.conv2d(dilation=d) per image / gathered gamma, beta."""
    g = torch.Generator().manual_seed(8)
    B, H, W, C, E = 5, 11, 9, 24, 16
    buf = torch.randn((B, H, W, C + 8), generator=g).half()
    x = buf[..., 8:]                                                              # channel-sliced view: pitch 32, 24 channels
    w = (torch.randn((E, 9, C), generator=g) * 0.3).half()
    dil = torch.tensor([1 + e // 2 for e in range(E)], dtype=torch.int32)
    routes = torch.tensor([[0, 3], [15, 1], [6, 6], [9, 14], [2, 11]], dtype=torch.int32)   # (B, top_k): column views have stride 2
    for j in range(2):
        got = ops.dwconv3_routed(x, w, routes[:, j], dil)
        for b in range(B):
            e = int(routes[b, j])
            d = int(dil[e])
            want = F.conv2d(x[b:b + 1].float().permute(0, 3, 1, 2), w[e].float().t().reshape(C, 1, 3, 3), None, 1, d, d, C)[0].permute(1, 2, 0)
            assert torch.equal(got[b], want.half()) or float((got[b].float() - want).abs().max()) < 2e-3, (j, b)   # fp32 sums, one fp16 rounding
    sc, sh = torch.rand((B, C), generator=g) + 0.5, torch.randn((B, C), generator=g)
    gamma, beta = torch.randn((E, C), generator=g), torch.randn((E, C), generator=g)
    r = routes[:, 1]
    want_sc, want_sh = sc * gamma[r.long()], sh * gamma[r.long()] + beta[r.long()]
    ops.route_affine(sc, sh, gamma, beta, r)
    torch.testing.assert_close(sc, want_sc, atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(sh, want_sh, atol=1e-6, rtol=1e-6)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        ops.dwconv3_routed(torch.zeros((1, 2, 2, 12), dtype=torch.float16), torch.zeros((2, 9, 12), dtype=torch.float16),
                           torch.zeros((1,), dtype=torch.int32), torch.ones((2,), dtype=torch.int32))


def test_scale_coords_kernel_and_pose_predictor(emu):
    """ym_scale_coords bit-exact against the oracle (which test_scale_coords_oracle_matches_reference pins to ops.scale_coords), and
    PosePredictor.construct_result: boxes through ym_scale_boxes, keypoints (n, 17, 3) through ym_scale_coords, visibility untouched."""
    from yolo_master_b200.engine import PosePredictor
    from yolo_master_b200.utils import ops as box_ops
    g = torch.Generator().manual_seed(6)
    for shape0 in ((480, 640), (1080, 1920), (100, 37)):
        for last in (2, 3):
            k = torch.rand((9, 17, last), generator=g) * 700 - 30
            for norm in (False, True):
                got = box_ops.scale_coords((640, 640), k.clone(), shape0, normalize=norm)
                assert np.array_equal(got.numpy(), L.scale_coords((640, 640), k.numpy(), shape0, normalize=norm)), (shape0, last, norm)
    assert box_ops.scale_coords((640, 640), torch.zeros((0, 17, 3)), (480, 640)).shape == (0, 17, 3)
    stub = _TaskStub(None, 1)
    stub.kpt_shape = (17, 3)
    pred = PosePredictor(stub, imgsz=640, device="cpu")
    rows = torch.cat([torch.rand((5, 4), generator=g) * 600, torch.rand((5, 1), generator=g), torch.zeros((5, 1)),
                      torch.rand((5, 51), generator=g) * 600], 1)
    frame = np.zeros((360, 500, 3), np.uint8)
    r = pred.construct_result(rows.clone(), torch.zeros((1, 3, 640, 640)), frame, None)
    want_k = L.scale_coords((640, 640), rows[:, 6:].reshape(5, 17, 3).numpy(), frame.shape)
    assert np.array_equal(r.keypoints.data.numpy(), want_k) and np.array_equal(want_k[..., 2], rows[:, 6:].reshape(5, 17, 3)[..., 2].numpy())
    assert np.array_equal(r.boxes.data[:, :4].numpy(), L.scale_boxes((640, 640), rows[:, :4].numpy(), frame.shape))
    assert r.keypoints.xy.shape == (5, 17, 2) and r.keypoints.conf.shape == (5, 17) and len(r[:2].keypoints) == 2
