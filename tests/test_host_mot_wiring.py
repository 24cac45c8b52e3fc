"""HOST logic of the MoT / MoA modules (weight packing, layer-scale folding into the projections, q|v and gate|value GEMM fusion,
head padding, window pad vectors, op order) checked on the CPU: the C-ABI ops are replaced by torch restatements of their
documented semantics (tools/cpu_emu.py - test infrastructure, never shipped), and every module is compared with the oracle
function of the same name.  The CUDA kernels behind those ops are verified on the GPU by tests/test_gpu_mot.py."""
import importlib.util
import os
import sys

import pytest
import torch

from _util import ROOT
from oracle import yolo_master_oracle as O


@pytest.fixture(scope="module")
def emu():
    spec = importlib.util.spec_from_file_location("cpu_emu", os.path.join(ROOT, "tools", "cpu_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from yolo_master_b200 import ops
    from yolo_master_b200.nn.modules import _base, block, conv, moa, mot
    saved_ops = {k: v for k, v in vars(ops).items() if callable(v) and not k.startswith("_")}
    saved_nhwc = {m: m.to_nhwc for m in (_base, block, conv, moa, mot) if hasattr(m, "to_nhwc")}
    mod.install()
    yield mod
    for k, v in saved_ops.items():          # the emulation must not leak into other tests of this session
        setattr(ops, k, v)
    for m, f in saved_nhwc.items():
        m.to_nhwc = f


def _run(emu, mod, fn, x, what):
    with torch.no_grad():
        y = mod.fwd_nhwc(x.half().permute(0, 2, 3, 1).contiguous()).float().permute(0, 3, 1, 2)
    ref = fn(x)
    with O.fp16_storage(), O.fp16_weights():
        sim = fn(x)
    assert emu.report(what, y, ref, sim), what


def test_mot_modules_host_wiring(emu):
    from yolo_master_b200.nn.modules import mot
    g = torch.Generator().manual_seed(0)
    dim, nh, H, W = 64, 8, 10, 9
    x = torch.randn((2, dim, H, W), generator=g).half().float()
    m, sd = emu.seeded(mot._LocalConvTransformerExpert(dim, nh), 1)
    _run(emu, m, lambda t: O.mot_local_expert(sd, "m", t, nh), x, "LocalConv expert")
    m, sd = emu.seeded(mot._LocalConvTransformerExpert(dim, nh, local_window_size=4), 2)
    _run(emu, m, lambda t: O.mot_local_expert(sd, "m", t, nh, 4), x, "LocalConv expert (windowed)")
    m, sd = emu.seeded(mot._WindowTransformerExpert(dim, nh, 7, shift_size=1), 3)
    _run(emu, m, lambda t: O.mot_window_expert(sd, "m", t, nh, 7, 3), x, "Window expert (shifted, padded)")
    m, sd = emu.seeded(mot._DeformableTransformerExpert(dim, nh), 5)
    _run(emu, m, lambda t: O.mot_deform_expert(sd, "m", t, nh), x, "Deformable expert")
    m, sd = emu.seeded(mot.MoTBlock(dim, nh, 2, temperature=0.8), 7)
    _run(emu, m, lambda t: O.mot_block(sd, "m", t, nh, 2), x, "MoTBlock")
    m, sd = emu.seeded(mot.C2fMoT(64, 128, 1, 8, 1), 8)
    _run(emu, m, lambda t: O.layer_c2f_mot(sd, "m", t, 64, 128, 1, 8, 1), x, "C2fMoT top-1")


@pytest.mark.parametrize("dim,heads,H,W", [(32, 3, 24, 24), (32, 3, 22, 22), (64, 3, 8, 10)])
def test_moa_modules_host_wiring(emu, dim, heads, H, W):
    """N = 576 -> linear attention, 484 -> exact/linear blend, 80 -> exact; dim 64 with 3 heads -> head_dim 21 padded to 24."""
    from yolo_master_b200.nn.modules import moa
    x = torch.randn((2, dim, H, W), generator=torch.Generator().manual_seed(dim + H)).half().float()
    m, sd = emu.seeded(moa.MoABlock(dim, heads, temperature=0.8), 5)
    _run(emu, m, lambda t: O.moa_block(sd, "m", t, heads, 0.8), x, f"MoABlock {dim} {H}x{W}")
    m, sd = emu.seeded(moa.C2fMoA(dim, dim, 1, 3, 2.0, 0.8, True), 9)
    _run(emu, m, lambda t: O.layer_c2f_moa(sd, "m", t, dim, dim, 1, 3, 2.0, 0.8, True), x, f"C2fMoA {dim} {H}x{W}")
