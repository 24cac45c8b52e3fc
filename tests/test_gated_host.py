"""Gated MoE family without a GPU.

1. The phase functions the kernels of yolo-master_b200/csrc/gated.cu run (gated_core.cuh) are compiled for the HOST with g++
   (tests/native/gated_host.cpp; every phase is executed for all 256 thread ids before the next one starts, which is what
   __syncthreads() gives the kernels) and compared with the oracle functions pinned to the reference's v0_10 model
   (tests/test_oracle_gated.py): router decisions exactly, everything else within fp32 / fp16 rounding.
2. The host mirror `VisualEnhancedAdaptiveGateMoE` (weight packing, BatchNorm folding, block-diagonal expansion of the grouped
   expert conv, shuffle permutation, op order) is run on CPU tensors with every C-ABI op replaced by its documented semantics
   (tools/cpu_emu.py; the gated ops by the host build above) against the oracle's `layer_visual_enhanced_gate_moe`.
3. The class tree loads the reference's v0_10 state-dict key table with strict=True."""
import ctypes as C
import importlib.util
import json
import os
import subprocess

import pytest
import torch
import torch.nn.functional as F

from _util import GOLD, ROOT, synth_sd_from_keys, yaml_of
from oracle import yolo_master_oracle as O


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("gated_host") / "libgated_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "yolo-master_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "gated_host.cpp"), "-o", so], check=True)
    return C.CDLL(so)


@pytest.fixture()
def emu(host):
    spec = importlib.util.spec_from_file_location("cpu_emu", os.path.join(ROOT, "tools", "cpu_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from yolo_master_b200 import ops
    from yolo_master_b200.nn.modules import _base, block, conv, gated, moa, mot
    saved_ops = {k: v for k, v in vars(ops).items() if callable(v) and not k.startswith("_")}
    saved_nhwc = {m: m.to_nhwc for m in (_base, block, conv, moa, mot, gated) if hasattr(m, "to_nhwc")}
    mod.install()
    mod.install_gated(host)
    yield mod
    for k, v in saved_ops.items():          # the emulation must not leak into other tests of this session
        setattr(ops, k, v)
    for m, f in saved_nhwc.items():
        m.to_nhwc = f


def _block(emu, c, E, k, seed):
    from yolo_master_b200.nn.modules.gated import VisualEnhancedAdaptiveGateMoE
    m, sd = emu.seeded(VisualEnhancedAdaptiveGateMoE(c, c, E, k, 0.5), seed)
    return m, sd


# ------------------------------------------------------------------------------------------------ 1. kernel bodies on the host
@pytest.mark.parametrize("c,E,k,H,W", [(128, 4, 2, 20, 24), (64, 8, 2, 9, 7), (256, 16, 2, 6, 5), (64, 4, 1, 4, 4)])
def test_gate_router_body_matches_oracle(emu, c, E, k, H, W):
    """ym_gate_router (three kernels) against DualStreamGateRouter + the batch-level complexity gate of the oracle: identical
    expert choices, weights to fp32 rounding.  (4, 4) maps exercise the un-pooled branch of gated.py:139-142."""
    from yolo_master_b200 import ops
    m, sd = _block(emu, c, E, k, 11)
    g = torch.Generator().manual_seed(5)
    xd = torch.randn((3, c // 2, H, W), generator=g).half()
    pk = m.get_pack()["router"]
    idx, w, probs = ops.gate_router(xd.permute(0, 2, 3, 1).contiguous(), pk, k)
    xf = xd.float()
    cx = torch.sigmoid(F.conv2d(xf.mean((2, 3), keepdim=True), sd["m.complexity_estimator.1.weight"], sd["m.complexity_estimator.1.bias"])).mean()
    rw, ri, rp = O.dual_stream_gate_router(sd, "m.routing", xf, k, max(1.2, 1e-3))
    rw = O.complexity_gate(rw, cx.clamp(0.3, 1.5))
    assert torch.equal(idx.long(), ri)
    torch.testing.assert_close(probs, rp, atol=2e-6, rtol=1e-4)
    torch.testing.assert_close(w, rw, atol=2e-6, rtol=1e-4)


def test_complexity_gate_drops_ranks(emu):
    """A strongly negative complexity bias gives c = 0.3 -> round(0.6) = 1 rank kept: the second weight is zeroed, the first 1."""
    from yolo_master_b200 import ops
    m, sd = _block(emu, 64, 4, 2, 3)
    pk = dict(m.get_pack()["router"])
    pk["cx_b"] = -50.0
    x = torch.randn((2, 6, 6, 32), generator=torch.Generator().manual_seed(1)).half()
    _, w, _ = ops.gate_router(x, pk, 2)
    assert torch.allclose(w, torch.tensor([[1.0, 0.0], [1.0, 0.0]]), atol=1e-6)
    pk["cx_b"] = 50.0                       # c = 1 -> round(2) = 2 ranks kept
    _, w, _ = ops.gate_router(x, pk, 2)
    assert bool((w > 0).all()) and torch.allclose(w.sum(1), torch.ones(2), atol=1e-5)


def test_fc_gate_body_matches_oracle(emu):
    from yolo_master_b200 import ops
    m, sd = _block(emu, 128, 4, 2, 7)
    x = torch.randn((3, 128, 5, 6), generator=torch.Generator().manual_seed(2)).half()
    pk = m.get_pack()
    v = x.float().mean((2, 3)).half().view(3, 1, 1, 128)                                 # what ym_adaptive_avgpool_nhwc stores
    got = ops.fc_gate(v, pk["se_w1"], pk["se_w2"], pk["se_b2"])
    want = torch.sigmoid(F.linear(F.silu(F.linear(v.view(3, 128).float(), sd["m.se_gate.2.weight"])), sd["m.se_gate.4.weight"], sd["m.se_gate.4.bias"]))
    torch.testing.assert_close(got, want, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(got, O.gated_se_gate(sd, "m.se_gate", x.float()), atol=2e-3, rtol=0)   # vs the un-rounded pooled vector


def test_gated_select_body_matches_oracle(emu):
    """The FusedExpertGroup tail on a given all-expert conv output (strided view of a wider buffer, idx repeated across images)."""
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, H, W, E, oc, G = 3, 7, 5, 4, 16, 8
    buf = torch.randn((B, H, W, E * oc + 8), generator=g).half()
    fo = buf[..., :E * oc]
    idx = torch.tensor([[2, 0], [1, 3], [0, 2]], dtype=torch.int32)
    w = torch.tensor([[0.7, 0.3], [0.55, 0.45], [1.0, 0.0]])
    gamma, beta = torch.randn((E, oc), generator=g), torch.randn((E, oc), generator=g)
    got = ops.gated_select(fo, idx, w, gamma, beta, E, oc, G).float()
    f5 = fo.float().permute(0, 3, 1, 2).reshape(B, E, oc, H, W)
    sel = torch.gather(f5, 1, idx.long().view(B, 2, 1, 1, 1).expand(B, 2, oc, H, W))
    nrm = F.group_norm(sel.reshape(B * 2, oc, H, W), G).view(B, 2, oc, H, W)
    nrm = nrm * gamma[idx.long()].view(B, 2, oc, 1, 1) + beta[idx.long()].view(B, 2, oc, 1, 1)
    want = (F.silu(nrm) * w.view(B, 2, 1, 1, 1)).sum(1).permute(0, 2, 3, 1)
    torch.testing.assert_close(got, want, atol=4e-3, rtol=2e-3)                          # one fp16 rounding of the output


def test_ctx_mean3_body_matches_torch(emu):
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(4)
    for H, W in ((8, 12), (5, 7), (3, 3), (1, 2)):
        a = torch.randn((2, H, W, 16), generator=g).half()
        h2, w2, h4, w4 = max(1, H // 2), max(1, W // 2), max(1, H // 4), max(1, W // 4)
        b, c = torch.randn((2, h2, w2, 16), generator=g).half(), torch.randn((2, h4, w4, 16), generator=g).half()
        up = lambda t: F.interpolate(t.float().permute(0, 3, 1, 2), size=(H, W), mode="nearest").permute(0, 2, 3, 1)
        want = torch.stack([a.float(), up(b), up(c)]).mean(0)
        torch.testing.assert_close(ops.ctx_mean3(a, b, c).float(), want, atol=2e-3, rtol=1e-3)


# ------------------------------------------------------------------------------------------------ 2. host mirror of the block
@pytest.mark.parametrize("c,E,k,H,W,seed", [(128, 4, 2, 20, 16, 1), (128, 8, 2, 10, 12, 2), (256, 16, 2, 6, 6, 3)])
def test_block_host_wiring_matches_oracle(emu, c, E, k, H, W, seed):
    """E = 4 / 8: low-rank fused expert group (dense block-diagonal conv + select); E = 16: shared-inverted group on the grouped
    expert GEMM.  Routing decisions must agree exactly; the output is held to the fp16 noise floor of the oracle."""
    m, sd = _block(emu, c, E, k, seed)
    x = torch.randn((2, c, H, W), generator=torch.Generator().manual_seed(20 + seed)).half().float()
    fn = lambda t, **kw: O.layer_visual_enhanced_gate_moe(sd, "m", t, c, c, E, k, 0.5, **kw)
    with torch.no_grad():
        y = m.fwd_nhwc(x.half().permute(0, 2, 3, 1).contiguous()).float().permute(0, 3, 1, 2)
    ref, rw, ri, _ = fn(x, return_route=True)
    assert torch.equal(m.last_routing_snapshot["topk_indices"].long(), ri)
    torch.testing.assert_close(m.last_routing_snapshot["topk_weights"], rw, atol=5e-3, rtol=0)
    with O.fp16_storage(), O.fp16_weights():
        sim = fn(x)
    assert emu.report(f"VisualEnhancedAdaptiveGateMoE c{c} E{E}", y, ref, sim)


FAMILY = torch.load(os.path.join(GOLD, "gated_family.golden.pt"))


@pytest.mark.parametrize("key", sorted(FAMILY), ids=lambda k: k.replace("AdaptiveGateMoE", "AGM"))
def test_family_class_matches_reference_module_golden(emu, key):
    """Every class of the AdaptiveGateMoE line (v0_4 ... v0_10 zoos), top-2 of 4 and of 16 experts: state-dict keys equal the
    reference module's, and the host mirror (emulated ops) reproduces the REAL module's output within the fp16 noise floor of the
    oracle with identical expert choices."""
    from yolo_master_b200.nn.modules import gated, moe as moe_mod
    from yolo_master_b200.utils.synth import fill_state_dict_
    name, E = key.split("/E")
    c = FAMILY[key]
    extra = () if name == "UltraOptimizedMoE" else (c["split"],)             # (in, out, num_experts, top_k[, split_ratio])
    m = getattr(gated, name, None) or getattr(moe_mod, name)
    m = m(64, 64, int(E), 2, *extra)
    assert getattr(m, "expert_backend", c["backend"]) == c["backend"]
    sd = m.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == c["keys"]
    fill_state_dict_(sd, c["seed"])
    sd.update({k: v.clone() for k, v in c["scalars"].items()})    # 0-dim parameters (and CrossPathGate's bias) as the fixture had them
    m.load_state_dict(sd, strict=True)
    m.eval()
    sdm = {"m." + k: v.clone().float() for k, v in sd.items()}
    x = torch.randn((2, 64, c["hw"], c["hw"]), generator=torch.Generator().manual_seed(c["xseed"]))
    with torch.no_grad():
        y = m.fwd_nhwc(x.half().permute(0, 2, 3, 1).contiguous()).float().permute(0, 3, 1, 2)
    xq = x.half().float()
    ref, _, ri, _ = O._LAYER_FN[name](sdm, "m", xq, 64, 64, int(E), 2, *extra, return_route=True)
    assert torch.equal(m.last_routing_snapshot["topk_indices"].long(), ri)
    with O.fp16_storage(), O.fp16_weights():
        sim = O._LAYER_FN[name](sdm, "m", xq, 64, 64, int(E), 2, *extra)
    assert emu.report(key, y, ref, sim)
    assert float((y - c["y"]).abs().max()) < 3e-2               # and stays close to the reference module's own fp32 output


def test_fused_group_dense_weight_equals_grouped_conv():
    from yolo_master_b200.nn.modules.gated import FusedExpertGroup
    g = FusedExpertGroup(32, 16, 4)
    torch.nn.init.normal_(g.fused_conv.weight)
    x = torch.randn(2, 32, 6, 5)
    assert g.num_groups == 8
    torch.testing.assert_close(F.conv2d(x, g.dense_weight(), None, 1, 1), g.fused_conv(x), atol=1e-5, rtol=1e-5)


# ------------------------------------------------------------------------------------------------ 3. model zoo
def test_v0_10_model_builds_and_loads_reference_keys():
    from yolo_master_b200.nn.modules.gated import VisualEnhancedAdaptiveGateMoE
    from yolo_master_b200.nn.tasks import DetectionModel
    m = DetectionModel("master/v0_10/det/yolo-master-n.yaml")
    assert [type(m.model[i]) for i in (5, 8, 11)] == [VisualEnhancedAdaptiveGateMoE] * 3
    assert [m.model[i].expert_backend for i in (5, 8, 11)] == ["low_rank_fused", "low_rank_fused", "shared_inverted"]
    keys = json.load(open(os.path.join(GOLD, "yolo-master-n-v0_10.keys.json")))
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert set(mine) == set(keys), (sorted(set(mine) ^ set(keys))[:10])
    assert all(mine[k] == keys[k][0] for k in keys)
    m.load_state_dict(synth_sd_from_keys(0, "yolo-master-n-v0_10"), strict=True)
    spec = O.parse_spec(yaml_of("master/v0_10/det/yolo-master-n.yaml"))
    assert [L["type"] for L in spec["layers"]].count("VisualEnhancedAdaptiveGateMoE") == 3
    for sc in "smlx":
        assert DetectionModel(f"master/v0_10/det/yolo-master-{sc}.yaml") is not None


def test_gated_zoo_yamls_build():
    """v0_4 ... v0_10 model zoos (35 stock YAMLs): every file parses into its family's block with the reference's back-end choice and
    parameter count (reference counts taken with ultralytics' own DetectionModel on the n scale)."""
    from yolo_master_b200.nn.tasks import DetectionModel
    want = {"v0_4": ("AdaptiveGateMoE", ["shared_inverted"] * 3, None), "v0_5": ("FusedAdaptiveGateMoE", ["fused"] * 3, None),
            "v0_6": ("HybridAdaptiveGateMoE", ["fused", "fused", "shared_inverted"], None),
            "v0_7": ("LowRankHybridAdaptiveGateMoE", ["low_rank_fused", "low_rank_fused", "shared_inverted"], 3106914),
            "v0_8": ("RefinedLowRankHybridAdaptiveGateMoE", ["low_rank_fused", "low_rank_fused", "shared_inverted"], 3137637),
            "v0_9": ("DetailAwareLowRankHybridAdaptiveGateMoE", ["low_rank_fused", "low_rank_fused", "shared_inverted"], 3116133),
            "v0_10": ("VisualEnhancedAdaptiveGateMoE", ["low_rank_fused", "low_rank_fused", "shared_inverted"], 3449963)}
    for ver, (cls, backends, nparam) in want.items():
        m = DetectionModel(f"master/{ver}/det/yolo-master-n.yaml")
        assert [type(m.model[i]).__name__ for i in (5, 8, 11)] == [cls] * 3
        assert [m.model[i].expert_backend for i in (5, 8, 11)] == backends
        if nparam:
            assert sum(p.numel() for p in m.parameters()) == nparam
        assert DetectionModel(f"master/{ver}/det/yolo-master-s.yaml").model[5].in_channels == 256
    assert type(DetectionModel("master/v0/det/yolo-master-n.yaml").model[3]).__name__ == "ES_MOE"
    with pytest.raises(FileNotFoundError, match="Multiple files match"):       # 16 zoo versions ship this file name: like check_yaml
        DetectionModel("yolo-master-n.yaml")
    for cfg, cls in (("master/v0_12/det/yolo-master-n.yaml", "OptimalHybridGateMoE"), ("master/exp/yolo-master-v0_11.yaml", "HybridAdaptiveGateMoEv2"),
                     ("master/v0_13/det/yolo-master-n.yaml", "MultiHeadRouterMoE"), ("master/v0_15/det/yolo-master-n.yaml", "GatedFusionMoE"),
                     ("master/v0_14/det/yolo-master-n.yaml", "DiversifiedExpertMoE")):
        m = DetectionModel(cfg)
        assert [type(m.model[i]).__name__ for i in (5, 8, 11)] == [cls] * 3 and m.model[11].dynamic_channels == 96
        assert type(m.model[5].routing).__name__ == ("MultiHeadRouterV3" if cls == "MultiHeadRouterMoE" else "DualStreamGateRouterV2")
    from yolo_master_b200.nn.modules.gated import SharedExpertMoE
    SharedExpertMoE.reset_shared_pools()                        # v0_8 shared-expert model: two blocks, one expert group, 2 618 441 parameters
    m = DetectionModel("master/v0_8/det/yolo-master-moe-mot-shared-n.yaml")
    assert m.model[5].fused_experts is m.model[8].fused_experts and (m.model[5]._is_pool_owner, m.model[8]._is_pool_owner) == (True, False)
    assert sum(p.numel() for p in m.parameters()) == 2618441 and len(m.state_dict()) == 838
    SharedExpertMoE.reset_shared_pools()
    for sc in "nsmlx":                                          # v0_3 zoo: UltimateOptimizedMoE
        m = DetectionModel(f"master/v0_3/det/yolo-master-{sc}.yaml")
        assert [type(m.model[i]).__name__ for i in (5, 8, 11)] == ["UltimateOptimizedMoE"] * 3
    for sc in "nsmlx":                                          # v0_1 zoo: ModularRouterExpertMoE is OptimizedMOEImproved
        m = DetectionModel(f"master/v0_1/det/yolo-master-{sc}.yaml")
        assert [type(m.model[i]).__name__ for i in (5, 8, 11)] == ["OptimizedMOEImproved"] * 3 and m.model[5].add_residual
