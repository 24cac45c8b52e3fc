"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host mirror keeps the reference's
names / signatures / state_dict keys, and the product path refuses to run without CUDA (no silent fallback)."""
import ctypes
import inspect
import json
import os
import re

import pytest
import torch

from _util import GOLD, ROOT, yaml_n


def test_library_exports_every_header_symbol():
    from yolo_master_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "ym_b200.h")).read()
    declared = set(re.findall(r"\b(ym_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ym_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.ym_version() >= 100
    # the library targets sm_100a only
    so = os.path.join(ROOT, "yolo-master_b200", "libym_b200.so")
    assert os.path.getsize(so) > 100_000


def test_state_dict_keys_match_reference():
    from yolo_master_b200.nn.tasks import DetectionModel

    m = DetectionModel("yolo26-master-n.yaml")
    ref = json.load(open(os.path.join(GOLD, "yolo26-master-n.keys.json")))
    sd = m.state_dict()
    assert set(sd) == set(ref)
    for k, (shape, dt) in ref.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dt, k
    assert m.stride.tolist() == [8.0, 16.0, 32.0]
    assert m.end2end is True


def test_signatures_match_reference():
    from yolo_master_b200.nn import modules as M

    expect = {
        "Conv": ["c1", "c2", "k", "s", "p", "g", "d", "act"],
        "C2f": ["c1", "c2", "n", "shortcut", "g", "e"],
        "C3k2": ["c1", "c2", "n", "c3k", "e", "attn", "g", "shortcut"],
        "SPPF": ["c1", "c2", "k", "n", "shortcut"],
        "C2PSA": ["c1", "c2", "n", "e"],
        "A2C2f": ["c1", "c2", "n", "a2", "area", "residual", "mlp_ratio", "e", "g", "shortcut"],
        "A2C2fMoE": ["c1", "c2", "n", "a2", "area", "residual", "mlp_ratio", "e", "g", "shortcut", "num_experts", "top_k",
                     "expert_type"],
        "Detect": ["nc", "reg_max", "end2end", "ch"],
        "Concat": ["dimension"],
    }
    for name, params in expect.items():
        got = [p for p in inspect.signature(getattr(M, name).__init__).parameters if p != "self"]
        assert got == params, (name, got)


def test_parse_model_matches_oracle_spec():
    from oracle import yolo_master_oracle as O
    from yolo_master_b200.nn.tasks import DetectionModel

    m = DetectionModel("yolo26-master-n.yaml")
    spec = O.parse_spec(yaml_n())
    assert len(spec["layers"]) == len(m.model) == 24
    assert spec["save"] == m.save
    for L, mod in zip(spec["layers"], m.model):
        assert L["type"].split(".")[-1] == mod.type, (L, mod.type)


def test_no_cpu_fallback():
    from yolo_master_b200.nn.tasks import DetectionModel

    m = DetectionModel("yolo26-master-n.yaml")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="inference-only"):
        m.train()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "yolo-master_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(root, f)


@pytest.mark.parametrize("name,cfg,scale", [
    ("yolo-master-n-v0", "master/v0/det/yolo-master-n.yaml", None), ("yolo-master-l-v0", "master/v0/det/yolo-master-l.yaml", None),
    ("yolo26-master-moa-mot-n", "yolo26-master-moa-mot-n.yaml", None), ("yolo26-master-moa-mot-s", "yolo26-master-moa-mot-n.yaml", [0.50, 0.50, 1024])])
def test_state_dict_keys_match_reference_all_families(name, cfg, scale):
    """Stock YAMLs of the v0 (ES_MOE / A2C2f / DFL) and MoT + MoA families build with the reference's exact state_dict layout
    (key tables exported from the real reference by tests/golden/make_golden.py)."""
    from yolo_master_b200.nn.tasks import DetectionModel, yaml_model_load

    d = yaml_model_load(cfg)
    if scale is not None:
        d["scales"]["s"] = scale
        d["scale"] = "s"
    m = DetectionModel(d)
    ref = json.load(open(os.path.join(GOLD, f"{name}.keys.json")))
    sd = m.state_dict()
    assert set(sd) == set(ref), sorted(set(sd) ^ set(ref))[:10]
    for k, (shape, dt) in ref.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dt, k
    assert m.stride.tolist() == [8.0, 16.0, 32.0]
    assert m.end2end is ("moa-mot" in name)


def test_mixture_module_signatures_match_reference():
    from yolo_master_b200.nn import modules as M

    expect = {
        "ES_MOE": ["in_channels", "out_channels", "num_experts", "reduction", "top_k", "use_sparse_inference", "dynamic_threshold",
                   "max_kernel_size", "expert_kernel_sizes"],
        "C2fMoT": ["c1", "c2", "n", "num_heads", "top_k", "window_size", "n_points", "mlp_ratio", "temperature", "balance_loss_coeff", "e",
                   "sparse_train", "scene_aware_router", "scene_hidden_dim", "scene_consistency_coeff", "sparse_train_warmup_steps",
                   "scene_inference_mode", "local_attn_window"],
        "C2fMoA": ["c1", "c2", "n", "num_heads", "mlp_ratio", "temperature", "shortcut", "e", "aux_loss_coeff", "local_window_size",
                   "sequential_heads", "regional_max_kv_tokens", "sparse_inference", "sparse_inference_threshold",
                   "inference_sparse_threshold"],
    }
    for name, params in expect.items():
        got = [p for p in inspect.signature(getattr(M, name).__init__).parameters if p != "self"]
        assert got == params, (name, got)


def test_host_side_helpers_of_the_c_abi_need_no_gpu():
    """Pure host entry points: the router's tile grid (ym_router_blocks, routers.py:289-292 pooling rule), the A/B switches (set returns the
    previous value and rejects out-of-range arguments), the expert-FFN strip count."""
    import ctypes as C
    from yolo_master_b200 import _lib
    L = _lib.load()
    for (H, W, pool) in [(80, 80, 4), (40, 40, 4), (20, 20, 4), (4, 4, 4), (37, 23, 4), (160, 96, 2)]:
        do_pool = H > pool and W > pool
        ps = pool if do_pool else 1
        Hp, Wp = H // ps, W // ps
        npix = C.c_int(0)
        nblk = L.ym_router_blocks(H, W, pool, C.addressof(npix))
        assert npix.value == Hp * Wp and nblk == ((Wp + 15) // 16) * ((Hp + 3) // 4), (H, W, pool, nblk, npix.value)
    for setter, good, bad in [(L.ym_set_small_conv_impl, (0, 1), 2), (L.ym_set_stem_impl, (0, 1), 5), (L.ym_set_dwconv_tc, (0, 1), -1),
                              (L.ym_set_conv2_epi_groups, (1, 2), 3), (L.ym_set_attention2_qtiles, (0, 1, 2), 3),
                              (L.ym_set_attention2_variant, tuple(range(8)), 8), (L.ym_set_kernel_priority, (0, -3, 3), 9)]:
        first = setter(good[0])
        try:
            for v in good:
                prev = setter(v)
                assert setter(bad) == v and setter(v) == v, (setter, v)      # a rejected value leaves the setting alone
                assert prev in good or prev == first
        finally:
            setter(first)
    for HW, P in [(6400, 64), (1600, 64), (400, 64), (117, 6)]:
        strips = L.ym_moe_ffn_strips(HW, P)
        assert 1 <= strips <= (HW + 127) // 128
