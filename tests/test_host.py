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
