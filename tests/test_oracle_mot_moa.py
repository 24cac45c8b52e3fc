"""Pins the oracle's Mixture-of-Transformers / Mixture-of-Attention restatement (C2fMoT, C2fMoA: routers, the three
transformer experts, the three attention head groups) to outputs of the REAL reference model
(tests/golden/make_golden.py: yolo26-master-moa-mot-n and its injected s scale)."""
import os

import pytest
import torch

from _util import GOLD, synth_sd_from_keys, yaml_of
from oracle import yolo_master_oracle as O
from yolo_master_b200.utils.synth import synth_images

CASES = [("yolo26-master-moa-mot-n", None, "b2_224"), ("yolo26-master-moa-mot-n", None, "b1_96"),
         ("yolo26-master-moa-mot-s", [0.50, 0.50, 1024], "b1_160")]


def spec_for(scale):
    d = yaml_of("26/yolo26-master-moa-mot-n.yaml")
    if scale is not None:
        d["scales"]["s"] = scale
        d["scale"] = "s"
    return O.parse_spec(d)


@pytest.mark.parametrize("name,scale,tag", CASES)
def test_oracle_matches_reference_mot_moa(name, scale, tag):
    c = torch.load(os.path.join(GOLD, f"{name}.golden.pt"))["cases"][tag]
    sd = synth_sd_from_keys(0, name)
    spec = spec_for(scale)
    x = synth_images(c["B"], c["H"], c["W"], c["seed"])
    O.ROUTE_TAP = {}
    try:
        y, ys = O.forward(spec, sd, x, return_layers=True)
        tap = O.ROUTE_TAP
    finally:
        O.ROUTE_TAP = None
    for i, ref in c["layers"].items():
        torch.testing.assert_close(ys[i], ref, atol=3e-4, rtol=1e-4, msg=lambda m, i=i: f"layer {i}: {m}")
    assert len(c["routes"]) == 4 and set(c["routes"]) == set(tap)
    for n, r in c["routes"].items():
        torch.testing.assert_close(tap[n][0], r[0], atol=1e-5, rtol=1e-4, msg=lambda m, n=n: f"{n} weights: {m}")
        if len(r) > 1:   # MoT: per-token top-k expert indices bit-exact
            assert torch.equal(tap[n][1].to(torch.int8), r[1]), n
    braw, sraw, _ = ys["detect_raw"]
    torch.testing.assert_close(braw, c["head_boxes"], atol=3e-4, rtol=1e-4)
    torch.testing.assert_close(sraw, c["head_scores"], atol=3e-4, rtol=1e-4)
    torch.testing.assert_close(y, c["final"], atol=2e-3, rtol=1e-4)
    assert torch.equal(y[..., 5], c["final"][..., 5])
