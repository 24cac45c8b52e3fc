"""Pins the ES_MOE oracle to the real reference module (fixture: tests/golden/make_golden.py::esmoe_golden)."""
import os

import torch

from _util import GOLD
from oracle import yolo_master_oracle as O
from yolo_master_b200.utils.synth import fill_state_dict_

ESMOE_CASES = [(64, 4, 2, 20, 24, 6), (32, 4, 2, 9, 7, 5), (128, 4, 2, 10, 10, 4)]   # C, E, top_k, H, W, B (make_golden.py)


def esmoe_case(i):
    """(state_dict with 'm.' prefix, input x, reference output y) rebuilt from key names + seeds."""
    c = torch.load(os.path.join(GOLD, "esmoe.golden.pt"))["cases"][i]
    C, E, k, H, W, B = ESMOE_CASES[i]
    sd = {kk: (torch.zeros(shape, dtype=torch.int64) if kk.endswith("num_batches_tracked") else torch.zeros(shape))
          for kk, shape in c["keys"].items()}
    fill_state_dict_(sd, 40 + c["seed"])
    for kk in sd:
        if kk.endswith("routing_network.2.weight"):
            sd[kk] *= 6
    x = torch.randn((B, C, H, W), generator=torch.Generator().manual_seed(c["seed"]))
    return sd, x, c["y"]


def test_esmoe_oracle_matches_reference():
    for i, (C, E, k, H, W, B) in enumerate(ESMOE_CASES):
        sd, x, y = esmoe_case(i)
        yo, (ti, w) = O.es_moe({"m." + kk: v for kk, v in sd.items()}, "m", x, C, C, E, 8, k)
        torch.testing.assert_close(yo, y, atol=2e-6, rtol=1e-5)
        assert set((w > 0).sum(1).tolist()) == {1, 2}, "fixture must exercise both sides of the dynamic threshold"
