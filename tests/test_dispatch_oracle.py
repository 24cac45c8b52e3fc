"""Pins the dispatch oracle to the real reference `BatchedExpertComputation` (fixture from tests/golden/make_golden.py)."""
import os

import torch

from _util import GOLD
from oracle.moe_dispatch_oracle import compute_sparse_experts_batched, conv1x1_experts


def test_dispatch_oracle_matches_reference():
    g = torch.load(os.path.join(GOLD, "dispatch.golden.pt"))
    for c in g["cases"]:
        out = compute_sparse_experts_batched(c["x"], conv1x1_experts(c["W"]), c["w"], c["idx"], c["W"].shape[1])
        torch.testing.assert_close(out, c["out"], atol=1e-6, rtol=1e-6)
