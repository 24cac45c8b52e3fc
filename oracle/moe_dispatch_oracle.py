"""CPU oracle for the generic image-level expert dispatcher.  TEST INFRASTRUCTURE ONLY (see yolo_master_oracle.py).

Restates `BatchedExpertComputation.compute_sparse_experts_batched` (reference moe/utils.py:119-209), eval branch:
routes with weight <= 0.01 are dropped (:172-173), each expert runs on the images routed to it, the output is multiplied
by the routing weight in fp32 and accumulated with index_add_ in x.dtype (:200-203), final clamp to +-1e4 (:207).
Pinned against the real reference class by tests/golden/make_golden.py -> dispatch.golden.pt.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


def compute_sparse_experts_batched(x: torch.Tensor, experts: Sequence[Callable[[torch.Tensor], torch.Tensor]],
                                   routing_weights: torch.Tensor, routing_indices: torch.Tensor, out_channels: int,
                                   training: bool = False) -> torch.Tensor:
    B, C, H, W = x.shape
    k = routing_indices.shape[1]
    idx = routing_indices.reshape(B, -1)[:, :k]
    wts = routing_weights.reshape(B, -1)[:, :k]
    thr = 0.0 if training else 0.01
    valid = wts > thr
    out = torch.zeros(B, out_channels, H, W, dtype=x.dtype, device=x.device)
    for e in range(len(experts)):
        mask = (idx == e) & valid
        if not mask.any():
            continue
        bi, ki = torch.where(mask)
        y = experts[e](x[bi])
        out.index_add_(0, bi, (y.float() * wts[bi, ki].view(-1, 1, 1, 1).float()).to(out.dtype))
    return out.clamp_(-1e4, 1e4)


def conv1x1_experts(weights: torch.Tensor):
    """Experts = nn.Conv2d(C, N, 1, bias=False) with weights [E, N, C] (the C5 microbenchmark configuration)."""
    return [lambda t, w=w: torch.nn.functional.conv2d(t, w.view(w.shape[0], w.shape[1], 1, 1)) for w in weights]
