"""GPU counterparts of the reference's own known-answer / property tests (SURVEY.md §4 "parity anchors")."""
import pytest
import torch

from yolo_master_b200 import ops
from yolo_master_b200.nn.modules import moe as MM
from yolo_master_b200.nn.modules import mot as MT
from yolo_master_b200.utils.synth import fill_state_dict_

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_mot_router_one_hot_and_normalised():
    """tests/test_mot.py:262-273 on the CUDA router: top_k=1 -> one non-zero fp32 weight per token, sum 1; top_k=2 -> two."""
    for k in (1, 2):
        r = MT._MoTRouter(64, 3, k)
        sd = r.state_dict()
        fill_state_dict_(sd, 5)
        r.load_state_dict(sd)
        r.to(DEV).eval()
        x = torch.randn((2, 64, 12, 10), generator=torch.Generator().manual_seed(0)).half().to(DEV)
        with torch.no_grad():
            w, idx = r(x)                                            # module API: NCHW weights in x.dtype, int64 indices
            wf, _ = r.route(x.permute(0, 2, 3, 1).contiguous())     # kernel output: dense fp32 (B,H,W,E)
        assert w.shape == (2, 3, 12, 10) and idx.shape == (2, k, 12, 10) and idx.dtype == torch.int64
        assert wf.dtype == torch.float32
        assert torch.allclose(wf.sum(-1), torch.ones_like(wf[..., 0]), atol=1e-5)
        assert int((wf > 0).sum(-1).max()) == k and int((wf > 0).sum(-1).min()) == k


def test_es_moe_route_threshold_known_answers():
    """tests/test_moe.py:409-420 on ym_esmoe_route: with crafted routing logits the rank-1 expert below the dynamic threshold is
    dropped and the retained mass is renormalised to 1; above it both stay."""
    m = MM.ES_MOE(16, 16, num_experts=3, top_k=2, dynamic_threshold=0.5)
    sd = m.state_dict()
    fill_state_dict_(sd, 1)
    # routing_network: GAP -> 1x1 (C -> Cr) -> SiLU -> 1x1 (-> E).  Zero the last weight and put the logits in its bias.
    sd["routing.routing_network.2.weight"].zero_()
    sd["routing.routing_network.2.bias"].copy_(torch.log(torch.tensor([0.6, 0.4, 1e-6])))
    m.load_state_dict(sd)
    m.to(DEV).eval()
    x = torch.randn((2, 16, 8, 8), generator=torch.Generator().manual_seed(0)).half().to(DEV)
    with torch.no_grad():
        m(x)
    snap = m.last_routing_snapshot
    w, idx = snap["topk_weights"].float().cpu(), snap["topk_indices"].cpu()
    assert idx[:, 0].tolist() == [0, 0]
    assert torch.allclose(w[:, 0], torch.ones(2), atol=1e-5) and torch.allclose(w[:, 1], torch.zeros(2), atol=1e-6)
    m.dynamic_threshold = 0.3
    with torch.no_grad():
        m(x)
    w = m.last_routing_snapshot["topk_weights"].float().cpu()
    assert torch.allclose(w, torch.tensor([[0.6, 0.4], [0.6, 0.4]]), atol=1e-4)


def test_dispatch_drops_low_weight_routes_in_eval():
    """moe/utils.py:172-173 (tests/test_moe.py:768-789 cover the training side): eval routes with weight <= 0.01 contribute nothing."""
    B, C, H, W, E = 4, 64, 8, 8, 4
    g = torch.Generator().manual_seed(0)
    x = torch.randn((B, H, W, C), generator=g).half().to(DEV)
    Wt = (torch.randn((E, C, C), generator=g) / C ** 0.5).half().to(DEV)
    idx = torch.tensor([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=torch.int32, device=DEV)
    w = torch.tensor([[0.995, 0.005], [0.5, 0.5], [0.005, 0.995], [0.009, 0.009]], device=DEV)
    out = ops.moe_dispatch(x, Wt, idx, w).float().cpu()
    xf, Wf = x.float().cpu(), Wt.float().cpu()
    y = lambda b, e: (xf[b].reshape(-1, C) @ Wf[e].t()).reshape(H, W, C)
    torch.testing.assert_close(out[0], y(0, 0) * 0.995, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(out[2], y(2, 3) * 0.995, atol=2e-2, rtol=2e-2)
    assert float(out[3].abs().max()) == 0.0
