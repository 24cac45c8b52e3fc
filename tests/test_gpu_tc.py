"""tcgen05 kernels (TMEM accumulators, swizzled smem operands) through the C ABI: plain GEMM and the ES-MoE dispatch."""
import pytest
import torch

from _util import assert_close
from oracle.moe_dispatch_oracle import compute_sparse_experts_batched, conv1x1_experts

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 64, 128), (1000, 128, 64), (6400, 64, 192), (777, 256, 256),
                                   (512, 192, 72), (128, 384, 384), (300, 80, 256)])
def test_tc_gemm_matches_fp32(M, N, K):
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((M, K), generator=g).half()
    b = (torch.randn((N, K), generator=g) / K ** 0.5).half()
    bias = torch.randn((N,), generator=g)
    out = ops.tc_gemm_nt(a.to(DEV), b.to(DEV), bias.to(DEV))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    assert_close(out, ref, what=f"tc_gemm {M}x{N}x{K}")


def test_tc_gemm_epilogue_and_pitched_views():
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(1)
    M, N, K = 640, 64, 128
    abuf = torch.randn((M, K + 64), generator=g).half().to(DEV)
    a = abuf[:, 32:32 + K]                      # channel slice of a wider NHWC buffer (pitch 192)
    b = (torch.randn((N, K), generator=g) / K ** 0.5).half().to(DEV)
    res = torch.randn((M, N), generator=g).half().to(DEV)
    obuf = torch.zeros((M, 2 * N), dtype=torch.float16, device=DEV)
    out = ops.tc_gemm_nt(a, b, None, res, act=True, out=obuf[:, N:])
    ref = torch.nn.functional.silu(a.float().cpu() @ b.float().cpu().t()) + res.float().cpu()
    assert_close(out, ref, what="tc_gemm silu+res")
    assert float(obuf[:, :N].abs().max()) == 0.0


@pytest.mark.parametrize("B,hw,C,E,k", [(8, (16, 16), 64, 4, 2), (64, (32, 32), 256, 8, 2), (5, (9, 13), 128, 8, 2), (6, (20, 20), 256, 16, 1)])
def test_moe_dispatch_vs_oracle(B, hw, C, E, k):
    """C5 configuration (x=(64,256,32,32), 8 experts, top-2) and ragged variants against the restated dispatcher."""
    from yolo_master_b200 import ops
    g = torch.Generator().manual_seed(B * 7 + E)
    x = torch.randn((B, C, *hw), generator=g).half()
    W = (torch.randn((E, C, C), generator=g) / C ** 0.5).half()
    idx = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(B)]).int()
    w = torch.rand((B, k), generator=g)
    w = w / w.sum(1, keepdim=True)
    w[0, -1] = 0.005                                   # below the 0.01 eval threshold: that route must be dropped
    out = ops.moe_dispatch(x.to(DEV).permute(0, 2, 3, 1).contiguous(), W.to(DEV), idx.to(DEV), w.to(DEV))
    torch.cuda.synchronize()
    ref = compute_sparse_experts_batched(x.float(), conv1x1_experts(W.float()), w, idx.long(), C)
    assert_close(out.permute(0, 3, 1, 2), ref, what="moe_dispatch")
