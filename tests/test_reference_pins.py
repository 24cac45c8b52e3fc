"""The reference's own known-answer / property tests for this path (SURVEY.md §4 "parity anchors"), replayed against the
oracle and the host mirror (CPU).  Their GPU counterparts live in tests/test_gpu_pins.py."""
import torch

from oracle import yolo_master_oracle as O
from yolo_master_b200.utils.nms import non_max_suppression
from yolo_master_b200.utils.synth import fill_state_dict_


def test_nms_end2end_classes_before_max_det():
    """tests/test_python.py:1249-1268 of the reference, verbatim inputs and expectations, on this package's function."""
    pred = torch.tensor(
        [[[0, 0, 9, 9, 0.9, 5], [1, 1, 9, 9, 0.8, 0], [2, 2, 9, 9, 0.7, 0], [3, 3, 9, 9, 0.6, 0]],
         [[0, 0, 9, 9, 0.9, 0], [1, 1, 9, 9, 0.8, 5], [2, 2, 9, 9, 0.7, 5], [3, 3, 9, 9, 0.6, 0]]], dtype=torch.float32)
    outputs, indices = non_max_suppression(pred, conf_thres=0.25, classes=[0], max_det=2, return_idxs=True)
    for out, idx, confs, expected in zip(outputs, indices, ([0.8, 0.7], [0.9, 0.6]), ([1, 2], [0, 3])):
        assert out.shape[0] == 2 and (out[:, 5] == 0).all()
        assert torch.allclose(out[:, 4], torch.tensor(confs))
        assert idx.tolist() == expected
    out = non_max_suppression(pred, conf_thres=0.25, max_det=2)[0]
    assert torch.allclose(out[:, 4], torch.tensor([0.9, 0.8]))


def test_es_moe_threshold_renormalises_retained_mass():
    """tests/test_moe.py:409-420: weights (0.6, 0.4, 0), top-2, threshold 0.5 -> the rank-1 expert is dropped and the kept one
    carries the whole mass (identity experts then return x)."""
    ti, w, kept = O.es_moe_retained_weights(torch.tensor([[0.6, 0.4, 0.0]]), 2, 0.5)
    assert ti.tolist() == [[0, 1]] and kept.tolist() == [[True, False, False]]
    assert torch.allclose(w, torch.tensor([[1.0, 0.0, 0.0]]))
    _, w, kept = O.es_moe_retained_weights(torch.tensor([[0.45, 0.55, 0.0]]), 2, 0.4)     # both above the threshold: unchanged
    assert kept.tolist() == [[True, True, False]] and torch.allclose(w, torch.tensor([[0.45, 0.55, 0.0]]))


def _mot_sd(dim, nh, seed):
    """State dict of one MoTBlock with the reference's key names, built from this package's mirror class (CPU construction)."""
    from yolo_master_b200.nn.modules.mot import MoTBlock
    sd = MoTBlock(dim, nh, 2).state_dict()
    fill_state_dict_(sd, seed)
    return {"m." + k: v.float() for k, v in sd.items()}


def test_mot_sample_sparse_equals_dense_blend():
    """tests/test_mot_sparse_parity.py:8-22: the eval sample-sparse dispatch equals sum_e expert_e(x) * w_e (atol 1e-5, rtol 1e-4)."""
    dim, nh = 24, 3
    sd = _mot_sd(dim, nh, 3)
    x = torch.randn((2, dim, 6, 6), generator=torch.Generator().manual_seed(0))
    w, idx, _ = O.mot_router(sd, "m.router", x, 1)      # top-1: some experts are inactive for a whole image
    dense = (O.mot_local_expert(sd, "m.experts.0", x, nh) * w[:, 0:1] + O.mot_window_expert(sd, "m.experts.1", x, nh, 7, 0) * w[:, 1:2]
             + O.mot_deform_expert(sd, "m.experts.2", x, nh, 4) * w[:, 2:3])
    import torch.nn.functional as F
    ref = O._gn(sd, "m.out_norm", F.conv2d(dense, sd["m.out_proj.weight"]), O.get_safe_groups(dim, 8)) + x
    torch.testing.assert_close(O.mot_block(sd, "m", x, nh, 1), ref, atol=1e-5, rtol=1e-4)


def test_mot_eval_routing_is_one_hot_for_top1():
    """tests/test_mot.py:262-273: top_k=1 -> exactly one non-zero weight per token and the weights sum to one."""
    sd = _mot_sd(24, 3, 4)
    x = torch.randn((1, 24, 6, 6), generator=torch.Generator().manual_seed(0))
    w, idx, _ = O.mot_router(sd, "m.router", x, 1)
    assert w.shape == (1, 3, 6, 6) and idx.shape == (1, 1, 6, 6)
    assert torch.allclose(w.sum(1), torch.ones_like(w[:, 0])) and int((w > 0).sum(1).max()) == 1


def test_router_fp32_contract():
    """tests/test_moe_router_boundaries.py:393-419, test_mixture_numeric.py:103-153: routing weights fp32, indices int64, sum 1."""
    from _util import synth_sd_from_keys
    sd = synth_sd_from_keys(0)
    x = torch.randn((3, 64, 16, 16), generator=torch.Generator().manual_seed(1)).half().float()
    w, idx, probs = O.efficient_spatial_router(sd, "model.4.m.0.0.mlp.routing", x, 2)
    assert w.dtype == torch.float32 and idx.dtype == torch.int64 and probs.dtype == torch.float32
    assert torch.allclose(w.sum(1), torch.ones(3)) and torch.allclose(probs.sum(1), torch.ones(3), atol=1e-6)


# ----------------------------------------------------------------------------------------------------------------------
# Error conventions of the routed modules: tests/test_moe_router_boundaries.py:66-112,175-190 of the reference,
# replayed on this package's router.  All of these fire before any kernel is launched, so they run without a GPU.
def test_router_input_validation_matches_reference_error_types():
    import pytest

    from yolo_master_b200.nn.modules.moe import EfficientSpatialRouter, _validate_router_input
    from yolo_master_b200.utils.errors import MoERouterError, ShapeMismatchError, YOLOMasterError
    C = 64
    _validate_router_input(torch.randn(2, C, 16, 16), C)                                   # valid: no exception
    for bad in (torch.randn(2, C, 16), torch.randn(2, C, 16, 16, 1)):
        with pytest.raises(MoERouterError, match="4-D"):
            _validate_router_input(bad, C)
    with pytest.raises(ShapeMismatchError) as e:
        _validate_router_input(torch.randn(2, 32, 16, 16), C)
    assert e.value.expected == "(N, 64, H, W)" and e.value.actual == (2, 32, 16, 16) and "router input" in str(e.value)
    for v in (float("nan"), float("inf")):
        x = torch.randn(2, C, 16, 16)
        x[0, 0, 0, 0] = v
        with pytest.raises(MoERouterError, match="NaN"):
            _validate_router_input(x, C)
    assert issubclass(MoERouterError, YOLOMasterError) and issubclass(ShapeMismatchError, YOLOMasterError)
    router = EfficientSpatialRouter(C, 4, top_k=2).eval()
    with pytest.raises(MoERouterError, match="4-D"):
        router(torch.randn(2, C, 16))
    with pytest.raises(ShapeMismatchError):
        router(torch.randn(2, 128, 16, 16))
    for v in (float("nan"), float("inf")):
        router.noise_std = v
        with pytest.raises(MoERouterError, match="noise_std"):
            router(torch.randn(2, C, 16, 16))


def test_error_types_are_the_references_when_it_is_importable():
    """Inside the reference tree the package re-exports ultralytics' own classes (identity, so `except` clauses written against
    ultralytics.utils.errors keep working); standalone it defines equivalents."""
    import importlib
    import sys

    from yolo_master_b200.utils import errors
    if "ultralytics" in sys.modules or importlib.util.find_spec("ultralytics") is not None:
        from ultralytics.utils.errors import MoERouterError
        assert errors.MoERouterError is MoERouterError
    else:
        assert errors.MoERouterError.__module__ == "yolo_master_b200.utils.errors"
        e = errors.ShapeMismatchError("(N, 8, H, W)", (1, 4, 2, 2), "ctx")
        assert str(e) == "Shape mismatch: expected (N, 8, H, W), got (1, 4, 2, 2) [ctx]"
