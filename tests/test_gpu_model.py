"""GPU parity of the whole yolo26-master-n forward (stock YAML -> DetectionModel) against the CPU oracle and the
committed reference goldens, through the public API; plus CUDA-graph replay and host-buffer entry points."""
import os

import pytest
import torch

from _util import GOLD, close_stats, synth_sd_from_keys, yaml_n
from oracle import yolo_master_oracle as O
from yolo_master_b200.nn.tasks import DetectionModel
from yolo_master_b200.utils.synth import synth_images

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def model_and_sd():
    sd = synth_sd_from_keys(0)
    m = DetectionModel("yolo26-master-n.yaml")
    m.load_state_dict(sd)
    return m.to(DEV).eval(), sd


def _layers(m, x):
    feats = {}
    hooks = [mod.register_forward_hook(lambda mod, i, o, k=k: feats.__setitem__(k, o)) for k, mod in enumerate(m.model)]
    with torch.no_grad():
        y = m(x)
    for h in hooks:
        h.remove()
    torch.cuda.synchronize()
    return y[0], feats


def _check_dets(y, ref, score_tol=1e-2, box_tol=1.0):
    """Detections clear of the cut-off must be reproduced (same class, score within tol, box within tol pixels)."""
    y = y.float().cpu()
    n_checked = 0
    for b in range(y.shape[0]):
        kth = ref[b, -1, 4]
        for r in ref[b]:
            if r[4] < kth + 2 * score_tol:
                continue
            same = (y[b, :, 5] == r[5]) & ((y[b, :, 4] - r[4]).abs() < score_tol)
            assert same.any(), f"image {b}: reference detection {r.tolist()} not reproduced"
            assert (y[b][same][:, :4] - r[:4]).abs().max(1)[0].min() < box_tol
            n_checked += 1
    assert n_checked > 0.5 * ref.shape[0] * ref.shape[1]
    s1, s2 = y[..., 4].sort(dim=1, descending=True)[0], ref[..., 4].sort(dim=1, descending=True)[0]
    assert (s1 - s2).abs().max() < score_tol


@pytest.mark.parametrize("tag", ["b2_160", "b1_64"])
def test_model_matches_reference_golden(model_and_sd, tag):
    """CUDA path vs outputs of the real reference model (fixtures made by tests/golden/make_golden.py)."""
    m, _ = model_and_sd
    c = torch.load(os.path.join(GOLD, "yolo26-master-n.golden.pt"))["cases"][tag]
    x = synth_images(c["B"], c["H"], c["W"], c["seed"]).half().to(DEV)
    y, feats = _layers(m, x)
    for i, ref in c["layers"].items():
        mx, _ = close_stats(feats[i], ref)
        rms = float(ref.pow(2).mean().sqrt())
        mean_err = float((feats[i].float().cpu() - ref).abs().mean())
        assert mean_err < 4e-3 * rms and mx < 8e-2 * max(rms, 1.0), f"layer {i}: max {mx:.3e} mean {mean_err:.3e} rms {rms:.3f}"
    _check_dets(y, c["final"])
    for name, (w_ref, i_ref) in c["routes"].items():
        mod = dict(m.named_modules())[name.replace(".routing", "")]
        snap = mod.last_routing_snapshot
        assert torch.equal(snap["topk_indices"].cpu(), i_ref), name     # router top-k indices bit-exact
        torch.testing.assert_close(snap["topk_weights"].cpu(), w_ref, atol=5e-3, rtol=0)


def test_model_640_vs_oracle(model_and_sd):
    """BASELINE config geometry (640x640) at batch 2, checked per layer against the oracle run on the same inputs."""
    m, sd = model_and_sd
    x = synth_images(2, 640, 640, 3)
    y, feats = _layers(m, x.half().to(DEV))
    ref, ys = O.forward(O.parse_spec(yaml_n()), sd, x.half().float(), return_layers=True)
    for i in range(23):
        if feats[i] is None or ys[i] is None:
            continue
        a, b = feats[i].float().cpu(), ys[i]
        rms = float(b.pow(2).mean().sqrt())
        mean_err = float((a - b).abs().mean())
        assert mean_err < 4e-3 * rms, f"layer {i}: mean err {mean_err:.3e} rms {rms:.3f}"
    _check_dets(y, ref)


def test_graph_replay_and_host_api(model_and_sd):
    m, _ = model_and_sd
    x = synth_images(4, 320, 320, 5).half()
    with torch.no_grad():
        eager = m(x.to(DEV))[0].clone()
    g = m.graphed(4, 320, 320)
    assert g.kernels_per_replay > 50
    out = g(x.to(DEV)).clone()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)                     # same kernels, same order: bit-identical
    host = g.run_host(x.pin_memory())
    assert torch.equal(host, eager.cpu())
    x2 = synth_images(4, 320, 320, 6).half()
    assert not torch.equal(g.run_host(x2.pin_memory()), eager.cpu())


def test_batch_sizes_and_empty(model_and_sd):
    m, _ = model_and_sd
    with torch.no_grad():
        y1 = m(synth_images(1, 96, 128, 1).half().to(DEV))[0]
        y3 = m(synth_images(3, 96, 128, 1).half().to(DEV))[0]
    assert y1.shape == (1, 252, 6) and y3.shape == (3, 252, 6)
    assert torch.equal(y1[0], y3[0])                   # per-image independence (routing / GroupNorm are per sample)
