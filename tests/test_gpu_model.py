"""GPU parity of the whole yolo26-master-n forward (stock YAML -> DetectionModel) against the CPU oracle and the
committed reference goldens, through the public API; plus CUDA-graph replay and host-buffer entry points."""
import os

import pytest
import torch

from _util import GOLD, assert_within_noise, close_stats, synth_sd_from_keys, yaml_n
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


def _repro_frac(y, ref, score_tol=1e-2, box_tol=1.0):
    """Fraction of reference detections clear of the score cut-off that `y` reproduces (same class, score and box)."""
    y = y.float().cpu()
    hit = tot = 0
    for b in range(y.shape[0]):
        kth = ref[b, -1, 4]
        for r in ref[b]:
            if r[4] < kth + 2 * score_tol:
                continue
            tot += 1
            same = (y[b, :, 5] == r[5]) & ((y[b, :, 4] - r[4]).abs() < score_tol)
            if same.any() and (y[b][same][:, :4] - r[:4]).abs().max(1)[0].min() < box_tol:
                hit += 1
    return hit / max(tot, 1), tot


def _check_dets(y, ref, sim):
    """Detections must agree with the fp32 oracle at least as well as the oracle's own fp16-storage model does."""
    f_ours, tot = _repro_frac(y, ref)
    f_sim, _ = _repro_frac(sim, ref)
    assert tot > 0.3 * ref.shape[0] * ref.shape[1]
    assert f_ours >= min(f_sim, 0.995) - 0.05, f"reproduced {f_ours:.3f} of {tot} detections (fp16 noise floor {f_sim:.3f})"
    s1, s2 = y.float().cpu()[..., 4].sort(dim=1, descending=True)[0], ref[..., 4].sort(dim=1, descending=True)[0]
    s3 = sim[..., 4].sort(dim=1, descending=True)[0]
    assert (s1 - s2).abs().max() <= 3 * (s3 - s2).abs().max() + 2e-3


@pytest.mark.parametrize("tag", ["b2_160", "b1_64"])
def test_model_matches_reference_golden(model_and_sd, tag):
    """CUDA path vs outputs of the real reference model (fixtures made by tests/golden/make_golden.py)."""
    m, _ = model_and_sd
    c = torch.load(os.path.join(GOLD, "yolo26-master-n.golden.pt"))["cases"][tag]
    x = synth_images(c["B"], c["H"], c["W"], c["seed"]).half().to(DEV)
    y, feats = _layers(m, x)
    with O.fp16_storage(), O.fp16_weights():   # noise floor of fp16 storage for this graph, from the oracle itself
        ysim, sim = O.forward(O.parse_spec(yaml_n()), model_and_sd[1], x.float().cpu(), return_layers=True)
    for i, ref in c["layers"].items():
        assert_within_noise(feats[i], ref, sim[i], what=f"layer {i} vs reference golden")
    _check_dets(y, c["final"], ysim)
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
    spec = O.parse_spec(yaml_n())
    ref, ys = O.forward(spec, sd, x.half().float(), return_layers=True)
    with O.fp16_storage(), O.fp16_weights():
        ysim, sim = O.forward(spec, sd, x.half().float(), return_layers=True)
    for i in range(23):
        if feats.get(i) is None:
            continue
        assert_within_noise(feats[i], ys[i], sim[i], what=f"layer {i} (end to end)")
    _check_dets(y, ref, ysim)


def test_layers_teacher_forced_640(model_and_sd):
    """Every top-level layer at the 640x640 geometry, fed the ORACLE's (fp16-rounded) input, against the oracle's output
    for that same input: per-layer parity without error accumulation from earlier layers."""
    m, sd = model_and_sd
    spec = O.parse_spec(yaml_n())
    x = synth_images(2, 640, 640, 4).half().float()
    _, ys = O.forward(spec, sd, x, return_layers=True)
    ys16 = {k: (v.half() if torch.is_tensor(v) else v) for k, v in ys.items()}
    for i, L in enumerate(spec["layers"][:-1]):
        f = L["f"]
        src = (lambda j: x.half() if (i == 0 and j == -1) else ys16[i - 1 if j == -1 else j])
        xin = src(f) if isinstance(f, int) else [src(j) for j in f]
        to_dev = lambda t: t.to(DEV).contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            y = m.model[i](to_dev(xin) if torch.is_tensor(xin) else [to_dev(t) for t in xin])
        xin32 = xin.float() if torch.is_tensor(xin) else [t.float() for t in xin]
        ref = O.forward_layer(spec, sd, i, xin32)
        with O.fp16_storage(), O.fp16_weights():
            sim = O.forward_layer(spec, sd, i, xin32)
        if L["type"] in ("Concat", "nn.Upsample"):
            assert torch.equal(y.float().cpu(), ref), f"layer {i} {L['type']}"
        else:
            assert_within_noise(y, ref, sim, what=f"layer {i} {L['type']} (teacher forced)")
            if L["type"] == "Conv":   # a single fused kernel: strict north-star tolerance vs the reference's fp16 weights
                with O.fp16_weights():
                    ref16 = O.forward_layer(spec, sd, i, xin32)
                mx, bad = close_stats(y, ref16)
                assert bad == 0.0, f"layer {i} Conv: {bad:.2e} outside tolerance (max {mx:.2e})"


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


def test_stream_host_pipeline_uint8(model_and_sd):
    """Pipelined host API (double-buffered H2D / compute / D2H) on uint8 frames returns, in order, exactly what the
    synchronous call returns for each batch; uint8 input == the same frames as fp16 /255 up to the fp16 rounding of x/255."""
    m, _ = model_and_sd
    frames = [(synth_images(2, 256, 256, 40 + i) * 255).round().to(torch.uint8).pin_memory() for i in range(5)]
    g8 = m.graphed(2, 256, 256, dtype=torch.uint8)
    want = [g8.run_host(f).clone() for f in frames]
    got = [o.clone() for o in g8.stream_host(iter(frames))]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    with torch.no_grad():
        y16 = m((frames[0].float() / 255).half().to(DEV))[0].float().cpu()
    s8, s16 = want[0][..., 4].sort(dim=1, descending=True)[0], y16[..., 4].sort(dim=1, descending=True)[0]
    assert (s8 - s16).abs().max() < 2e-2


def test_pipelined_forward_two_instances_bit_identical_and_in_order(model_and_sd):
    """PipelinedForward (two graph instances on two streams): every batch gets exactly what the single graph returns for it, in order,
    for the device-resident call and for the host-buffer stream with an odd number of batches (slot reuse, flush)."""
    m, _ = model_and_sd
    frames = [(synth_images(2, 256, 256, 60 + i) * 255).round().to(torch.uint8).pin_memory() for i in range(7)]
    g8 = m.graphed(2, 256, 256, dtype=torch.uint8)
    want = [g8.run_host(f).clone() for f in frames]
    pipe = m.pipelined(2, 256, 256, dtype=torch.uint8, depth=2)
    assert pipe.kernels_per_replay == g8.kernels_per_replay and pipe.depth == 2
    got = [o.clone() for o in pipe.stream_host(iter(frames))]
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        assert torch.equal(a, b), f"batch {i}"
    got2 = [o.clone() for o in pipe.stream_host(iter(frames[:3]))]           # second use of the same pipeline, fewer batches
    assert all(torch.equal(a, b) for a, b in zip(got2, want[:3]))
    outs = pipe.run_device([f.to(DEV) for f in frames[:2]])
    torch.cuda.synchronize()
    assert torch.equal(outs[0].cpu(), want[0]) and torch.equal(outs[1].cpu(), want[1])


def test_empty_batch_and_bad_rank(model_and_sd):
    m, _ = model_and_sd
    with torch.no_grad():
        y, aux = m(torch.zeros((0, 3, 96, 128), dtype=torch.float16, device=DEV))
    assert y.shape == (0, 252, 6) and aux["boxes"] == []
    with pytest.raises(ValueError):
        m(torch.zeros((3, 96, 128), dtype=torch.float16, device=DEV))
    with pytest.raises(RuntimeError):
        m(torch.zeros((1, 3, 96, 128), dtype=torch.float16))      # CPU tensor: no fallback
