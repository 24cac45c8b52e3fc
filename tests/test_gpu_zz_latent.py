"""GPU parity of LatentMixture (the yolo26-master-latent-n* zoo): `ym_latent_router` and the whole model against the
reference golden and the CPU oracle.  The kernel also runs on the CUDA-on-host emulation and the whole model on
emulated ops in the CPU suite; on the B200 since round 2 (profiles/r02_gpu_suite.txt)."""
import os

import pytest
import torch

from _util import GOLD, assert_within_noise, synth_sd_from_keys, yaml_of
from oracle import yolo_master_oracle as O
from yolo_master_b200.nn.tasks import DetectionModel
from yolo_master_b200.utils.synth import synth_images

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAME, CFG = "yolo26-master-latent-n", "26/yolo26-master-latent-n-resinit010.yaml"


def test_latent_model_matches_reference_golden():
    m = DetectionModel(CFG)
    sd = synth_sd_from_keys(0, NAME)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    c = torch.load(os.path.join(GOLD, f"{NAME}.golden.pt"))["cases"]["b2_128"]
    x = synth_images(c["B"], c["H"], c["W"], c["seed"]).half()
    feats = {}
    hooks = [mod.register_forward_hook(lambda mod, i, o, k=k: feats.__setitem__(k, o)) for k, mod in enumerate(m.model)]
    with torch.no_grad():
        y = m(x.to(DEV))[0].float().cpu()
    for h in hooks:
        h.remove()
    spec = O.parse_spec(yaml_of(CFG))
    ref, ys = O.forward(spec, sd, x.float(), return_layers=True)
    with O.fp16_storage(), O.fp16_weights():
        _, sim = O.forward(spec, sd, x.float(), return_layers=True)
    for i in (23, 24, 25):
        assert_within_noise(feats[i], c["layers"][i], sim[i], what=f"LatentMixture layer {i} vs reference golden", outlier_frac=0.02)
        assert_within_noise(feats[i], ys[i], sim[i], what=f"LatentMixture layer {i}", outlier_frac=0.02)
    assert y.shape == ref.shape == (2, 300, 6)
    torch.testing.assert_close(y[:, :20, 4], ref[:, :20, 4], atol=2e-2, rtol=5e-2)
    out = m.graphed(2, 128, 128)(x.to(DEV)).clone()
    torch.cuda.synchronize()
    assert torch.equal(out.float().cpu(), y)
