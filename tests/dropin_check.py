"""Drop-in check, run in its OWN process by tests/test_dropin_reference.py (CPU) and tests/test_gpu_dropin.py (GPU), so that the
reference package (`oracle/_ref/ultralytics`) is never imported into the pytest process.

    python tests/dropin_check.py cpu|gpu

Performs INTEGRATION.md's swap (`yolo_master_b200.integration.install()`) INSIDE the unmodified reference and builds the REFERENCE's
`ultralytics.nn.tasks.DetectionModel('yolo26-master-n.yaml')` (nn/tasks.py:530-577) out of the derived operator classes.

cpu: * the constructor's stride forward (tasks.py:555-559: CPU, 256x256, Detect in training mode) runs through the fallback;
     * CPU / fp32 / training inputs give bit-identical results to the stock reference (the fallback IS the reference's code);
     * the accelerated methods run on the reference-built instances (plain nn.Sequential / nn.Conv2d children, reference constructor
       attributes) through the torch emulation of the C-ABI ops (tools/cpu_emu.py) and land on the reference's fp32 output within the
       fp16 noise floor, before and after the reference's own `model.fuse()`;
     * `uninstall()` restores every binding.
gpu: * the same model `.eval().half().cuda()`: the reference's `_predict_once` now launches this package's kernels (counted), result
       within the fp16 noise floor of this package's own DetectionModel and of the stock reference on the same GPU;
     * predict-style inference: checkpoint saved with `torch.save`, loaded and run by the reference's own `YOLO(...).predict(...)`
       (AutoBackend -> fuse -> half -> LetterBox -> model -> Results), detections equal to the stock reference's within noise.
Prints one line `DROPIN OK {...}` on success.
"""
import importlib.util
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from _util import synth_sd_from_keys  # noqa: E402
from oracle import reference_runner as R  # noqa: E402
from yolo_master_b200.utils.synth import synth_images  # noqa: E402


def top(t):
    return t[..., 4].float().sort(1, descending=True)[0].cpu()


def main(mode):
    R.import_reference()
    import ultralytics.nn.tasks as T
    from yolo_master_b200 import integration

    sd = synth_sd_from_keys(0)
    cfg = R.reference_yaml()
    x = synth_images(2, 160, 160, 1)
    stock = T.DetectionModel(cfg, verbose=False)
    stock.load_state_dict(sd)
    stock.eval()
    with torch.no_grad():
        y_stock = stock(x)[0]
    stock_conv = T.Conv
    info = {}

    if mode == "cpu":
        spec = importlib.util.spec_from_file_location("cpu_emu", os.path.join(ROOT, "tools", "cpu_emu.py"))
        emu = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(emu)
        emu.install_model()                      # torch restatement of the C-ABI ops, BEFORE the derived classes capture the forwards

    derived = integration.install()
    assert T.Conv is derived["Conv"] and issubclass(T.Conv, stock_conv) and T.Conv.__name__ == "Conv"
    import ultralytics.nn.modules.block as RB
    assert RB.Conv is derived["Conv"], "nested constructors must resolve the derived class (module globals of block.py)"
    m = T.DetectionModel(cfg, verbose=False)      # runs the CPU stride forward with Detect.training = True
    assert m.stride.tolist() == [8.0, 16.0, 32.0], m.stride
    assert type(m.model[0]) is derived["Conv"] and type(m.model[4]) is derived["A2C2fMoE"] and type(m.model[-1]) is derived["Detect"]
    assert list(m.state_dict().keys()) == list(stock.state_dict().keys())
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        y_fb = m(x)[0]
    assert torch.equal(y_fb, y_stock), "CPU fp32 eval through the derived classes must be the stock reference bit for bit"
    m.train()
    out_train = m(x)
    assert isinstance(out_train, dict) and "one2one" in out_train, "training forward falls back to the reference"
    m.load_state_dict(sd, strict=True)            # the training forward moved the BatchNorm running statistics
    m.eval()

    if mode == "cpu":
        real_rule = integration._accelerated
        integration._accelerated = lambda mod, args: not mod.training       # CPU tensors take the accelerated methods (emulated ops)
        try:
            with torch.no_grad():
                y_acc = m(x.half())[0]
                noise = float((top(y_acc) - top(y_stock)).abs().max())
                assert noise < 2.5e-2, noise
                m.fuse(verbose=False)                                        # the reference's own Conv+BN fusion, Detect.fuse()
                assert not hasattr(m.model[0], "bn") and m.model[-1].cv2 is None
                y_fused = m(x.half())[0]
                d_fuse = float((top(y_fused) - top(y_acc)).abs().max())
                assert d_fuse < 2.5e-2, d_fuse
        finally:
            integration._accelerated = real_rule
        info.update(emulated_vs_reference=noise, after_reference_fuse=d_fuse)
    else:
        from yolo_master_b200 import ops
        from yolo_master_b200.nn.tasks import DetectionModel as OursDM
        dev = torch.device("cuda", 0)
        xg = synth_images(4, 640, 640, 3).to(dev)
        stock_g = stock.to(dev)
        with torch.no_grad():
            y_ref = stock_g(xg)[0]                                           # stock reference, fp32, same GPU
        m = m.half().to(dev)
        k0 = ops.KERNELS
        with torch.no_grad():
            y_acc = m(xg.half())[0]
        torch.cuda.synchronize()
        launched = ops.KERNELS - k0
        assert launched > 100, f"the reference's _predict_once did not reach the CUDA kernels ({launched} launches)"
        ours = OursDM("yolo26-master-n.yaml")
        ours.load_state_dict(sd)
        ours.to(dev).eval()
        with torch.no_grad():
            y_ours = ours(xg.half())[0]
        d_ours = float((top(y_acc) - top(y_ours)).abs().max())
        d_ref = float((top(y_acc) - top(y_ref)).abs().max())
        # not bit-identical: `model.half()` rounds the reference's parameters to fp16 BEFORE this package folds BatchNorm (its own
        # DetectionModel folds the fp32 parameters), and nn.Upsample stays torch's kernel - both inside the fp16 noise floor
        assert d_ours < 2.5e-2, f"reference-built model vs this package's DetectionModel: {d_ours}"
        assert d_ref < 2.5e-2, f"vs the stock reference on the same GPU: {d_ref}"
        info.update(kernels=launched, vs_own_model=d_ours, vs_stock_reference_gpu=d_ref)

        # ---- predict-style: the reference's own YOLO(...).predict(...) on a checkpoint of the derived-class model
        import numpy as np
        from ultralytics import YOLO
        with tempfile.TemporaryDirectory() as td:
            pt = os.path.join(td, "yolo26-master-n-synth.pt")
            ck = T.DetectionModel(cfg, verbose=False)
            ck.load_state_dict(sd)
            ck.names = {i: f"c{i}" for i in range(80)}
            ck.args = {"imgsz": 640}
            torch.save({"model": ck.half(), "train_args": {}, "date": None, "version": "drop-in"}, pt)
            g = np.random.default_rng(0)
            frames = [g.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(2)]
            k1 = ops.KERNELS
            res_acc = YOLO(pt).predict(frames, half=True, device=0, conf=0.05, verbose=False, imgsz=640)
            launched_pred = ops.KERNELS - k1
            assert launched_pred > 100, launched_pred
            integration.uninstall()
            res_ref = YOLO(pt).predict(frames, half=False, device=0, conf=0.05, verbose=False, imgsz=640)
            integration.install()
            for a, b in zip(res_acc, res_ref):
                ca, cb = a.boxes.conf.float().cpu().sort(descending=True)[0], b.boxes.conf.float().cpu().sort(descending=True)[0]
                n = min(len(ca), len(cb), 50)
                assert n > 0 and abs(len(ca) - len(cb)) <= max(3, len(cb) // 10), (len(ca), len(cb))
                assert float((ca[:n] - cb[:n]).abs().max()) < 2.5e-2
            info.update(predict_kernels=launched_pred, predict_dets=[len(r.boxes) for r in res_acc])

    integration.uninstall()
    assert T.Conv is stock_conv and not integration.installed()
    print("DROPIN OK " + json.dumps(info))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cpu")
