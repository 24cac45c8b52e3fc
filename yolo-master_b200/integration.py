"""Drop-in binding INTO the reference: `install()` rebinds the reference's operator classes to accelerated subclasses.

The reference (an ultralytics fork) resolves every operator of a model YAML by NAME through module globals and the mixture registry
(`ultralytics/nn/tasks.py:2122-2130`, `nn/mixture_registry.py:39-81`), and its constructors resolve their children the same way through
the globals of `nn/modules/{conv,block,head}.py`, `nn/modules/moe/*.py`, ...  `install()` therefore

  1. derives, for each operator class `X` this package mirrors, `class X(reference.X)`: the reference's OWN constructor, parameters,
     buffers, `state_dict` layout, `fuse()`, training forward - everything - plus the methods of `yolo_master_b200.nn.modules.X`
     (weight packing, `fwd_nhwc`) under their own names, and a dispatching `forward` / `forward_fuse`:
         CUDA fp16 tensor, `.eval()`, not tracing / exporting   -> the sm_100a kernels (this package's forward)
         anything else (CPU tensors, training, fp32, ONNX export) -> `super().forward`, i.e. the reference's stock code
     so the constructor's CPU stride forward (`tasks.py:555-559`, `Detect.training` true), training, export and `model.fuse()` keep
     working unmodified.  The fallback is the REFERENCE's path, not a second implementation in this package: the product has no CPU
     arithmetic of its own (`tests/test_host.py::test_no_cpu_fallback` still holds for `yolo_master_b200.nn`).
  2. rebinds every global of every loaded `ultralytics.*` module (and every registry dict value) that IS the original class to the
     derived class, so `parse_model`, the registries and nested constructors all build the accelerated classes.

`uninstall()` restores every binding.  Nothing is imported from the reference until `install()` is called.

    import ultralytics                                   # the reference
    from yolo_master_b200 import integration
    integration.install()                                # before the model is built
    model = ultralytics.nn.tasks.DetectionModel("yolo26-master-n.yaml")   # reference class, reference YAML, accelerated operators
    model.load_state_dict(ckpt); model.eval().half().cuda()
    y = model(images_fp16_cuda)[0]                       # (B, 300, 6), through the reference's own _predict_once

Tested: tests/test_dropin_reference.py (CPU: swap, build, stride forward, fallback == stock reference bit for bit, accelerated
methods on the reference-built instances through the op emulation) and tests/test_gpu_dropin.py (GPU: predict-style inference).
"""
from __future__ import annotations

import sys

import torch
import torch.nn as nn

# operator classes of the detection hot path verified inside the reference by the tests above
DEFAULT_CLASSES = ("Conv", "DWConv", "Concat", "Bottleneck", "C2f", "C3", "C3k", "C3k2", "SPPF", "Attention", "PSABlock", "C2PSA",
                   "AAttn", "ABlock", "A2C2f", "EfficientSpatialRouter", "SimpleExpert", "OptimizedMOEImproved", "ABlockMoE",
                   "A2C2fMoE", "Detect")

_STATE = {"installed": None}


def _accelerated(mod: nn.Module, args) -> bool:
    """The dispatch rule of every derived class (module docstring)."""
    if mod.training or torch.jit.is_tracing() or torch.onnx.is_in_onnx_export():
        return False
    x = args[0] if args else None
    while isinstance(x, (list, tuple)) and x:
        x = x[0]
    return torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float16


def _graft_namespace(ym_cls, ref_cls):
    """Methods / class attributes of the mirror class (and of its non-torch bases) that the reference class does not define."""
    ns = {}
    for klass in reversed(ym_cls.__mro__):
        if klass in (object, nn.Module) or klass.__module__.startswith("torch"):
            continue
        for k, v in vars(klass).items():
            if k.startswith("__") and k.endswith("__"):
                continue
            ns[k] = v
    fwd = ns.pop("forward", None)
    ns.pop("forward_fuse", None)
    ns = {k: v for k, v in ns.items() if not hasattr(ref_cls, k)}      # the reference's own attributes always win
    return ns, fwd


def _derive(ref_cls, ym_cls):
    ns, ym_forward = _graft_namespace(ym_cls, ref_cls)
    if ym_forward is None:
        raise TypeError(f"{ym_cls.__name__}: the mirror class defines no forward")
    ref_forward = ref_cls.forward
    ref_forward_fuse = getattr(ref_cls, "forward_fuse", None)

    def forward(self, *args, **kw):
        if _accelerated(self, args):
            return ym_forward(self, *args, **kw)
        return ref_forward(self, *args, **kw)

    ns["forward"] = forward
    if ref_forward_fuse is not None:           # Conv: BaseModel.fuse() rebinds `m.forward = m.forward_fuse` (tasks.py:285-320)
        def forward_fuse(self, *args, **kw):
            if _accelerated(self, args):
                return ym_forward(self, *args, **kw)
            return ref_forward_fuse(self, *args, **kw)
        ns["forward_fuse"] = forward_fuse
    ns["_ym_forward"] = ym_forward
    ns["_ym_reference_class"] = ref_cls
    ns["__module__"] = ref_cls.__module__      # checkpoints pickle classes by module + qualname: they keep resolving either way
    ns["__qualname__"] = ref_cls.__qualname__
    ns["__doc__"] = ref_cls.__doc__
    return type(ref_cls.__name__, (ref_cls,), ns)


def _reference_class(name: str):
    """The class object the reference DEFINES under `name` (several modules re-export it)."""
    for modname, mod in list(sys.modules.items()):
        if mod is None or not modname.startswith("ultralytics"):
            continue
        c = vars(mod).get(name)
        if isinstance(c, type) and c.__module__.startswith("ultralytics") and not hasattr(c, "_ym_reference_class"):
            return getattr(sys.modules.get(c.__module__), name, c)
    return None


def install(classes=DEFAULT_CLASSES):
    """Rebind the reference's operator classes (module docstring).  Idempotent; returns {name: derived class}."""
    if _STATE["installed"] is not None:
        return _STATE["installed"]["derived"]
    import ultralytics.nn.mixture_registry  # noqa: F401  (the reference; must be importable)
    import ultralytics.nn.modules  # noqa: F401
    import ultralytics.nn.tasks  # noqa: F401

    from .nn import modules as YM

    derived, originals = {}, {}
    for name in classes:
        ref_cls, ym_cls = _reference_class(name), getattr(YM, name, None)
        if ref_cls is None or ym_cls is None:
            raise LookupError(f"integration.install: no class named {name} in {'the reference' if ref_cls is None else 'yolo_master_b200.nn.modules'}")
        derived[name] = _derive(ref_cls, ym_cls)
        originals[name] = ref_cls
    by_id = {id(originals[n]): derived[n] for n in derived}
    undo = []
    for modname, mod in list(sys.modules.items()):
        if mod is None or not modname.startswith("ultralytics"):
            continue
        for k, v in list(vars(mod).items()):
            if isinstance(v, type) and id(v) in by_id:
                undo.append((vars(mod), k, v))
                setattr(mod, k, by_id[id(v)])
            elif isinstance(v, dict) and not k.startswith("__"):         # registries: MIXTURE_MODULES and friends
                for kk, vv in list(v.items()):
                    if isinstance(vv, type) and id(vv) in by_id:
                        undo.append((v, kk, vv))
                        v[kk] = by_id[id(vv)]
            elif isinstance(v, (set, frozenset)) and any(isinstance(e, type) and id(e) in by_id for e in v):
                undo.append((vars(mod), k, v))                           # module-level class sets (MIXTURE_BASE_MODULES ...)
                setattr(mod, k, type(v)(by_id.get(id(e), e) if isinstance(e, type) else e for e in v))
    _STATE["installed"] = {"derived": derived, "originals": originals, "undo": undo}
    return derived


def uninstall():
    st = _STATE["installed"]
    if st is None:
        return
    for container, key, value in reversed(st["undo"]):
        container[key] = value
    _STATE["installed"] = None


def installed() -> bool:
    return _STATE["installed"] is not None
