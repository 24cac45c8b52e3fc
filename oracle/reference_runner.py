"""Loader for the UNMODIFIED reference (Tencent/YOLO-Master `ultralytics/` tree) as shipped in `oracle/_ref/`.

TEST INFRASTRUCTURE ONLY - imported by tests/, `bench.py --impl reference`, bench.py's `torch_eager_gpu` / `cpu_baseline` legs and
nothing else.  The product package never imports this module (tests/test_host.py::test_product_does_not_import_oracle).

`oracle/_ref/ultralytics` is produced by `make -C oracle` (oracle/Makefile) from `/root/reference/ultralytics` in the build container;
it is git-ignored and travels to the GPU box with the gpurun snapshot, where `/root/reference` does not exist.  Nothing here reads
`/root/reference` at run time.

The model this returns is the reference's own `ultralytics.nn.tasks.DetectionModel` (nn/tasks.py:530-577) built from the
reference's own YAML, running the reference's own `_predict_once` (nn/tasks.py:182-218): stock code path, none of this
repository's kernels, modules or engine on it.  Only the WEIGHTS come from this repository: the key-seeded synthetic
state_dict (`utils/synth.py`, a pure function of key names) that every parity fixture uses.
"""
from __future__ import annotations

import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(_HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_ROOT, "ultralytics", "nn", "tasks.py"))


def import_reference():
    """Put oracle/_ref first on sys.path and import the reference package (settings dir redirected to /tmp)."""
    if not available():
        raise RuntimeError("oracle/_ref/ultralytics is missing: run `make -C oracle` in the build container (needs /root/reference)")
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ulcfg")
    os.environ.setdefault("YOLO_VERBOSE", "false")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import ultralytics  # noqa: F401

    assert os.path.realpath(ultralytics.__file__).startswith(os.path.realpath(REF_ROOT)), \
        f"another ultralytics shadows the reference: {ultralytics.__file__}"
    return ultralytics


def reference_yaml(rel: str = "26/yolo26-master-n.yaml") -> str:
    return os.path.join(REF_ROOT, "ultralytics", "cfg", "models", rel)


def build_reference_model(state_dict, rel: str = "26/yolo26-master-n.yaml", fuse: bool = True):
    """`DetectionModel(yaml).load_state_dict(sd).eval().fuse()` of the real reference (what `AutoBackend` runs at predict time,
    nn/autobackend.py: `model.fuse()`)."""
    import_reference()
    from ultralytics.nn.tasks import DetectionModel

    m = DetectionModel(reference_yaml(rel), verbose=False)
    m.load_state_dict(state_dict, strict=True)
    m.eval()
    if fuse:
        m.fuse(verbose=False)
    for p in m.parameters():
        p.requires_grad_(False)
    return m
