"""The drop-in on hardware: the reference's own DetectionModel / YOLO(...).predict(...) built out of the derived operator classes
(`yolo_master_b200.integration.install()`), launching this package's kernels - see tests/dropin_check.py (mode gpu)."""
import os
import subprocess
import sys

import pytest

from _util import ROOT

pytestmark = pytest.mark.gpu


def test_reference_predict_runs_on_the_cuda_kernels():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ultralytics", "nn", "tasks.py")):
        pytest.skip("oracle/_ref/ultralytics not shipped (make -C oracle in the build container)")
    env = dict(os.environ, YOLO_CONFIG_DIR="/tmp/ulcfg", YOLO_VERBOSE="false", YOLO_OFFLINE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_check.py"), "gpu"], capture_output=True, text=True, env=env, timeout=1500)
    assert r.returncode == 0 and "DROPIN OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    print(r.stdout.strip().splitlines()[-1])
