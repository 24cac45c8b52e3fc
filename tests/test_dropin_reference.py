"""INTEGRATION.md's class swap performed INSIDE the unmodified reference (oracle/_ref/ultralytics), on the CPU: see tests/dropin_check.py.
Runs in a subprocess so that this pytest process never imports the reference package."""
import os
import subprocess
import sys

import pytest

from _util import ROOT


def test_reference_builds_and_runs_on_the_derived_classes():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ultralytics", "nn", "tasks.py")):
        pytest.skip("oracle/_ref/ultralytics not built (make -C oracle needs /root/reference)")
    env = dict(os.environ, YOLO_CONFIG_DIR="/tmp/ulcfg", YOLO_VERBOSE="false")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_check.py"), "cpu"], capture_output=True, text=True, env=env, timeout=1500)
    assert r.returncode == 0 and "DROPIN OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
