#!/bin/bash
# GPU call 15 of round 2: finer timing probe of the persistent conv kernel's epilogue.
mkdir -p gpurun_out
timeout 600 python tools/conv2_probe.py gpurun_out/conv2_probe_r02o.json > gpurun_out/conv2_probe_r02o.log 2>&1; cut -c1-150 gpurun_out/conv2_probe_r02o.log
