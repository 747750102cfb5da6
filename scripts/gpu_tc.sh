#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_fused.py -q -m gpu -x -s --timeout=120 > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
grep -E "^\[tc\]|passed|failed|Error|error|assert" gpurun_out/pytest_tc.log | head -40
