#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q --timeout=600 --timeout-method=thread -p no:cacheprovider tests/test_gpu_baseline_batch.py -m gpu -k "fp32tc or bf16_matches" -s > gpurun_out/tests_tc.log 2>&1
echo "tests rc=$?"; grep -E "parity at|passed|failed|Error|error|assert" gpurun_out/tests_tc.log | tail -30
PRECISION=fp32tc timeout 300 python scripts/timeline.py updown > gpurun_out/timeline_fp32tc_updown.txt 2>&1; tail -28 gpurun_out/timeline_fp32tc_updown.txt
timeout 300 python scripts/timeline.py updown > gpurun_out/timeline_updown.txt 2>&1; tail -12 gpurun_out/timeline_updown.txt
