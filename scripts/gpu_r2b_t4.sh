#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest -q --timeout=300 --timeout-method=thread -p no:cacheprovider tests/test_gpu_baseline_batch.py -m gpu -k "fp32tc" -s > gpurun_out/tests_tc.log 2>&1
echo "tests rc=$?"; grep -E "parity at|passed|failed|Error|error|assert" gpurun_out/tests_tc.log | tail -30
timeout 300 python -m pytest -q --timeout=300 --timeout-method=thread -p no:cacheprovider tests/test_gpu_ops.py -m gpu -k "split" -s 2>&1 | grep -E "linear_split|passed|failed" | tail
PRECISION=fp32tc timeout 300 python scripts/timeline.py updown > gpurun_out/timeline_fp32tc_updown.txt 2>&1; tail -14 gpurun_out/timeline_fp32tc_updown.txt
VQA_B200_GRU_SPLIT_PERSIST=0 VQA_B200_SPLIT_SMALL_PAIR=0 PRECISION=fp32tc timeout 300 python scripts/timeline.py updown 2>&1 | tail -8
PRECISION=fp32tc timeout 300 python scripts/timeline.py regat > gpurun_out/timeline_fp32tc_regat.txt 2>&1; tail -12 gpurun_out/timeline_fp32tc_regat.txt
