#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1 ONLY14=1
timeout 600 python -m pytest -q --timeout=300 --timeout-method=thread -p no:cacheprovider tests/test_gpu_ops.py -m gpu -k "gru" -s > gpurun_out/tests_gru.log 2>&1
echo "tests rc=$?"; grep -E "token table|passed|failed|Error|error" gpurun_out/tests_gru.log | tail -12
for d in 0 16; do VQA_B200_GRU_DEBUG=$d TABLE=1 timeout 120 python scripts/time_gru.py 2>&1 | tail -1; done
TABLE=0 timeout 120 python scripts/time_gru.py 2>&1 | tail -1
for c in 32x1 64x2 32x2; do VQA_B200_GRU_CFG=$c TABLE=1 timeout 120 python scripts/time_gru.py 2>&1 | tail -1; done
PRECISION=fp32tc timeout 300 python scripts/timeline.py updown > gpurun_out/timeline_fp32tc_updown.txt 2>&1; tail -45 gpurun_out/timeline_fp32tc_updown.txt
