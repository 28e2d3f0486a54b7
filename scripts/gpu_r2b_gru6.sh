#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 60 python scripts/time_gru_split.py 2>&1 | tail -1
timeout 400 python -m pytest -q --timeout=100 --timeout-method=thread -p no:cacheprovider tests/test_gpu_baseline_batch.py tests/test_gpu_forward.py -m gpu -k "fp32tc or ragged" -s > gpurun_out/tests_tc.log 2>&1
echo "tests rc=$?"; grep -E "parity at|passed|failed|Error|error" gpurun_out/tests_tc.log | tail -8
for d in 1 4; do VQA_B200_GRUS_DEBUG=$d timeout 60 python scripts/time_gru_split.py 2>&1 | tail -1; done
VQA_B200_GRU_SPLIT_PERSIST=0 timeout 60 python scripts/time_gru_split.py 2>&1 | tail -1
B=2048 timeout 60 python scripts/time_gru_split.py 2>&1 | tail -1
