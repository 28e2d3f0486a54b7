#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest -q --timeout=600 --timeout-method=thread -p no:cacheprovider tests/test_gpu_forward.py tests/test_gpu_ops.py tests/test_gpu_fallbacks.py tests/test_gpu_baseline_batch.py -m gpu -k "regat_full_batch or token_table or single_cta or fp32tc or gru_tile" > gpurun_out/tests_fix.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/tests_fix.log
for b in 128 512; do B=$b ONLY14=1 TABLE=1 timeout 100 python scripts/time_gru.py | tail -1; B=$b ONLY14=1 TABLE=0 timeout 100 python scripts/time_gru.py | tail -1; done
