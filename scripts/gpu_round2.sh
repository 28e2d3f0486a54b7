#!/bin/bash
# relation kernel tests + op timings + GRU load-phase experiments (run under gpurun, 1 GPU)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest -q --timeout=300 --timeout-method=thread -p no:cacheprovider tests/test_gpu_ops.py tests/test_gpu_modules.py -m gpu -k "relation" > gpurun_out/tests_rel.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests_rel.log
timeout 600 python scripts/time_ops.py > gpurun_out/time_ops.log 2>&1; tail -32 gpurun_out/time_ops.log
for d in 0 8 16 24 25; do VQA_B200_GRU_DEBUG=$d timeout 120 python scripts/time_gru.py 2>&1 | tail -1; done | tee gpurun_out/time_gru3.log
