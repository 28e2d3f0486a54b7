#!/bin/bash
# token-table GRU: parity tests, timing with / without the table, short bench
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest -q --timeout=300 --timeout-method=thread -p no:cacheprovider tests/test_gpu_ops.py -m gpu -k "gru" -s > gpurun_out/tests_gru.log 2>&1
echo "tests rc=$?"; grep -E "token table|passed|failed|Error|error" gpurun_out/tests_gru.log | tail -20
TABLE=1 timeout 120 python scripts/time_gru.py 2>&1 | tail -3
TABLE=0 timeout 120 python scripts/time_gru.py 2>&1 | tail -3
for d in 1 4 8; do VQA_B200_GRU_DEBUG=$d TABLE=1 timeout 120 python scripts/time_gru.py 2>&1 | tail -1; done
timeout 120 python scripts/probes/tc_accum_probe.py 2>&1 | tail -8
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_tab.json 2> gpurun_out/bench_tab.err; echo "bench rc=$?"; cat gpurun_out/bench_tab.json; tail -3 gpurun_out/bench_tab.err
