#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 300 python bench.py --workloads regat --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-exact-block > gpurun_out/bench_regat.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_regat.json')); print('regat', round(d['value']), d['ms_per_step'])"
for pm in 250 280 340; do timeout 300 python bench.py --workloads regat --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-exact-block --no-parity --side-permille $pm > gpurun_out/bench_regat_$pm.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_regat_$pm.json')); print('permille $pm', round(d['value']), d['ms_per_step'])"; done
VQA_B200_GRU_TABLE=0 timeout 300 python bench.py --workloads regat --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-exact-block --no-parity > gpurun_out/bench_regat_notab.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_regat_notab.json')); print('no table', round(d['value']), d['ms_per_step'])"
timeout 200 python scripts/timeline.py regat > gpurun_out/timeline_regat.txt 2>&1; tail -13 gpurun_out/timeline_regat.txt
