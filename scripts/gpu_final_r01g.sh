#!/bin/bash
# final verification of round 1 (state r01g = r01f + fused answer selection, LSTM / stacked encoders, base-cap predictor):
# full GPU suite, smoke, bench lines, Up-Down timeline and ncu launch list
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout -k 5 300 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout -k 5 60 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout -k 5 120 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_updown.json 2> gpurun_out/bench_updown.err; echo "bench rc=$?"
timeout -k 5 120 python bench.py --workload regat --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_regat.json 2> gpurun_out/bench_regat.err; echo "regat rc=$?"
timeout -k 5 60 python scripts/timeline.py updown 2>/dev/null | tail -11 > gpurun_out/timeline_updown.txt
timeout -k 5 60 python scripts/timeline.py regat 2>/dev/null | tail -13 > gpurun_out/timeline_regat.txt
timeout -k 5 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_updown.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_updown.log 2>&1
echo "ncu updown rc=$?"
