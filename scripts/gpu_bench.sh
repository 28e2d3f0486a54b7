#!/bin/bash
# tests + bench + ncu launch list (run under gpurun, 1 GPU)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q --timeout=300 --timeout-method=thread -p no:cacheprovider tests -m gpu > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_updown.json 2> gpurun_out/bench_updown.err; echo "bench rc=$?"; cat gpurun_out/bench_updown.json; tail -3 gpurun_out/bench_updown.err
python bench.py --workload regat --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_regat.json 2> gpurun_out/bench_regat.err; echo "regat rc=$?"; cat gpurun_out/bench_regat.json; tail -3 gpurun_out/bench_regat.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_updown.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_updown.log 2>&1
echo "ncu rc=$?"
