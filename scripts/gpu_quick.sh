#!/bin/bash
# quick GPU iteration: selected tests + bench (1 GPU)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q --timeout=300 --timeout-method=thread -p no:cacheprovider tests -m gpu ${TEST_ARGS} > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; tail -${TAIL:-8} gpurun_out/tests.log
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_updown.json 2> gpurun_out/bench_updown.err; echo "bench rc=$?"; cat gpurun_out/bench_updown.json; tail -3 gpurun_out/bench_updown.err
python bench.py --workload regat --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_regat.json 2> gpurun_out/bench_regat.err; echo "regat rc=$?"; cat gpurun_out/bench_regat.json; tail -3 gpurun_out/bench_regat.err
