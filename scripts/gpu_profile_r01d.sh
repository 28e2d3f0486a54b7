#!/bin/bash
# r01d: bench lines + ncu launch lists (updown, regat) + ncu --set full of the hot kernels
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_updown.json 2> gpurun_out/bench_updown.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_updown.json
python bench.py --workload regat --steps 50 --warmup 5 > gpurun_out/bench_regat.json 2> gpurun_out/bench_regat.err; echo "regat rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 400 gpurun_out/bench_ref.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_updown.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_updown.log 2>&1
echo "ncu updown rc=$?"
python bench.py --workload regat --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_regat.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_regat.csv python bench.py --workload regat --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_regat.log 2>&1
echo "ncu regat rc=$?"
KERNELS="gru relation pool wv gat" bash scripts/gpu_profile.sh
