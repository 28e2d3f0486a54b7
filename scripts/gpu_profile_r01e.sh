#!/bin/bash
# r01e: full GPU test suite, smoke, bench lines, ncu launch lists (updown, regat), ncu --set full of the hot kernels
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q --timeout=300 -p no:cacheprovider tests -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_updown.json 2> gpurun_out/bench_updown.err; echo "bench rc=$?"
python bench.py --workload regat --steps 50 --warmup 5 > gpurun_out/bench_regat.json 2> gpurun_out/bench_regat.err; echo "regat rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
B=512 python scripts/time_qcap.py 2>/dev/null | tail -1 > gpurun_out/bench_qcap.json
python scripts/train_bench.py 2>/dev/null | tail -1 > gpurun_out/bench_train.json
python scripts/timeline.py updown 2>/dev/null | tail -12 > gpurun_out/timeline_updown.txt
python scripts/timeline.py regat 2>/dev/null | tail -14 > gpurun_out/timeline_regat.txt
python scripts/timeline_train.py 2>/dev/null | tail -24 > gpurun_out/timeline_train.txt
python scripts/time_ops.py > gpurun_out/time_ops.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_updown.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_updown.log 2>&1
echo "ncu updown rc=$?"
python bench.py --workload regat --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_regat.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_regat.csv python bench.py --workload regat --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_regat.log 2>&1
echo "ncu regat rc=$?"
KERNELS="gru wide wv relation pool" bash scripts/gpu_profile.sh
