#!/bin/bash
# final verification of round 1 (state r01f): full GPU suite, smoke, bench lines, timelines, launch lists, pool ncu
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout -k 5 400 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout -k 5 60 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout -k 5 200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_updown.json 2> gpurun_out/bench_updown.err; echo "bench rc=$?"
timeout -k 5 200 python bench.py --workload regat --steps 50 --warmup 5 > gpurun_out/bench_regat.json 2> gpurun_out/bench_regat.err; echo "regat rc=$?"
timeout -k 5 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
B=512 timeout -k 5 100 python scripts/time_qcap.py 2>/dev/null | tail -1 > gpurun_out/bench_qcap.json
timeout -k 5 100 python scripts/train_bench.py 2>/dev/null | tail -1 > gpurun_out/bench_train.json
timeout -k 5 100 python scripts/train_bench.py --torch-optim 2>/dev/null | tail -1 > gpurun_out/bench_train_torch_optim.json
B=128 timeout -k 5 100 python scripts/time_decoder.py 2>/dev/null | tail -1 > gpurun_out/bench_decoder_b128.json
B=512 timeout -k 5 100 python scripts/time_decoder.py 2>/dev/null | tail -1 > gpurun_out/bench_decoder_b512.json
timeout -k 5 60 python scripts/timeline.py updown 2>/dev/null | tail -12 > gpurun_out/timeline_updown.txt
timeout -k 5 60 python scripts/timeline.py regat 2>/dev/null | tail -14 > gpurun_out/timeline_regat.txt
timeout -k 5 100 python scripts/time_ops.py > gpurun_out/time_ops.log 2>&1
timeout -k 5 100 python bench.py --workload regat --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_regat.log 2>&1 &&
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_regat.csv python bench.py --workload regat --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_regat.log 2>&1
echo "ncu regat rc=$?"
timeout -k 5 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_updown.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_updown.log 2>&1
echo "ncu updown rc=$?"
KERNELS="pool" bash scripts/gpu_profile.sh
