#!/bin/bash
set -u
O=gpurun_out/r02h
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_baseline_batch.py tests/test_gpu_modules.py -q -m gpu -x > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/status.txt
for d in 0 2 4 8; do VQA_B200_GAT_DEBUG=$d timeout 120 python scripts/gat_probe.py >> $O/gat_probe.txt 2>&1; done
timeout 300 python scripts/timeline.py regat > $O/timeline_regat.txt 2>&1
timeout 600 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_100.json 2> $O/bench_100.err; echo "bench100 rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_20.json 2> $O/bench_20.err; echo "bench20 rc=$?" >> $O/status.txt
cat $O/status.txt; cat $O/gat_probe.txt; tail -3 $O/tests.log
