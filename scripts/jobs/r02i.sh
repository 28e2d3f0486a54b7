#!/bin/bash
set -u
O=gpurun_out/r02i
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_baseline_batch.py tests/test_gpu_train.py tests/test_gpu_fallbacks.py -q -m gpu -x > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/status.txt
timeout 300 python scripts/timeline.py updown > $O/timeline_updown.txt 2>&1
timeout 600 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_100.json 2> $O/bench_100.err; echo "bench100 rc=$?" >> $O/status.txt
for g in 32x2 64x2; do
  VQA_B200_GRU_CFG=$g timeout 300 python bench.py --workloads updown --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-parity > $O/bench_updown_gru$g.json 2> $O/bench_updown_gru$g.err; echo "bench gru $g rc=$?" >> $O/status.txt
done
cat $O/status.txt; tail -5 $O/tests.log
