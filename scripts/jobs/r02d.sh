#!/bin/bash
# round-2 GPU job D: graph attention with an 8-warp KxK stage
set -u
O=gpurun_out/r02d
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_baseline_batch.py tests/test_gpu_modules.py -q -m gpu -x > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/status.txt
for c in 0 16 24; do
  VQA_B200_GAT_CHASE=$c timeout 300 python scripts/timeline.py regat > $O/timeline_regat_chase$c.txt 2>&1; echo "timeline chase $c rc=$?" >> $O/status.txt
done
for c in 0 8 12 16 20 24 32; do
  timeout 300 python bench.py --workloads regat --steps 100 --warmup 5 --chase $c --no-e2e --no-cpu-baseline --no-parity > $O/bench_regat_chase$c.json 2> $O/bench_regat_chase$c.err; echo "bench chase $c rc=$?" >> $O/status.txt
done
cat $O/status.txt
