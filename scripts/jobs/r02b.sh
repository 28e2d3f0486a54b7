#!/bin/bash
# round-2 GPU job B: chase schedule, GRU tile configurations, full suite
set -u
O=gpurun_out/r02b
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_baseline_batch.py -q -m gpu -s > $O/tests_new.log 2>&1; echo "tests_new rc=$?" >> $O/status.txt
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_baseline_batch.py > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/status.txt
for c in 0 8 12 16 24 32; do
  VQA_B200_GAT_CHASE=$c timeout 300 python scripts/timeline.py regat > $O/timeline_regat_chase$c.txt 2>&1; echo "timeline chase $c rc=$?" >> $O/status.txt
  timeout 300 python bench.py --workloads regat --steps 100 --warmup 5 --chase $c --no-e2e --no-cpu-baseline --no-parity > $O/bench_regat_chase$c.json 2> $O/bench_regat_chase$c.err; echo "bench chase $c rc=$?" >> $O/status.txt
done
for g in 64x1 32x1 64x2 32x2; do
  VQA_B200_GRU_CFG=$g timeout 300 python bench.py --workloads updown --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-parity > $O/bench_updown_gru$g.json 2> $O/bench_updown_gru$g.err; echo "bench gru $g rc=$?" >> $O/status.txt
done
timeout 900 python bench.py --steps 100 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" >> $O/status.txt
cat $O/status.txt
