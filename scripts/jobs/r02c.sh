#!/bin/bash
# round-2 GPU job C: fewer, larger TMA boxes in the GRU and the graph attention
set -u
O=gpurun_out/r02c
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_baseline_batch.py -q -m gpu -x > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/status.txt
for c in 0 16 32; do
  VQA_B200_GAT_CHASE=$c timeout 300 python scripts/timeline.py regat > $O/timeline_regat_chase$c.txt 2>&1; echo "timeline chase $c rc=$?" >> $O/status.txt
done
timeout 300 python scripts/timeline.py updown > $O/timeline_updown.txt 2>&1
for c in 0 16 24 32 48; do
  timeout 300 python bench.py --workloads regat --steps 100 --warmup 5 --chase $c --no-e2e --no-cpu-baseline --no-parity > $O/bench_regat_chase$c.json 2> $O/bench_regat_chase$c.err; echo "bench chase $c rc=$?" >> $O/status.txt
done
for g in 64x1 32x2 64x2; do
  VQA_B200_GRU_CFG=$g timeout 300 python bench.py --workloads updown --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-parity > $O/bench_updown_gru$g.json 2> $O/bench_updown_gru$g.err; echo "bench gru $g rc=$?" >> $O/status.txt
done
cat $O/status.txt
