#!/bin/bash
# round-2 GPU job A: new tests, full suite, timelines and benches of both schedules
set -u
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_baseline_batch.py -x -q -m gpu -s > $O/tests_new.log 2>&1; echo "tests_new rc=$?" >> $O/status.txt
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_baseline_batch.py > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/status.txt
for wl in updown regat; do
  VQA_B200_OVERLAP=1 timeout 300 python scripts/timeline.py $wl > $O/timeline_${wl}_overlap.txt 2>&1; echo "timeline $wl overlap rc=$?" >> $O/status.txt
  VQA_B200_OVERLAP=1 VQA_B200_GRU_COOP=0 timeout 300 python scripts/timeline.py $wl > $O/timeline_${wl}_overlap_nocoop.txt 2>&1; echo "timeline $wl nocoop rc=$?" >> $O/status.txt
  VQA_B200_OVERLAP=0 timeout 300 python scripts/timeline.py $wl > $O/timeline_${wl}_serial.txt 2>&1; echo "timeline $wl serial rc=$?" >> $O/status.txt
done
timeout 900 python bench.py --steps 100 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 100 --warmup 5 --overlap 0 --no-e2e --no-cpu-baseline --no-parity > $O/bench_serial.json 2> $O/bench_serial.err; echo "bench serial rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 100 --warmup 5 --no-graph --no-e2e --no-cpu-baseline --no-parity > $O/bench_nograph.json 2> $O/bench_nograph.err; echo "bench nograph rc=$?" >> $O/status.txt
VQA_B200_GRU_COOP=0 timeout 600 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-parity > $O/bench_nocoop.json 2> $O/bench_nocoop.err; echo "bench nocoop rc=$?" >> $O/status.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?" >> $O/status.txt
cat $O/status.txt
