#!/bin/bash
set -u
O=gpurun_out/r02f
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_baseline_batch.py -q -m gpu -x > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/status.txt
B="python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-parity"
timeout 300 $B --workloads updown,regat --chase 0 > $O/bench_serial_nofence.json 2> $O/e1.err; echo "serial nofence rc=$?" >> $O/status.txt
VQA_B200_GRU_FENCE=1 timeout 300 $B --workloads updown,regat --chase 0 > $O/bench_serial_fence.json 2> $O/e2.err; echo "serial fence rc=$?" >> $O/status.txt
for pm in 200 260 320; do
  timeout 300 $B --workloads regat --chase 0 --overlap 1 --side-sms 64 --side-permille $pm > $O/bench_regat_ov64_$pm.json 2> $O/e3.err; echo "ov64 $pm rc=$?" >> $O/status.txt
done
for pm in 780 820 860; do
  timeout 300 $B --workloads regat --chase 0 --overlap 1 --side-sms 128 --side-permille $pm > $O/bench_regat_ov128_$pm.json 2> $O/e4.err; echo "ov128 $pm rc=$?" >> $O/status.txt
done
timeout 300 $B --workloads updown --overlap 1 --side-sms 128 --side-permille 800 > $O/bench_updown_ov128_800.json 2> $O/e5.err; echo "updown ov128 rc=$?" >> $O/status.txt
cat $O/status.txt
