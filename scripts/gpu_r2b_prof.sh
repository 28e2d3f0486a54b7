#!/bin/bash
# round 2, second session: ncu launch list of the bench command + ncu --set full of the new kernels (1 GPU)
# (ncu cannot launch cooperative cluster kernels: the bf16 GRU runs as the single-CTA kernel, the fp32tc GRU one launch per step)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workloads updown,regat --no-graph > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workloads updown,regat --no-graph --no-parity > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench rc=$?"
for k in wv_split gru_split pool_split; do
  case $k in wv_split) pat=linear_tc_kernel;; gru_split) pat=linear_tc_kernel;; pool_split) pat=attention_pool_split_kernel;; esac
  timeout 200 python scripts/prof_kernel.py $k > gpurun_out/plain_$k.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 2 -f -o gpurun_out/prof_$k python scripts/prof_kernel.py $k > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"; tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
