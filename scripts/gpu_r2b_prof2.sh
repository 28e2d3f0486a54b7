#!/bin/bash
# final-state ncu launch list of the bench command + ncu --set full of the fp32tc GRU step kernel (32-unit form)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workloads updown,regat --no-graph > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workloads updown,regat --no-graph --no-parity > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench rc=$?"
timeout 200 python scripts/prof_kernel.py gru_split > gpurun_out/plain_gru_split.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 2 -c 2 -f -o gpurun_out/prof_gru_split python scripts/prof_kernel.py gru_split > gpurun_out/ncu_gru_split.log 2>&1
echo "gru_split rc=$?"; tail -2 gpurun_out/ncu_gru_split.log
