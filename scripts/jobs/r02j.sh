#!/bin/bash
# round-2 GPU job J: ncu launch list of the bench step + `--set full` captures of the hot kernels (one GPU)
set -u
O=gpurun_out/r02j
mkdir -p $O
BENCH="python bench.py --steps 3 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-parity --workloads updown,regat"
$BENCH > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv $BENCH > $O/ncu_bench.log 2>&1
echo "launch list rc=$?" >> $O/status.txt
for k in gat wv wide pool gru; do
  case $k in gat) pat=graph_attention;; wv) pat=linear_tc;; wide) pat=linear_tc;; pool) pat=attention_pool;; gru) pat=gru_;; esac
  python scripts/prof_kernel.py $k > $O/plain_$k.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 2 -o $O/prof_$k python scripts/prof_kernel.py $k > $O/ncu_$k.log 2>&1
  echo "ncu $k rc=$?" >> $O/status.txt
done
cat $O/status.txt
