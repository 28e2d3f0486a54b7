#!/bin/bash
# ncu --set full captures of the hot kernels (run under gpurun, 1 GPU); each only after its plain run exits 0
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
for k in ${KERNELS:-gat pool wide cls1}; do
  case $k in
    gat) pat=graph_attention_tc_kernel;; pool) pat=attention_pool;; wide|cls1|wv) pat=linear_tc_kernel;;
    gru) pat="gru_(pair|persistent)_kernel";; relation) pat=relation_labels_kernel;; esac
  timeout 200 python scripts/prof_kernel.py $k > gpurun_out/plain_$k.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 2 -f -o gpurun_out/prof_$k python scripts/prof_kernel.py $k > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"; tail -2 gpurun_out/ncu_$k.log
done
