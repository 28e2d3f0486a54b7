#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 200 python scripts/prof_kernel.py relation > gpurun_out/plain_relation.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relation_labels_kernel -s 1 -c 1 -f -o gpurun_out/prof_relation python scripts/prof_kernel.py relation > gpurun_out/ncu_relation.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_relation.log
