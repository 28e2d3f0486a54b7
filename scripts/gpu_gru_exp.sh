#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
for c in 1 2 4 8; do VQA_B200_GRU_CLUSTER=$c timeout 120 python scripts/time_gru.py 2>&1 | tail -1 | sed "s/^/cluster=$c /"; done 2>&1 | tee gpurun_out/gru_exp2.log
