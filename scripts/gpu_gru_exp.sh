#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest -q --timeout=300 -p no:cacheprovider tests -m gpu -x 2>&1 | tail -3
for d in 0 1 2 3; do VQA_B200_GRU_DEBUG=$d timeout 120 python scripts/time_gru.py 2>&1 | tail -1; done 2>&1 | tee gpurun_out/gru_exp.log
