#!/bin/bash
export PYTHONDONTWRITEBYTECODE=1 ONLY14=1
for d in 0 16 64 128 256 192; do VQA_B200_GRU_DEBUG=$d TABLE=1 timeout 120 python scripts/time_gru.py 2>&1 | tail -1; done
TOKMAX=64 TABLE=1 timeout 120 python scripts/time_gru.py 2>&1 | tail -1
TOKMAX=2000 TABLE=1 timeout 120 python scripts/time_gru.py 2>&1 | tail -1
