#!/bin/bash
export PYTHONDONTWRITEBYTECODE=1
for d in 0 1 2 4 8 5 12 13; do VQA_B200_GRUS_DEBUG=$d timeout 100 python scripts/time_gru_split.py 2>&1 | tail -1; done
VQA_B200_GRU_SPLIT_PERSIST=0 timeout 100 python scripts/time_gru_split.py 2>&1 | tail -1
