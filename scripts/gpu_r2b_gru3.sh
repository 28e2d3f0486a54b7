#!/bin/bash
export PYTHONDONTWRITEBYTECODE=1 ONLY14=1
for ns in 0 300 700 1200 2000; do echo -n "pause=$ns "; VQA_B200_GRU_GI_PAUSE=$ns TABLE=1 timeout 120 python scripts/time_gru.py 2>&1 | tail -1; done
TABLE=0 timeout 120 python scripts/time_gru.py 2>&1 | tail -1
