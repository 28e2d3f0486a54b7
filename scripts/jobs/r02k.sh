#!/bin/bash
set -u
O=gpurun_out/r02k
mkdir -p $O
for d in 0 1 2 3 4 8 12 15; do VQA_B200_GRU_DEBUG=$d timeout 120 python scripts/gru_probe.py >> $O/gru_probe.txt 2>&1; done
for d in 0 4 8 12; do VQA_B200_GRU_CFG=64x2 VQA_B200_GRU_DEBUG=$d timeout 120 python scripts/gru_probe.py >> $O/gru_probe.txt 2>&1; done
cat $O/gru_probe.txt
