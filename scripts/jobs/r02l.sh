#!/bin/bash
set -u
O=gpurun_out/r02l
mkdir -p $O
for c in 1 2 4 8; do VQA_B200_GRU_PAIR=0 VQA_B200_GRU_CLUSTER=$c timeout 120 python scripts/gru_probe.py >> $O/gru_probe.txt 2>&1; echo "single-CTA cluster=$c" >> $O/gru_probe.txt; done
cat $O/gru_probe.txt
