#!/bin/bash
set -u
O=gpurun_out/r02e
mkdir -p $O
for d in 0 1 2 4 6 8 14 15; do
  VQA_B200_GAT_DEBUG=$d timeout 120 python scripts/gat_probe.py >> $O/gat_probe.txt 2>&1
done
cat $O/gat_probe.txt
