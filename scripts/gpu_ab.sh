#!/bin/bash
# A/B of one environment switch on the forward timelines: bash scripts/gpu_ab.sh VAR
export PYTHONDONTWRITEBYTECODE=1
V=$1
for wl in updown regat; do
  echo "== $wl default"; python scripts/timeline.py $wl 2>&1 | grep -E "pool|graph_att|step span"
  echo "== $wl $V=1"; env $V=1 python scripts/timeline.py $wl 2>&1 | grep -E "pool|graph_att|step span"
done
