#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 400 python -m pytest -q --timeout=100 --timeout-method=thread -p no:cacheprovider tests -m gpu -k "relation or labels or regat_full_batch or smoke" > gpurun_out/tests_rel.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/tests_rel.log
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from oracle import vqa_oracle as O
from vqa_collection_b200 import ops
import numpy as np
for B, K in ((1 << 18, 36), (1024, 36), (4096, 20), (2048, 45), (1024, 64)):
    boxes = torch.from_numpy(O.make_boxes(B, K, 1)).cuda()
    for _ in range(3): lab = ops.relation_labels(boxes, 640, 480)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): lab = ops.relation_labels(boxes, 640, 480)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = B * (K * 16 + K * K)
    n = min(B, 512)
    ref = O.relation_graph_batch(boxes[:n].cpu().numpy(), 640, 480).astype(np.uint8)
    mism = int((lab[:n].cpu().numpy() != ref).sum())
    print(f"B={B} K={K}: {ms*1e3:.1f} us, {byts/ms/1e6:.0f} GB/s, mismatches in first {n} images: {mism} of {n*K*K}")
PY
