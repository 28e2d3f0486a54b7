#!/bin/bash
# A/B of the attention pooling kernel: streaming (bulk-copy) kernel vs the register-staged one
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout -k 5 200 python -m pytest -q --timeout=60 -p no:cacheprovider tests/test_gpu_ops.py -m gpu -x -k "attention_pool" 2>&1 | tail -5
echo "--- stream"; timeout -k 5 100 python scripts/time_ops.py 2>&1 | grep -i "attention_pool"
echo "--- register-staged"; VQA_B200_POOL_STREAM=0 timeout -k 5 100 python scripts/time_ops.py 2>&1 | grep -i "attention_pool"
echo "--- bench stream"; timeout -k 5 150 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
echo "--- bench old"; VQA_B200_POOL_STREAM=0 timeout -k 5 150 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
