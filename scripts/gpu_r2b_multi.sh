#!/bin/bash
# 8-GPU box: NCCL tests (2 ranks), the driver's bench command at N = 8, 4 and 2 (--steps 20 --warmup 3)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 300 python -m pytest -q --timeout=200 -p no:cacheprovider tests/test_gpu_multi.py -m gpu 2>&1 | tail -3
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], "updown", round(d["value"]), d["ms_per_step"], "regat", round(d["regat"]["value"]), d["regat"]["ms_per_step"], "e2e", round(d["e2e"]["value"]))
print("  per-rank", d["per_rank_ms_total"], "train", d["train"].get("ms_per_step"))
print("  fp32tc", {k: (round(v["value"]), v["ms_per_step"]) for k, v in d["fp32tc"].items() if isinstance(v, dict) and "value" in v})
PY
done
