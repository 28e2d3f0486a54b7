#!/bin/bash
# the rest of the GPU suite (after the first failure), fp32tc timeline with the 8-warp GRU epilogue
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1800 python -m pytest -q --timeout=600 --timeout-method=thread -p no:cacheprovider tests -m gpu > gpurun_out/tests_full.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/tests_full.log
PRECISION=fp32tc timeout 300 python scripts/timeline.py updown > gpurun_out/timeline_fp32tc_updown.txt 2>&1; tail -12 gpurun_out/timeline_fp32tc_updown.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --workloads updown > gpurun_out/bench_q.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_q.json"))
print("updown", round(d["value"]), d["ms_per_step"])
print("fp32tc", {k: (round(v["value"]), v["ms_per_step"], v["parity"]["n_equal"], v["parity"]["max_rel_logit_err"]) for k, v in d["fp32tc"].items() if isinstance(v, dict) and "value" in v})
PY
