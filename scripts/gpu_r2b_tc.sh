#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q --timeout=600 --timeout-method=thread -p no:cacheprovider tests/test_gpu_baseline_batch.py -m gpu -k "fp32tc or fp32_matches" -s > gpurun_out/tests_tc.log 2>&1
echo "tests rc=$?"; grep -E "fp32tc parity|passed|failed|Error|error|assert" gpurun_out/tests_tc.log | tail -30
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --workloads updown,regat > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_tc.json"))
print("updown", d["value"], d["ms_per_step"], "regat", d.get("regat", {}).get("value"))
print(json.dumps(d.get("fp32tc"), indent=1)[:3000])
PY
tail -3 gpurun_out/bench_tc.err
