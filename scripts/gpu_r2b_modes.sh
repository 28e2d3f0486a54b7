#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python bench.py --precision fp32tc --steps 50 --warmup 5 --no-cpu-baseline --workloads updown,regat > gpurun_out/bench_fp32tc.json 2> gpurun_out/bench_fp32tc.err; echo "fp32tc rc=$?"
timeout 600 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --workloads updown,regat > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "fp32 rc=$?"
python - <<'PY'
import json
for n in ("fp32tc", "fp32"):
    d = json.load(open(f"gpurun_out/bench_{n}.json"))
    print(n, "updown", round(d["value"]), d["ms_per_step"], d.get("parity", {}).get("n_equal"), d.get("parity", {}).get("max_rel_logit_err"),
          "regat", round(d["regat"]["value"]), d["regat"]["ms_per_step"], d["regat"].get("parity", {}).get("n_equal"), "e2e", d.get("e2e", {}).get("value"))
PY
tail -2 gpurun_out/bench_fp32tc.err gpurun_out/bench_fp32.err
