#!/bin/bash
# full GPU suite + smoke + bench (1 GPU)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest -q --timeout=600 --timeout-method=thread -p no:cacheprovider tests -m gpu > gpurun_out/tests_full.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_full.json"))
print("updown", round(d["value"]), d["ms_per_step"], "regat", round(d["regat"]["value"]), d["regat"]["ms_per_step"])
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["kernel"][:40], d["roofline"]["frac"], d["roofline"]["launch_ms"])
for k in d["roofline_kernels"]: print("  ", k["kernel"][:50], round(k["frac"], 3), round(k["launch_ms"] * 1e3, 1))
print("fp32tc", {k: (round(v["value"]), v["ms_per_step"], v["parity"]["n_equal"], v["parity"]["max_rel_logit_err"]) for k, v in d["fp32tc"].items() if isinstance(v, dict) and "value" in v})
print("train", d["train"].get("ms_per_step"), "cpu", d["cpu_baseline"])
PY
tail -3 gpurun_out/bench_full.err
