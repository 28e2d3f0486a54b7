#!/bin/bash
# final state of the session: whole GPU suite (bounded), smoke, the driver's bench command, timelines
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest -q --timeout=200 --timeout-method=thread -p no:cacheprovider tests -m gpu > gpurun_out/tests_full.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_full.json"))
print("updown", round(d["value"]), d["ms_per_step"], "regat", round(d["regat"]["value"]), d["regat"]["ms_per_step"], "e2e", round(d["e2e"]["value"]))
for k in d["roofline_kernels"]: print("  ", k["kernel"][:50], round(k["frac"], 3), round(k["launch_ms"] * 1e3, 1))
print("fp32tc", {k: (round(v["value"]), v["ms_per_step"], v["parity"]["n_equal"], v["parity"]["max_rel_logit_err"]) for k, v in d["fp32tc"].items() if isinstance(v, dict) and "value" in v})
for k in d["fp32tc"]["updown"]["roofline_kernels"]: print("  ", k["kernel"][:60], round(k["frac"], 3), round(k["launch_ms"] * 1e3, 1))
print("train", d["train"].get("ms_per_step"), "cpu", d["cpu_baseline"]["value"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; head -c 600 gpurun_out/bench_ref.json; echo
PRECISION=fp32tc timeout 200 python scripts/timeline.py updown > gpurun_out/timeline_fp32tc_updown.txt 2>&1
timeout 200 python scripts/timeline.py updown > gpurun_out/timeline_updown.txt 2>&1
timeout 200 python scripts/timeline.py regat > gpurun_out/timeline_regat.txt 2>&1; tail -14 gpurun_out/timeline_regat.txt
