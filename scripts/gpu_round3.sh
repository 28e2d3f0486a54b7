#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q --timeout=300 -p no:cacheprovider tests -m gpu -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests.log
timeout 300 python scripts/time_ops.py 2>&1 | grep -E "gemm|argmax|pool" 
VQA_B200_NO_PDL=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NO_PDL updown', d['value'], d['ms_per_step'])"
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_updown.json 2> gpurun_out/bench_updown.err; python -c "import json; d=json.load(open('gpurun_out/bench_updown.json')); print('PDL updown', d['value'], d['ms_per_step'], d['e2e']['value'])"
VQA_B200_NO_PDL=1 python bench.py --workload regat --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NO_PDL regat', d['value'], d['ms_per_step'])"
python bench.py --workload regat --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_regat.json 2> gpurun_out/bench_regat.err; python -c "import json; d=json.load(open('gpurun_out/bench_regat.json')); print('PDL regat', d['value'], d['ms_per_step'])"
timeout 300 python scripts/time_qcap.py 2>&1 | tail -1
