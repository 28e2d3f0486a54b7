#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest -q --timeout=150 --timeout-method=thread -p no:cacheprovider tests/test_gpu_forward.py -m gpu -k "golden or ragged" > gpurun_out/tests_t5.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/tests_t5.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
