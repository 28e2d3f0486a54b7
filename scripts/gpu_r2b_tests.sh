#!/bin/bash
# the whole GPU suite (1 GPU), bounded: 200 s per test
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest -q --timeout=200 --timeout-method=thread -p no:cacheprovider tests -m gpu > gpurun_out/tests_full.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
