#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 120 python scripts/small_cases_new_kernels.py > gpurun_out/san_plain.log 2>&1; echo "plain rc=$?"; tail -14 gpurun_out/san_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python scripts/small_cases_new_kernels.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|error|Error" gpurun_out/san_memcheck.log | head -20; tail -3 gpurun_out/san_memcheck.log
