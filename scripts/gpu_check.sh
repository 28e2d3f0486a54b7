#!/bin/bash
# Staged GPU check (run under gpurun): safe kernels first, tcgen05 GEMM in its own
# process so that a hang there does not lose the other results.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
export PYTHONDONTWRITEBYTECODE=1
T="timeout 600 python -m pytest -q --timeout=300 --timeout-method=thread -p no:cacheprovider"
echo "== stage 1: fp32 / non-tensor kernels (VQA_B200_FORCE_SIMT=1)"
VQA_B200_FORCE_SIMT=1 $T tests/test_gpu_ops.py tests/test_gpu_forward.py -m gpu > gpurun_out/stage1.log 2>&1; tail -15 gpurun_out/stage1.log
echo "== stage 2: tcgen05 GEMM"
$T tests/test_gpu_ops.py -m gpu -k "linear" > gpurun_out/stage2.log 2>&1; tail -5 gpurun_out/stage2.log
echo "== stage 3: everything on the tensor-core path"
$T tests -m gpu > gpurun_out/stage3.log 2>&1; tail -15 gpurun_out/stage3.log
