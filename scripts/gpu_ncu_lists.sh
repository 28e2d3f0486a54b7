#!/bin/bash
# ncu launch lists of the bench command.  ncu cannot launch the cooperative cluster kernel gru_pair_kernel
# ("LaunchFailed" under the profiler), so these runs use the single-CTA GRU (VQA_B200_GRU_PAIR=0: 144 vs 138 us);
# the CUPTI timelines of the default path are in timeline_*.txt.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1 VQA_B200_GRU_PAIR=0
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_updown.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_updown.log 2>&1
echo "ncu updown rc=$?"
python bench.py --workload regat --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_regat.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_regat.csv python bench.py --workload regat --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_regat.log 2>&1
echo "ncu regat rc=$?"
KERNELS="gru" bash scripts/gpu_profile.sh
