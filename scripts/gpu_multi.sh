#!/bin/bash
# 2-GPU checks: NCCL tests, forward bench (weak scaling), training-step bench (strong scaling, config 4)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
N=${N:-2}
timeout 600 python -m pytest -q --timeout=300 -p no:cacheprovider tests/test_gpu_multi.py -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; python -c "import json; d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
python scripts/train_bench.py > gpurun_out/train_1gpu.json 2> gpurun_out/train_1gpu.err; tail -1 gpurun_out/train_1gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/train_bench.py > gpurun_out/train_${N}gpu.json 2> gpurun_out/train_${N}gpu.err; tail -1 gpurun_out/train_${N}gpu.json
