#!/bin/bash
# round-2 GPU job M (8 GPUs): the driver's scaling command at N = 8 (and 4)
set -u
O=gpurun_out/r02m
mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench 8gpu rc=$?" >> $O/status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > $O/bench_4gpu.json 2> $O/bench_4gpu.err; echo "bench 4gpu rc=$?" >> $O/status.txt
cat $O/status.txt; tail -c 600 $O/bench_8gpu.err
