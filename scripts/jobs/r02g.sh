#!/bin/bash
# round-2 GPU job G (2 GPUs): NCCL paths — sharded forward, two-bucket gradient all-reduce, the bench at N = 2
set -u
O=gpurun_out/r02g
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $O/tests_multi.log 2>&1; echo "tests_multi rc=$?" >> $O/status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench 2gpu rc=$?" >> $O/status.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench 1gpu rc=$?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/tests_multi.log
