#!/bin/bash
# round-2 first check on 2 GPUs: GPU tests (incl. the 2-GPU DP step), default bench at N=1 and N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 2>&1 | tail -15 > gpurun_out/r2_gpu_tests.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --batch 8192 --humans 10 --steps 40 --dp-comm nccl > gpurun_out/r2_train_n2_nccl.log 2>&1
cat gpurun_out/r2_gpu_tests.log; tail -c 1500 gpurun_out/r2_bench_n1.log; tail -c 1500 gpurun_out/r2_bench_n2.log; tail -c 800 gpurun_out/r2_train_n2_nccl.log
