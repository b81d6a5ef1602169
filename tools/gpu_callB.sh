#!/bin/bash
# 2-GPU validation: bank-sharded core over NCCL (2 ranks) + the driver's N=2 bench launch (c3 clip-parallel + the c5 leg)
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 300 python -m pytest tests/test_sharded_gpu.py tests/test_memory_read_gpu.py -m gpu -x -q > gpurun_out/cB_tests.log 2>&1; tail -4 gpurun_out/cB_tests.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/cB_bench_2gpu.json 2> gpurun_out/cB_bench_2gpu.err; tail -c 2500 gpurun_out/cB_bench_2gpu.json; tail -3 gpurun_out/cB_bench_2gpu.err
