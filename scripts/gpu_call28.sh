#!/bin/bash
# round 2, GPU call 28 (2 GPUs): the bounded wait of k_exchange_combine -- multi-GPU test file, emulated-rank test, 2-GPU bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -k "multi or peer or sharded or nccl or emulated" 2>&1 | tail -3 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_timeout_2gpu.json 2> gpurun_out/r2_timeout_2gpu.err
grep -v "^\*\|OMP_NUM\|NCCL version\|^$" gpurun_out/r2_timeout_2gpu.err | tail -3 | cut -c1-300; cut -c1-220 gpurun_out/r2_timeout_2gpu.json
