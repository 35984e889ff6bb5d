#!/bin/bash
# round 2, GPU call 25 (8 GPUs): the slab decomposition of axis 0 (the north star's wording) through bench.py, beside the
# default tile shards measured in calls 16 / 22
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 8 --steps 20 --warmup 5 --shard slab --no-cpu-baseline > gpurun_out/r2_slab_8gpu.json 2> gpurun_out/r2_slab_8gpu.err
grep -v "^\*\|OMP_NUM\|NCCL version\|^$" gpurun_out/r2_slab_8gpu.err | tail -3 | cut -c1-300; cut -c1-300 gpurun_out/r2_slab_8gpu.json
