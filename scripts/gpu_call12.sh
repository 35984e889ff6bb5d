#!/bin/bash
# round 2, GPU call 12 (2 GPUs): sharded API with the cached parameter block, multi-process tests, smoke()
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_multi_2gpu.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench2_c.json 2> gpurun_out/r2_bench2_c.err
tail -3 gpurun_out/r2_smoke.log gpurun_out/r2_pytest_gpu_multi_2gpu.log; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench2_c.err | tail -3; cut -c1-200 gpurun_out/r2_bench2_c.json
