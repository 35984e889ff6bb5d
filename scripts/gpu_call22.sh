#!/bin/bash
# round 2, GPU call 22 (8 GPUs): final tree -- the multi-GPU test file, scaling points 2 / 4 / 8 with the driver's flags,
# config 5 at 8 GPUs (rolling batch pipeline)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -rs > gpurun_out/r2_pytest_gpu_multi_final.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu_multi_final.log | cut -c1-200
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_final_${N}gpu.json 2> gpurun_out/r2_final_${N}gpu.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29630 bench.py --gpus 8 --config sweep_opt --steps 5 --warmup 3 > gpurun_out/r2_final_cfg5_8gpu.json 2> gpurun_out/r2_final_cfg5_8gpu.err
for f in r2_final_2gpu r2_final_4gpu r2_final_8gpu r2_final_cfg5_8gpu; do echo "== $f"; grep -v "^\*\|OMP_NUM\|NCCL version\|^$" gpurun_out/$f.err | tail -2 | cut -c1-300; cut -c1-200 gpurun_out/$f.json; done
