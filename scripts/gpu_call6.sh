#!/bin/bash
# round 2, GPU call 6 (8 GPUs): the scaling series 1/2/4/8 with the driver's flags, NCCL exchange beside the peer exchange at 8,
# config 5 at 1 and 8 GPUs
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then CMD="python"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"; fi
  timeout 600 $CMD bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scale_${N}gpu.json 2> gpurun_out/r2_scale_${N}gpu.err
done
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530"
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl > gpurun_out/r2_scale_8gpu_nccl.json 2> gpurun_out/r2_scale_8gpu_nccl.err
IWB_BENCH_VERBOSE=1 timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_scale_8gpu_b.json 2> gpurun_out/r2_scale_8gpu_b.err
timeout 600 $TR8 bench.py --gpus 8 --config sweep_opt --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5_8gpu.json 2> gpurun_out/r2_bench_cfg5_8gpu.err
timeout 600 python bench.py --config sweep_opt --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5_1gpu_b.json 2> gpurun_out/r2_bench_cfg5_1gpu_b.err
for f in r2_scale_1gpu r2_scale_2gpu r2_scale_4gpu r2_scale_8gpu r2_scale_8gpu_nccl r2_scale_8gpu_b r2_bench_cfg5_8gpu r2_bench_cfg5_1gpu_b; do echo "== $f"; grep -v "^\*\|OMP_NUM\|NCCL version" gpurun_out/$f.err | tail -4 | cut -c1-300; cut -c1-260 gpurun_out/$f.json; done
