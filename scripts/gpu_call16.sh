#!/bin/bash
# round 2, GPU call 16 (8 GPUs): final records -- GPU suite (incl. the 2-GPU NCCL / peer test), scaling series 1/2/4/8 with
# the driver's flags, NCCL exchange beside the peer exchange at 8, config 5 at 8
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_final8.log 2>&1
timeout 300 python scripts/f_rows_probe.py 2>&1 | head -3 > gpurun_out/r2_f_rows_probe_final.log
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then CMD="python"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"; fi
  timeout 600 $CMD bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scale_${N}gpu.json 2> gpurun_out/r2_scale_${N}gpu.err
done
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530"
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl --no-cpu-baseline > gpurun_out/r2_scale_8gpu_nccl.json 2> gpurun_out/r2_scale_8gpu_nccl.err
timeout 600 $TR8 bench.py --gpus 8 --config sweep_opt --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5_8gpu.json 2> gpurun_out/r2_bench_cfg5_8gpu.err
tail -4 gpurun_out/r2_pytest_gpu_final8.log | cut -c1-200; cat gpurun_out/r2_f_rows_probe_final.log
for f in r2_scale_1gpu r2_scale_2gpu r2_scale_4gpu r2_scale_8gpu r2_scale_8gpu_nccl r2_bench_cfg5_8gpu; do echo "== $f"; grep -v "^\*\|OMP_NUM\|NCCL version\|^$" gpurun_out/$f.err | tail -3 | cut -c1-300; cut -c1-200 gpurun_out/$f.json; done
