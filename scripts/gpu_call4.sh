#!/bin/bash
# round 2, GPU call 4 (2 GPUs): NCCL / peer-exchange multi-process paths, bench at N = 2, config 5, float32 table
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.log 2>&1
nvidia-smi topo -m >> gpurun_out/r2_gpus.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_exchange" > gpurun_out/r2_pytest_peer_emulated.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_multi_2gpu.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench2_peer.json 2> gpurun_out/r2_bench2_peer.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl > gpurun_out/r2_bench2_nccl.json 2> gpurun_out/r2_bench2_nccl.err
timeout 600 $TR bench.py --gpus 2 --config sweep_opt --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg5_2gpu.json 2> gpurun_out/r2_bench_cfg5_2gpu.err
timeout 600 python bench.py --config sweep_opt --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg5_1gpu.json 2> gpurun_out/r2_bench_cfg5_1gpu.err
timeout 300 python scripts/f32_error_table.py > gpurun_out/r2_f32_error_table.log 2>&1
timeout 300 python scripts/e2e_probe.py 1024 > gpurun_out/r2_e2e_probe.log 2>&1
timeout 300 python scripts/e2e_probe.py 256 >> gpurun_out/r2_e2e_probe.log 2>&1
tail -5 gpurun_out/r2_pytest_peer_emulated.log gpurun_out/r2_pytest_gpu_multi_2gpu.log
for f in r2_bench2_peer r2_bench2_nccl r2_bench_cfg5_2gpu r2_bench_cfg5_1gpu; do echo "== $f"; tail -3 gpurun_out/$f.err | cut -c1-300; cut -c1-400 gpurun_out/$f.json; done
cat gpurun_out/r2_f32_error_table.log gpurun_out/r2_e2e_probe.log
