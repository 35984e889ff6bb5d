#!/bin/bash
# round 2, GPU call 14: final single-GPU records -- full GPU suite, convergence table (config 3), both bench arms with the
# driver's flags, config 5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1
timeout 600 python -m irrotational_warp_lab_b200.cli convergence --n-values 40,60,80,100,128,200,256,400,512 --out gpurun_out/r2_convergence/convergence.json > gpurun_out/r2_convergence.log 2>&1
timeout 600 python -m irrotational_warp_lab_b200.cli reproduce --mode 3d --n 100 --dtype float32 --out gpurun_out/r2_convergence/repro_n100_f32.json >> gpurun_out/r2_convergence.log 2>&1
timeout 600 python -m irrotational_warp_lab_b200.cli reproduce --mode axisym --nx 2400 --ny 1200 --out gpurun_out/r2_convergence/repro_axisym_2400x1200.json >> gpurun_out/r2_convergence.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench1_final.json 2> gpurun_out/r2_bench1_final.err
timeout 600 python bench.py --config sweep_opt --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5_1gpu_final.json 2> gpurun_out/r2_bench_cfg5_1gpu_final.err
tail -4 gpurun_out/r2_pytest_gpu_final.log | cut -c1-200; grep -A14 "^=====" gpurun_out/r2_convergence.log | head -20
cut -c1-200 gpurun_out/r2_bench1_final.json; cut -c1-200 gpurun_out/r2_bench_ref_final.json; cut -c1-200 gpurun_out/r2_bench_cfg5_1gpu_final.json
