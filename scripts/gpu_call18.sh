#!/bin/bash
# round 2, GPU call 18 (1 GPU): full GPU suite on the new library (axisym: grouped phi with the host proof, 3 CTAs per SM;
# batches of small grids share the SMs, rolling pipeline; tile reduction with 8 loads in flight); batch A/B
# (IWB_BATCH_SPLIT = 0 / 1 / 2); config 5 with and without the split; axisym bench line, launch list and ncu capture;
# k_energy3d with the {x, x*x} pair loaded right before task P (IWB_XLATE) against the default; headline bench
mkdir -p gpurun_out
D=irrotational-warp-lab_b200
timeout 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_call18.log 2>&1
out=gpurun_out/r2_ab18_batch_split.log; : > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $out 2>&1
for rep in 1 2; do
  for sp in 0 1 2; do
    IWB_BATCH_SPLIT=$sp timeout 300 python scripts/batch_probe.py 50 5 2>&1 | cut -c1-200 >> $out
  done
done
for sp in 0 1; do
  IWB_BATCH_SPLIT=$sp timeout 600 python bench.py --config sweep_opt --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5_1gpu_split$sp.json 2> gpurun_out/r2_bench_cfg5_1gpu_split$sp.err
done
timeout 300 python scripts/axisym_probe.py 30 > gpurun_out/r2_axisym_probe_final.log 2>&1
timeout 600 python bench.py --config axisym --steps 50 --warmup 5 > gpurun_out/r2_bench_axisym.json 2> gpurun_out/r2_bench_axisym.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_axisym.csv python bench.py --config axisym --steps 5 --warmup 3 > gpurun_out/r2_ncu_axisym.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_axisym -s 3 -c 1 -f -o gpurun_out/r2_axisym_final python scripts/axisym_probe.py 2 > gpurun_out/r2_ncu_axisym_final.log 2>&1
bash scripts/ab2.sh gpurun_out/r2_ab18_xlate.log $D/libiwb200.so $D/libiwb200_xlate.so > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_call18.json 2> gpurun_out/r2_bench_call18.err
tail -3 gpurun_out/r2_pytest_gpu_call18.log | cut -c1-200
cat $out
for sp in 0 1; do cut -c1-330 gpurun_out/r2_bench_cfg5_1gpu_split$sp.json; done
cat gpurun_out/r2_axisym_probe_final.log | cut -c1-200
cut -c1-400 gpurun_out/r2_bench_axisym.json
grep "k_axisym\|k_reduce" gpurun_out/r2_launches_axisym.csv | tail -2 | cut -c60-300
grep -v "^name\|^NVIDIA" gpurun_out/r2_ab18_xlate.log | cut -c1-180
cut -c1-300 gpurun_out/r2_bench_call18.json; tail -2 gpurun_out/r2_bench_call18.err
