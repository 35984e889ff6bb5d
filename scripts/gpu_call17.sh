#!/bin/bash
# round 2, GPU call 17 (1 GPU): axisymmetric evaluator -- A/B of the phi phase (one point at a time / the four rows of a
# column together / the same with the host's operand-range proof and seeded divisions / 3 CTAs per SM), parity of every
# variant, bench line for BASELINE config 2 and the per-kernel launch list
mkdir -p gpurun_out
D=irrotational-warp-lab_b200
out=gpurun_out/r2_ab17_axisym.log; : > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $out 2>&1
for v in ax0 ax1 ax2 ax2m3 ax0m3; do
  echo "=== $v" >> $out
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$D/libiwb200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "axisym or unusual or tiny_sigma" 2>&1 | tail -2 | cut -c1-200 >> $out
  for rep in 1 2; do
    IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$D/libiwb200_$v.so timeout 300 python scripts/axisym_probe.py 30 2>&1 | cut -c1-200 >> $out
  done
done
timeout 600 python bench.py --config axisym --steps 50 --warmup 5 > gpurun_out/r2_bench_axisym.json 2> gpurun_out/r2_bench_axisym.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_axisym.csv python bench.py --config axisym --steps 5 --warmup 3 > gpurun_out/r2_ncu_axisym.log 2>&1
for v in ax0 ax2; do
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$D/libiwb200_$v.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_axisym -s 3 -c 1 -f -o gpurun_out/r2_axisym_$v python scripts/axisym_probe.py 2 > gpurun_out/r2_ncu_axisym_$v.log 2>&1
done
cat $out; cut -c1-600 gpurun_out/r2_bench_axisym.json; tail -3 gpurun_out/r2_bench_axisym.err
grep -c k_axisym gpurun_out/r2_launches_axisym.csv; grep "k_axisym\|k_reduce" gpurun_out/r2_launches_axisym.csv | tail -4 | cut -c1-300
