#!/bin/bash
# round 2, GPU call 27 (1 GPU): ncu evidence of the final tree -- launch list of a bench run, full captures of k_energy3d
# (n = 1024^3) and of the fused-profile kernel
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_plain_short.json 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches_final.log 2>&1
python scripts/k3_probe.py 1024 3 > gpurun_out/r2_k3_plain_final.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_energy3d -s 2 -c 1 -f -o gpurun_out/r2c_k3 python scripts/k3_probe.py 1024 3 > gpurun_out/r2c_k3_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_energy3d_prof -s 1 -c 1 -f -o gpurun_out/r2c_prof python -c "
import sys; sys.path.insert(0, '.')
from irrotational_warp_lab_b200 import tail
for _ in range(2): p = tail.radial_profile_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=1024)
print(p['path'], p['kernel_ms'])
" > gpurun_out/r2c_prof_ncu.log 2>&1
cat gpurun_out/r2_k3_plain_final.log | cut -c1-200; tail -2 gpurun_out/r2c_prof_ncu.log; grep -c k_energy3d gpurun_out/r2_launches_bench_final.csv
