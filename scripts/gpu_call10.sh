#!/bin/bash
# round 2, GPU call 10: full GPU suite, radial profile timing after the two-stage reduce, ncu of the histogram kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_10.log 2>&1
python scripts/radial_probe.py 1024 > gpurun_out/r2_radial_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_radial_hist_fast -s 8 -c 1 -o gpurun_out/r2_hist python scripts/radial_probe.py 1024 > gpurun_out/r2_hist_ncu.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_10.log | cut -c1-200; cat gpurun_out/r2_radial_plain2.log; tail -2 gpurun_out/r2_hist_ncu.log
