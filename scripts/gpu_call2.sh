#!/bin/bash
# round 2, GPU call 2: A/B of the tally / j-face-tile variants, full GPU test-suite on the new default build
mkdir -p gpurun_out
scripts/ab2.sh gpurun_out/r2_ab2.log libiwb200_r1.so libiwb200_seed.so libiwb200_t2.so libiwb200_t2sr.so libiwb200_t2jf.so libiwb200.so > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_2.log 2>&1
python scripts/fma_peak_probe.py > gpurun_out/r2_fma_peak_plain2.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu_2.log
cat gpurun_out/r2_fma_peak_plain2.log
cat gpurun_out/r2_ab2.log
