#!/bin/bash
# round 2, GPU call 1: FP64 pipe probe, A/B of the first kernel variants, phi self-test, parity of the new default,
# ncu capture of the FMA-peak microbenchmark
mkdir -p gpurun_out
scripts/probes/fp64_probe > gpurun_out/r2_fp64_probe.log 2>&1
python - > gpurun_out/r2_selftest_phi.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
from irrotational_warp_lab_b200.engine import get_engine
e = get_engine()
print(e.selftest_phi(count=1 << 28, seed=1))
print(e.selftest_phi(count=1 << 28, seed=2, coord_lo=1e-3, coord_hi=10.0))
print(e.selftest_phi(count=1 << 27, seed=3, rho=10.0, sigma=3.0, v=1.5, coord_lo=0.01, coord_hi=30.0))
print(e.selftest_divsqrt(count=1 << 26))
PY
scripts/ab2.sh gpurun_out/r2_ab1.log libiwb200_r1.so libiwb200.so libiwb200_nb.so libiwb200_nb2.so libiwb200_sr.so libiwb200_nbsr.so libiwb200_nbq.so libiwb200_nbs.so > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_1.log 2>&1
IWB_LIB=libiwb200_nb.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_pytest_gpu_nb.log 2>&1
python scripts/fma_peak_probe.py > gpurun_out/r2_fma_peak_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fma_peak -c 3 -o gpurun_out/r2_fma_peak python scripts/fma_peak_probe.py > gpurun_out/r2_fma_peak_ncu.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu_1.log gpurun_out/r2_pytest_gpu_nb.log
cat gpurun_out/r2_ab1.log gpurun_out/r2_selftest_phi.log
