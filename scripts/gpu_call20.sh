#!/bin/bash
# round 2, GPU call 20 (1 GPU): full GPU suite on the library with the fused radial profile; f3 / f4 timings with the
# automatic path selection, the forced fused kernel and the two-pass path; headline bench as a regression check
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_call20.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_call20.log | cut -c1-200
{ echo "# automatic (fused for >= 2^29 points)"; timeout 300 python scripts/f_rows_probe.py 2>&1;
  echo "# IWB_PROFILE_FUSED=2 (fused whenever eligible)"; IWB_PROFILE_FUSED=2 timeout 300 python scripts/f_rows_probe.py 2>&1 | head -3;
  echo "# IWB_PROFILE_FUSED=0 (two-pass)"; IWB_PROFILE_FUSED=0 timeout 300 python scripts/f_rows_probe.py 2>&1 | head -3; } > gpurun_out/r2_f_rows_probe_call20.log
cat gpurun_out/r2_f_rows_probe_call20.log | cut -c1-220
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_call20.json 2> gpurun_out/r2_bench_call20.err
cut -c1-260 gpurun_out/r2_bench_call20.json; tail -2 gpurun_out/r2_bench_call20.err
