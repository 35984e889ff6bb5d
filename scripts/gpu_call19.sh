#!/bin/bash
# round 2, GPU call 19 (1 GPU): fused radial profile (k_energy3d_prof) -- parity tests, a small case of every kernel variant,
# timing against the two-pass path at n = 256 / 512 / 1024
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "radial or tail_correction or batch_equals" > gpurun_out/r2_pytest_radial_fused.log 2>&1
tail -15 gpurun_out/r2_pytest_radial_fused.log | cut -c1-250
timeout 600 python scripts/sanitize_case.py > gpurun_out/r2_racecheck_call19.log 2>&1
tail -6 gpurun_out/r2_racecheck_call19.log | cut -c1-200
echo "--- fused"; timeout 300 python scripts/f_rows_probe.py 2>&1 | head -3 | tee gpurun_out/r2_f_rows_fused.log
echo "--- two-pass"; IWB_PROFILE_FUSED=0 timeout 300 python scripts/f_rows_probe.py 2>&1 | head -3 | tee gpurun_out/r2_f_rows_twopass.log
