#!/bin/bash
# round 2, GPU call 15: run-length histogram kernel -- parity, timing against the lane-private one, per-kernel times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "radial or tail_correction" > gpurun_out/r2_pytest_radial3.log 2>&1
timeout 600 python scripts/f_rows_probe.py 2>&1 | head -4 > gpurun_out/r2_f_rows_probe_rle.log
IWB_PROF_LANE_PRIVATE=1 timeout 600 python scripts/radial_probe.py 1024 > gpurun_out/r2_radial_lane_private.log 2>&1
python scripts/radial_probe.py 1024 > gpurun_out/r2_radial_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_radial_rle.csv python scripts/radial_probe.py 1024 > gpurun_out/r2_ncu_radial3.log 2>&1
tail -4 gpurun_out/r2_pytest_radial3.log; cat gpurun_out/r2_f_rows_probe_rle.log gpurun_out/r2_radial_lane_private.log gpurun_out/r2_radial_plain3.log
grep -E "k_radial|k_energy3d" gpurun_out/r2_launches_radial_rle.csv | awk -F'","' '{print $5, $NF}' | cut -c1-100 | tail -12
