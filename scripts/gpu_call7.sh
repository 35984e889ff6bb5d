#!/bin/bash
# round 2, GPU call 7: fast radial-histogram kernel (parity + timing), span-schedule sweep for the 8-rank tile shards
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "radial or tail_correction" > gpurun_out/r2_pytest_radial.log 2>&1
timeout 600 python scripts/f_rows_probe.py > gpurun_out/r2_f_rows_probe_after.log 2>&1
: > gpurun_out/r2_shard_static_sweep.log
for pct in 60 70 80 85 90 95; do
  IWB_K3_STATIC_PCT=$pct timeout 200 python scripts/shard_probe.py 8 4 >> gpurun_out/r2_shard_static_sweep.log 2>&1
done
for pct in 70 80 90; do
  IWB_K3_STATIC_PCT=$pct timeout 200 python scripts/k3_probe.py 1024 4 2>&1 | cut -c1-120 >> gpurun_out/r2_shard_static_sweep.log
  IWB_K3_STATIC_PCT=$pct timeout 200 python scripts/k3_probe.py 256 6 2>&1 | cut -c1-120 >> gpurun_out/r2_shard_static_sweep.log
done
tail -4 gpurun_out/r2_pytest_radial.log; cat gpurun_out/r2_f_rows_probe_after.log gpurun_out/r2_shard_static_sweep.log
