#!/bin/bash
# round 2, GPU call 8: lane-private histogram v2, span-schedule sweep for 2 / 4 / 8 tile shards, host cost of the sharded API
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "radial or tail_correction" > gpurun_out/r2_pytest_radial2.log 2>&1
timeout 600 python scripts/f_rows_probe.py > gpurun_out/r2_f_rows_probe_after2.log 2>&1
: > gpurun_out/r2_shard_static_sweep2.log
for P in 8 4 2; do for pct in 50 60 65 70 75 80; do
  IWB_K3_STATIC_PCT=$pct timeout 200 python scripts/shard_probe.py $P 4 >> gpurun_out/r2_shard_static_sweep2.log 2>&1
done; done
timeout 300 python scripts/sharded_host_probe.py 1024 > gpurun_out/r2_sharded_host_probe.log 2>&1
timeout 300 python scripts/sharded_host_probe.py 256 >> gpurun_out/r2_sharded_host_probe.log 2>&1
tail -4 gpurun_out/r2_pytest_radial2.log; head -4 gpurun_out/r2_f_rows_probe_after2.log; cat gpurun_out/r2_shard_static_sweep2.log gpurun_out/r2_sharded_host_probe.log
