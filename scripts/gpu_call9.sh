#!/bin/bash
# round 2, GPU call 9: full GPU suite on the build with the preparation cache, bench, per-kernel times of the radial profile,
# host cost of the sharded API
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_9.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench1_9.json 2> gpurun_out/r2_bench1_9.err
timeout 300 python scripts/sharded_host_probe.py 1024 > gpurun_out/r2_sharded_host_probe2.log 2>&1
timeout 300 python scripts/sharded_host_probe.py 256 >> gpurun_out/r2_sharded_host_probe2.log 2>&1
timeout 300 python scripts/e2e_probe.py 1024 > gpurun_out/r2_e2e_probe3.log 2>&1
timeout 300 python scripts/e2e_probe.py 256 >> gpurun_out/r2_e2e_probe3.log 2>&1
python scripts/radial_probe.py 1024 > gpurun_out/r2_radial_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_radial.csv python scripts/radial_probe.py 1024 > gpurun_out/r2_ncu_radial.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_9.log | cut -c1-200
tail -3 gpurun_out/r2_bench1_9.err; cut -c1-250 gpurun_out/r2_bench1_9.json
cat gpurun_out/r2_sharded_host_probe2.log gpurun_out/r2_e2e_probe3.log gpurun_out/r2_radial_plain.log
grep -E "k_radial|k_energy3d" gpurun_out/r2_launches_radial.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | tail -24
