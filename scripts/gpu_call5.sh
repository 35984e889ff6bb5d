#!/bin/bash
# round 2, GPU call 5: validation of the cleaned-up build -- full GPU suite, bench both arms with the driver's flags, launch
# list, rho-emission / f-row probes, final ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_5.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench1_final.json 2> gpurun_out/r2_bench1_final.err
timeout 300 python scripts/emit_probe.py 1024 > gpurun_out/r2_emit_probe.log 2>&1
timeout 600 python scripts/f_rows_probe.py > gpurun_out/r2_f_rows_probe.log 2>&1
timeout 300 python scripts/e2e_probe.py 1024 > gpurun_out/r2_e2e_probe2.log 2>&1
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_plain_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
python scripts/k3_probe.py 1024 3 > gpurun_out/r2_k3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_energy3d -s 2 -c 1 -o gpurun_out/r2b_k3 python scripts/k3_probe.py 1024 3 > gpurun_out/r2b_k3_ncu.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu_5.log | cut -c1-200
tail -3 gpurun_out/r2_bench1_final.err; cut -c1-300 gpurun_out/r2_bench1_final.json; cut -c1-300 gpurun_out/r2_bench_ref_final.json
cat gpurun_out/r2_emit_probe.log gpurun_out/r2_f_rows_probe.log gpurun_out/r2_e2e_probe2.log gpurun_out/r2_k3_plain.log
