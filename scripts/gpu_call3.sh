#!/bin/bash
# round 2, GPU call 3: A/B of tally sub-options and the mbarrier neighbour sync, GPU test-suite (large goldens, emulated
# peer exchange), bench line, ncu capture of the current default kernel
mkdir -p gpurun_out
scripts/ab2.sh gpurun_out/r2_ab3.log libiwb200.so libiwb200_t3.so libiwb200_t4.so libiwb200_nb3.so libiwb200_nb3q.so libiwb200_sr.so > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_3.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python scripts/k3_probe.py 1024 3 > gpurun_out/r2_k3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_energy3d -s 2 -c 1 -o gpurun_out/r2a_k3 python scripts/k3_probe.py 1024 3 > gpurun_out/r2a_k3_ncu.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu_3.log
tail -5 gpurun_out/r2_bench1.err; cut -c1-600 gpurun_out/r2_bench1.json
grep -E "===|n=1024|slab|n=256" gpurun_out/r2_ab3.log | sed -E 's/all=\[.*\] e_pos/e_pos/' | cut -c1-150
