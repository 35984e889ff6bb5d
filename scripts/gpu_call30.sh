#!/bin/bash
# round 2, GPU call 30 (1 GPU): soak (repeatability) on the final tree, and 200 + 100 further random configurations
mkdir -p gpurun_out
timeout 900 python scripts/soak.py > gpurun_out/r2_soak_final.log 2>&1; tail -5 gpurun_out/r2_soak_final.log | cut -c1-200
timeout 900 python scripts/fuzz_oracle.py 200 777 > gpurun_out/r2_fuzz_oracle_200.log 2>&1; tail -1 gpurun_out/r2_fuzz_oracle_200.log; grep -c "points differing 0," gpurun_out/r2_fuzz_oracle_200.log; grep "^BAD" gpurun_out/r2_fuzz_oracle_200.log | head -5 | cut -c1-220
timeout 900 python scripts/fuzz_axisym.py 100 778 > gpurun_out/r2_fuzz_axisym_100.log 2>&1; tail -1 gpurun_out/r2_fuzz_axisym_100.log; grep "^BAD" gpurun_out/r2_fuzz_axisym_100.log | head -8 | cut -c1-220
