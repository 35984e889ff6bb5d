#!/bin/bash
# round 2, GPU call 24 (1 GPU): random configurations on grids with interior tiles (n = 120..300)
mkdir -p gpurun_out
timeout 900 python scripts/fuzz_oracle.py 12 2028 120 301 > gpurun_out/r2_fuzz_oracle_large.log 2>&1
cat gpurun_out/r2_fuzz_oracle_large.log | cut -c1-230
