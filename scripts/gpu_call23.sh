#!/bin/bash
# round 2, GPU call 23 (1 GPU): seeded random configurations against the C oracle
mkdir -p gpurun_out
timeout 900 python scripts/fuzz_oracle.py 60 2026 > gpurun_out/r2_fuzz_oracle.log 2>&1
tail -70 gpurun_out/r2_fuzz_oracle.log | cut -c1-230
