#!/bin/bash
# round 2, GPU call 23 (1 GPU): seeded random configurations against the C oracle
mkdir -p gpurun_out
timeout 900 python scripts/fuzz_axisym.py 60 2027 > gpurun_out/r2_fuzz_axisym.log 2>&1; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k random_configurations 2>&1 | tail -2
grep -c "^ok" gpurun_out/r2_fuzz_axisym.log; grep -v "^ok" gpurun_out/r2_fuzz_axisym.log | cut -c1-230
