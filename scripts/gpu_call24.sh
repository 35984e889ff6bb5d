#!/bin/bash
# round 2, GPU call 24 (1 GPU): the two seeded random-configuration tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "random" 2>&1 | tail -15 | cut -c1-250
