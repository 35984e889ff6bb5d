#!/bin/bash
# round 2, GPU call 29 (1 GPU): time-out of the fused exchange (child process), emulated-rank exchange, full suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_final.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu_final.log | cut -c1-250
