#!/bin/bash
# round 2, GPU call 21 (1 GPU): interleaved shared-memory column layout (a lane's two columns adjacent, 128-bit accesses) --
# parity of the variant (whole parity file through the variant library), A/B timing against the default
mkdir -p gpurun_out
D=irrotational-warp-lab_b200
IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$D/libiwb200_il.so timeout 900 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r2_pytest_interleave.log 2>&1
tail -12 gpurun_out/r2_pytest_interleave.log | cut -c1-220
bash scripts/ab2.sh gpurun_out/r2_ab21_interleave.log $D/libiwb200.so $D/libiwb200_il.so > /dev/null 2>&1
grep -v "^name\|^NVIDIA" gpurun_out/r2_ab21_interleave.log | cut -c1-180
