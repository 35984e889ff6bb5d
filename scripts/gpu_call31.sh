#!/bin/bash
# round 2, GPU call 31 (1 GPU): k_energy3d without the carried flush_from variable (one live register less in the march) against the default
mkdir -p gpurun_out
D=irrotational-warp-lab_b200
IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$D/libiwb200_noff.so timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or slab or shard or odd_grid or full_size" 2>&1 | tail -2 | cut -c1-200
bash scripts/ab2.sh gpurun_out/r2_ab31_noff.log $D/libiwb200.so $D/libiwb200_noff.so > /dev/null 2>&1
grep -v "^name\|^NVIDIA" gpurun_out/r2_ab31_noff.log | cut -c1-150
