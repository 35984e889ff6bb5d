#!/bin/bash
# A/B timing of kernel builds on one box (development aid): scripts/ab_probe.sh lib1.so lib2.so ...
set -u
out=gpurun_out/ab_probe.log
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $out 2>&1
for lib in "$@"; do
  echo "=== $lib" >> $out
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$lib python scripts/k3_probe.py 1024 4 float64 0 128 >> $out 2>&1
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$lib python scripts/k3_probe.py 1024 4 float32 >> $out 2>&1
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$lib IWB_MODE=fast python scripts/k3_probe.py 1024 4 >> $out 2>&1
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$lib IWB_K3_CONFIG=0 python scripts/k3_probe.py 1024 4 >> $out 2>&1
done
cat $out
