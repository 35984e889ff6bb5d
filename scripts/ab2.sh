#!/bin/bash
# A/B of kernel builds on one box at n = 1024 (+ the 128-plane slab of an 8-GPU run), n = 256 and n = 400: scripts/ab2.sh out.log lib1.so lib2.so ...
out=$1; shift
: > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $out 2>&1
for rep in 1 2; do
for lib in "$@"; do
  echo "=== $lib (pass $rep)" >> $out
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$lib python scripts/k3_probe.py 1024 5 float64 0 128 2>&1 | cut -c1-230 >> $out
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$lib python scripts/k3_probe.py 256 8 float64 2>&1 | cut -c1-110 >> $out
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$lib python scripts/k3_probe.py 400 6 float64 2>&1 | cut -c1-110 >> $out
done
done
cat $out
