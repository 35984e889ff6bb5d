#!/bin/bash
# quick A/B of kernel builds (fp64 exact, n=1024^3 and the 128-plane slab of an 8-GPU run): scripts/ab_quick.sh lib1.so lib2.so ...
out=gpurun_out/ab_quick.log
: > $out
for lib in "$@"; do
  echo "=== $lib" >> $out
  IWB_ALLOW_LIB_OVERRIDE=1 IWB_LIB=$lib python scripts/k3_probe.py 1024 5 float64 0 128 >> $out 2>&1
done
cat $out
