#!/bin/bash
# round 2, GPU call 13: flush blocks of 8 planes against 16, and the static share of the span schedule on small grids
mkdir -p gpurun_out
out=gpurun_out/r2_ab13.log; : > $out
export IWB_ALLOW_LIB_OVERRIDE=1
for lib in libiwb200.so libiwb200_fl8.so; do
  for pct in auto 90 100; do
    echo "=== $lib static_pct=$pct" >> $out
    if [ $pct = auto ]; then unset IWB_K3_STATIC_PCT; else export IWB_K3_STATIC_PCT=$pct; fi
    for n in 1024 512 400 256 200 128; do IWB_LIB=$lib python scripts/k3_probe.py $n 6 2>&1 | cut -c1-115 >> $out; done
  done
done
unset IWB_K3_STATIC_PCT
for lib in libiwb200.so libiwb200_fl8.so; do echo "=== $lib slab" >> $out; IWB_LIB=$lib python scripts/k3_probe.py 1024 4 float64 0 128 2>&1 | cut -c1-115 >> $out; done
cat $out
