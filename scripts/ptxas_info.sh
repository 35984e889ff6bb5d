#!/bin/bash
# ptxas resource summary of the headline instantiation of k_energy3d for a set of -D flags (development aid)
# usage: scripts/ptxas_info.sh [-DIWB_X=1 ...]
set -e
out=/tmp/ptxas_$$.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas=-v -fmad=false \
  -I include "$@" -c irrotational-warp-lab_b200/csrc/energy3d.cu -o $out 2>&1 | grep -A3 "k_energy3dILi0ELi2ELi40ELi20ELb0ELi1ELi0E" | grep -E "spill|Used"
cuobjdump -sass -fun '_ZN3iwb10k_energy3dILi0ELi2ELi40ELi20ELb0ELi1ELi0EEEvNS_8K3ParamsE' $out > /tmp/k3_variant.sass
grep -cE "^\s+/\*[0-9a-f]{4,5}\*/" /tmp/k3_variant.sass
rm -f $out
