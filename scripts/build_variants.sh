#!/bin/bash
# builds libiwb200_<name>.so for every "name:-DFLAG=..;-DFLAG2=.." argument, in parallel (development aid, see build.py)
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}; flags=${flags//;/ }
  ( IWB_VARIANT=$name IWB_EXTRA_FLAGS="$flags" python irrotational-warp-lab_b200/build.py > /tmp/build_$name.log 2>&1 && echo "built $name [$flags]" || { echo "FAILED $name"; tail -5 /tmp/build_$name.log; } ) &
done
wait
