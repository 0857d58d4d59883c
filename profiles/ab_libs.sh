#!/bin/bash
# A/B of library builds at the headline workload: bash profiles/ab_libs.sh <lib suffix> ... ("" = default build)
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --no-batched --steps 20 --warmup 3"
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'us/launch', round(d['roofline']['avg_launch_us'],2), 'frac', round(d['roofline']['frac'],4), 'pdl_grid', round(d['launch_chaining']['pdl_grid_wait']['avg_launch_us'],2))"; }
for cfg in "" "--pitch 1.2" ${ABCFG:+"$ABCFG"}; do
  for V in default "$@"; do
    if [ "$V" = default ]; then L=phaze_b200/libphaze_b200.so; else L=phaze_b200/libphaze_b200_$V.so; fi
    [ -f $L ] || { echo "missing $L"; continue; }
    PVB_LIBRARY=$L $B $cfg 2>/dev/null | tail -1 | line "$V [$cfg]"
  done
done
