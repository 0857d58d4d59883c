#!/bin/bash
# exact first-writer classification (-DPVB_RING_EXACT=1, phaze_b200/libphaze_b200_exact.so) vs the
# default build, contracting pitch factors (the expanding path is unchanged)
cd "$(dirname "$0")/.."
L=$PWD/phaze_b200
line() { python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$1', 'us/launch %.2f' % r['avg_launch_us'], 'frac %.4f' % r['frac'], 'frames/s %.3e' % d['value'])"; }
B="python bench.py --steps ${STEPS:-2000} --warmup 300 --no-cpu-baseline --no-e2e --no-other-configs"
for cfg in "--channels 4096 --pitch 0.8" "--channels 32768 --pitch 0.9" "--frame 2048 --channels 2048 --pitch 0.8" \
           "--frame 512 --channels 8192 --pitch 0.8" "--frame 256 --channels 8192 --pitch 0.8" "--frame 4096 --channels 8192 --pitch 0.8 --steps 500"; do
  $B $cfg 2>/dev/null | line "default [$cfg]"
  PVB_LIBRARY=$L/libphaze_b200_exact.so $B $cfg 2>/dev/null | line "exact   [$cfg]"
done
