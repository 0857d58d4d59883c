#!/bin/bash
# Y-plane padding of the ring-order kernel: default build (PVB_RING_YSHIFT=5) vs -DPVB_RING_YSHIFT=4
# (phaze_b200/libphaze_b200_ys4.so), over the configurations the pad matters for.
cd "$(dirname "$0")/.."
L=$PWD/phaze_b200
line() { python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$1', 'us/launch %.2f' % r['avg_launch_us'], 'frac %.4f' % r['frac'], 'frames/s %.3e' % d['value'])"; }
B="python bench.py --steps ${STEPS:-2000} --warmup 300 --no-cpu-baseline --no-e2e --no-other-configs"
for cfg in "--channels 4096 --pitch 0.8" "--channels 8192 --pitch 1.2" "--channels 32768 --pitch 1.25" \
           "--frame 2048 --channels 2048 --pitch 1.5" "--frame 512 --channels 8192 --pitch 1.2" "--frame 512 --channels 8192 --pitch 0.8"; do
  $B $cfg 2>/dev/null | line "ys5 [$cfg]"
  PVB_LIBRARY=$L/libphaze_b200_ys4.so $B $cfg 2>/dev/null | line "ys4 [$cfg]"
done
