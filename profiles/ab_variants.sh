#!/bin/bash
# A/B of library builds of the ring-order kernel (`make -C phaze_b200/csrc variants`; macros in
# phaze_b200/csrc/pv_kernel_ring.cuh), selected at run time with PVB_LIBRARY:
#   libphaze_b200_p8.so         -DPVB_RING_PAIRS_1024=8   (16 warps per SM with PVB_RING_WPC=8)
#   libphaze_b200_lanefence.so  -DPVB_RING_LANE_FENCE=1   (a fence in every thread before the release store;
#                                                          the first version, 1 % slower)
# Measured in round 1 (4096 channels, pitch factor 0.8): default 16.28 us with the release by thread 0
# only against 16.44 with the lane fence; 8 pairs per CTA 16.41 against 16.44 with 7 (same build).
cd "$(dirname "$0")/.."
L=$PWD/phaze_b200
line() { python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$1', 'us/launch %.2f' % r['avg_launch_us'], 'frac %.4f' % r['frac'], 'frames/s %.3e' % d['value'], 'l2res %.3e' % (d.get('l2_resident_value') or 0))"; }
B="python bench.py --steps ${STEPS:-3000} --warmup 300 --no-cpu-baseline --no-e2e --no-other-configs"
$B 2>/dev/null | line "default (7 pairs, release by thread 0)"
PVB_LIBRARY=$L/libphaze_b200_lanefence.so $B 2>/dev/null | line "lane fence"
PVB_LIBRARY=$L/libphaze_b200_p8.so PVB_RING_WPC=7 $B 2>/dev/null | line "p8 lib, 7 pairs per CTA"
PVB_LIBRARY=$L/libphaze_b200_p8.so PVB_RING_WPC=8 $B 2>/dev/null | line "p8 lib, 8 pairs per CTA"
PVB_LIBRARY=$L/libphaze_b200_p8.so PVB_RING_WPC=8 $B --channels 32768 --pitch 1.25 2>/dev/null | line "p8 lib, 8 pairs, 32768 ch pf 1.25"
$B --channels 32768 --pitch 1.25 2>/dev/null | line "default, 32768 ch pf 1.25"
