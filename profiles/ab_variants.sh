#!/bin/bash
# A/B of library builds of the ring-order kernel (see phaze_b200/csrc/pv_kernel_ring.cuh macros):
#   libphaze_b200_p8.so   -DPVB_RING_PAIRS_1024=8            (16 warps per SM with PVB_RING_WPC=8)
#   libphaze_b200_nf.so   -DPVB_RING_LANE_FENCE=0            (release by thread 0 only; the default since)
#   libphaze_b200_p8nf.so both
# (`make -C phaze_b200/csrc variants` builds today's set: p8, lanefence, ys4, exact)
# usage: profiles/ab_variants.sh   (prints one line per build / setting)
cd "$(dirname "$0")/.."
L=$PWD/phaze_b200
line() { python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$1', 'us/launch %.2f' % r['avg_launch_us'], 'frac %.4f' % r['frac'], 'frames/s %.3e' % d['value'], 'l2res %.3e' % (d.get('l2_resident_value') or 0))"; }
B="python bench.py --steps ${STEPS:-3000} --warmup 300 --no-cpu-baseline --no-e2e"
$B 2>/dev/null | line "base(7 pairs, lane fence)"
PVB_LIBRARY=$L/libphaze_b200_nf.so $B 2>/dev/null | line "nf(7 pairs, no lane fence)"
PVB_LIBRARY=$L/libphaze_b200_p8.so PVB_RING_WPC=7 $B 2>/dev/null | line "p8 lib, 7 pairs per CTA"
PVB_LIBRARY=$L/libphaze_b200_p8.so PVB_RING_WPC=8 $B 2>/dev/null | line "p8 lib, 8 pairs per CTA"
PVB_LIBRARY=$L/libphaze_b200_p8nf.so PVB_RING_WPC=8 $B 2>/dev/null | line "p8nf lib, 8 pairs per CTA"
PVB_LIBRARY=$L/libphaze_b200_p8nf.so PVB_RING_WPC=8 $B --channels 32768 --pitch 1.25 2>/dev/null | line "p8nf, 8 pairs, 32768 ch pf 1.25"
$B --channels 32768 --pitch 1.25 2>/dev/null | line "base, 32768 ch pf 1.25"
