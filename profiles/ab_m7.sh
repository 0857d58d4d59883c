#!/bin/bash
# MULTI / DEEP instances of frame 1024: 6 pairs per CTA at 168 registers (default) against 7 at 144 (make m7)
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 20 --warmup 3"
for V in "" _m7; do
  L=phaze_b200/libphaze_b200$V.so
  PVB_LIBRARY=$L $B 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lib$V', 'us/launch', round(d['roofline']['avg_launch_us'],2), 'batched us/call', round(d['batched']['us_per_call'],2), 'vs per-call', round(d['batched']['vs_one_launch_per_call'],3))"
  for pf in 0.74 0.6 0.5; do
    PVB_LIBRARY=$L $B --no-batched --pitch $pf 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lib$V pitch $pf', 'us/launch', round(d['roofline']['avg_launch_us'],2))"
  done
done
