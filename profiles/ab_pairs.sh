#!/bin/bash
# pairs (warps) per CTA of the frame-1024 ring-order kernel: 7 (128 registers, 14 warps per SM, default) against
# builds with 8 / 9 / 10 (128 / 113 / 96 registers, 16 / 18 / 20 warps per SM; make pairs)
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --no-batched --steps 1000 --warmup 20"
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'us/step', round(1e3*d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4), 'pdl_grid', round(d['launch_chaining']['pdl_grid_wait']['avg_launch_us'],2))"; }
for cfg in "" "--pitch 1.2" "--channels 32768 --pitch 1.25 --steps 200"; do
  $B $cfg 2>/dev/null | tail -1 | line "default(7) [$cfg]"
  for P in 8 9 10; do
    PVB_LIBRARY=phaze_b200/libphaze_b200_p$P.so $B $cfg 2>/dev/null | tail -1 | line "pairs=$P [$cfg]"
  done
done
