#!/bin/bash
# frame 4096 / hop 1024 at 8192 channels (BASELINE config 5 row): ring-order kernel vs the CTA kernel
cd "$(dirname "$0")/.."
line() { python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$1', r['kernel'], 'us/launch %.2f' % r['avg_launch_us'], 'frac %.4f' % r['frac'], 'frames/s %.3e' % d['value'])"; }
B="python bench.py --frame 4096 --channels 8192 --steps 600 --warmup 100 --no-cpu-baseline --no-e2e --no-other-configs"
for pf in 1.2 0.8; do
  $B --pitch $pf 2>/dev/null | line "4096/1024 pf $pf ring"
  PVB_RING_4096=0 $B --pitch $pf 2>/dev/null | line "4096/1024 pf $pf cta"
done
