#!/bin/bash
# frame 2048 configurations (config 3, the reference-native 2048/128, the 8192-channel sweep point)
for cfg in "2048 512 2048 1.5" "2048 128 2048 1.2" "2048 128 2048 0.8" "2048 512 8192 1.2"; do
  set -- $cfg
  python bench.py --frame $1 --hop $2 --channels $3 --pitch $4 --steps 300 --warmup 30 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%s hop=%s C=%s pf=%s: %.3e frames/s  %.1f us/launch  frac %.3f  kernel %s  l2res %.3e' % ('$1','$2','$3','$4', d['value'], r['avg_launch_us'], r['frac'], r['kernel'], d['l2_resident_value']))"
done
