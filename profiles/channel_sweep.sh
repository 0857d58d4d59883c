#!/bin/bash
# per-launch time of the default kernel vs channels per launch (one line per count)
for c in ${@:-1024 2048 4096 8192 32768}; do
  python bench.py --channels $c --steps 300 --warmup 50 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('channels', $c, 'us/launch %.2f' % d['roofline']['avg_launch_us'], 'frac %.4f' % d['roofline']['frac'], 'l2_resident %.3e' % (d.get('l2_resident_value') or 0), 'concurrent %.3e' % (d.get('concurrent_streams_value') or 0))"
done
