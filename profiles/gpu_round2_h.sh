#!/bin/bash
# bench at the driver's settings and at the defaults (step = 64 calls), side by side
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs 2>$O/h_driver.err | tail -1 > $O/h_driver.json
timeout 600 python bench.py --no-other-configs --no-cpu-baseline 2>$O/h_long.err | tail -1 > $O/h_long.json
python - <<PY
import json
for n in ("h_driver", "h_long"):
    d = json.load(open("$O/%s.json" % n))
    print(n, 'steps', d['steps'], 'ms/step', d['ms_per_step'], 'us/launch', d['roofline']['avg_launch_us'], 'frac', d['roofline']['frac'],
          'e2e', d['e2e']['value'], 'chain', {k: v['avg_launch_us'] for k, v in d['launch_chaining'].items()})
PY
tail -n 3 $O/h_driver.err; tail -n 3 $O/h_long.err
