#!/bin/bash
# round 2: peak guard on the GPU -- parity tests, guard on/off timing, fixed cost of a short timed region
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests/test_gpu_peak_guard.py -x -q -s 2>&1 | tail -60) > $O/b_guard_tests.log
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/b_tests.log
B="--no-cpu-baseline --no-other-configs --no-e2e"
for K in 10 20 40 80 160 1000; do
  timeout 300 python bench.py --steps $K --warmup 5 $B 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('K=$K', 'ms_total', d['ms_per_step']*d['steps'], 'us/step', 1e3*d['ms_per_step'], 'frac', d['roofline']['frac'], d['launch_chaining'])"
done > $O/b_ksweep.log 2>&1
cat $O/b_guard_tests.log $O/b_tests.log $O/b_ksweep.log
