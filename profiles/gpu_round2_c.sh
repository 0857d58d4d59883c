#!/bin/bash
# round 2: peak guard policies -- parity tests, then the headline bench with the guard at default / off / strict / always
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests/test_gpu_peak_guard.py -x -q -s 2>&1 | grep -v "^\.N=\|^N=" | tail -30) > $O/c_guard_tests.log
B="--no-cpu-baseline --no-other-configs --no-e2e --steps 1000 --warmup 20"
for G in 0 1 3 2; do
  timeout 300 python bench.py $B --peak-guard $G 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('guard=$G', 'us/step', 1e3*d['ms_per_step'], 'frac', d['roofline']['frac'], d['peak_guard']['frames_redecided_in_float64'], d['launch_chaining'])"
done > $O/c_guard_bench.log 2>&1
timeout 300 python bench.py $B --pitch 1.2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pf1.2 guard=0', 'us/step', 1e3*d['ms_per_step'], 'frac', d['roofline']['frac'], d['peak_guard']['frames_redecided_in_float64'])" >> $O/c_guard_bench.log 2>&1
cat $O/c_guard_tests.log $O/c_guard_bench.log
