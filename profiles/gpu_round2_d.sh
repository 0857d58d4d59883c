#!/bin/bash
# round 2: per-channel pitch factors + guard threshold 5: tests and headline timing
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
(timeout 1200 python -m pytest tests/test_gpu_per_channel_pitch.py tests/test_gpu_peak_guard.py -x -q -s 2>&1 | grep -v "^\.N=\|^N=\|^\.$" | tail -40) > $O/d_new_tests.log
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/d_tests.log
B="--no-cpu-baseline --no-other-configs --no-e2e --steps 1000 --warmup 20"
for G in 0 1; do
  timeout 300 python bench.py $B --peak-guard $G 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('guard=$G', 'us/step', 1e3*d['ms_per_step'], 'frac', d['roofline']['frac'], d['peak_guard']['frames_redecided_in_float64'], d['launch_chaining'])"
done > $O/d_bench.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-other-configs --no-e2e --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('K=20', 'us/step', 1e3*d['ms_per_step'], 'frac', d['roofline']['frac'])" >> $O/d_bench.log 2>&1
cat $O/d_new_tests.log $O/d_tests.log $O/d_bench.log
