#!/bin/bash
# round 2: several calls per launch (MULTI kernels): tests and the batched bench entry
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > $O/e_tests.log
B="--no-cpu-baseline --no-other-configs --no-e2e --steps 1000 --warmup 20"
timeout 300 python bench.py $B 2>$O/e_bench.err | tail -1 > $O/e_bench.json
python - <<PY >> $O/e_tests.log
import json
d=json.load(open("$O/e_bench.json"))
print('us/step', 1e3*d['ms_per_step'], 'frac', d['roofline']['frac'], 'batched', d['batched'])
PY
cat $O/e_tests.log; tail -5 $O/e_bench.err
