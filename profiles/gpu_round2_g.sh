#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > $O/g_tests.log
B="--no-cpu-baseline --no-other-configs --no-e2e --steps 1000 --warmup 20"
timeout 300 python bench.py $B 2>$O/g_bench.err | tail -1 > $O/g_bench.json
python - <<PY >> $O/g_tests.log
import json
d=json.load(open("$O/g_bench.json"))
print('us/step', 1e3*d['ms_per_step'], 'frac', d['roofline']['frac'], 'batched', d['batched'])
PY
cat $O/g_tests.log; tail -5 $O/g_bench.err
(time timeout 300 python profiles/sanitize_workload.py) 2>&1 | tail -12
bash profiles/sanitize.sh > /dev/null 2>&1; cat $O/sanitizer.txt
