#!/bin/bash
# parity tests + a short bench of the current build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > $O/i_tests.log
cat $O/i_tests.log
B="--no-cpu-baseline --no-other-configs --no-e2e --no-batched --steps 20 --warmup 3"
for PF in 0.8 1.2; do
timeout 300 python bench.py $B --pitch $PF 2>$O/i_bench.err | tail -1 > $O/i_bench_$PF.json
python - <<PY
import json
d=json.load(open("$O/i_bench_$PF.json"))
print('pf', $PF, 'us/launch', d['roofline']['avg_launch_us'], 'frac', d['roofline']['frac'], 'chain', {k: round(v['avg_launch_us'],2) for k, v in d['launch_chaining'].items()})
PY
done
tail -n 5 $O/i_bench.err
