#!/bin/bash
# round 2, first GPU call: API / bench changes only (kernels unchanged): tests, driver-style bench, default bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo skipped-tests > $O/a_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs 2>$O/a_bench20.err | tail -1 > $O/a_bench20.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e 2>>$O/a_bench20.err | tail -1 > $O/a_bench20b.json
timeout 400 python bench.py --no-cpu-baseline --no-other-configs --no-e2e 2>$O/a_bench.err | tail -1 > $O/a_bench.json
cat $O/a_tests.log
for f in a_bench20 a_bench20b a_bench; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json"))
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["launch_chaining"], d.get("e2e") and d["e2e"]["value"], d["clocks"])
except Exception as e:
    print("$f failed", e); print(open("$O/$f.err").read()[-2000:] if "$f"!="a_bench20b" else "")
PY
done
