#!/bin/bash
# round 2, N GPUs of one box: NCCL / multi-device tests, then the bench arm at N ranks (driver-style and long)
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -8
(timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s 2>&1 | tail -8) > $O/multi_tests_$N.log
cat $O/multi_tests_$N.log
run() { # label, bench args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N $2 2>$O/multi_bench_$N.err | tail -1 > $O/multi_bench_${N}_$1.json
  python - <<PY
import json
try:
    d=json.load(open("$O/multi_bench_${N}_$1.json"))
    rs=d.get("root_scatter") or {}
    print("$1 N=$N value", d["value"], "us/step", 1e3*d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"] and d["e2e"]["value"], d["e2e"] and d["e2e"].get("pcie_probe_gbs_each_way_per_gpu"))
    print("   root_scatter", {k: rs.get(k) for k in ("value","nvlink_gbs_at_root")}, "streamed", rs.get("streamed"))
except Exception as e:
    print("$1 failed", e); print(open("$O/multi_bench_$N.err").read()[-3000:])
PY
}
run driver "--steps 20 --warmup 5"
run long "--steps 1000 --warmup 20 --no-e2e"
