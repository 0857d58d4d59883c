#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU), N from $1
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 2>$O/k_bench_$N.err | tail -1 > $O/k_bench_$N.json
python - <<PY
import json
d = json.load(open("$O/k_bench_$N.json"))
print('n', d['n_gpus'], 'value', d['value'], 'per gpu', d['value'] / d['n_gpus'], 'us/launch', d['roofline']['avg_launch_us'], 'frac', d['roofline']['frac'])
print('e2e', d['e2e']['value'], d['e2e'].get('gbs_each_way_per_gpu'), d['e2e'].get('pcie_probe_gbs_each_way_per_gpu'))
print('root_scatter', json.dumps(d.get('root_scatter'))[:900])
PY
tail -n 5 $O/k_bench_$N.err
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5)
