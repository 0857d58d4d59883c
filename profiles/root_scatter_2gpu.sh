#!/bin/bash
# single-root mode on 2 GPUs (bench.py --root-scatter): per-call exchange vs the streamed mode
# (K calls per NCCL message), with NCCL's point-to-point channel count and stream priority varied
cd "$(dirname "$0")/.."
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 50)) \
        bench.py --gpus 2 --steps 1000 --warmup 100 --root-scatter --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['root_scatter']; s = r['streamed']
print('$1', 'resident %.3e' % d['value'], 'per-call %.3e (%.0f GB/s)' % (r['value'], r['nvlink_gbs_at_root']), 'streamed K=%d %.3e (%.0f GB/s at the root)' % (s['calls_per_message'], s['value'], s['nvlink_gbs_at_root']))"; }
run "default"
NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32 run "p2p-channels-16..32"
TORCH_NCCL_HIGH_PRIORITY=1 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32 run "high-priority+channels"
PVB_BENCH_STREAM_CALLS=64 TORCH_NCCL_HIGH_PRIORITY=1 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32 run "K=64 high-priority+channels"
PVB_BENCH_STREAM_CALLS=4 TORCH_NCCL_HIGH_PRIORITY=1 run "K=4 high-priority"
