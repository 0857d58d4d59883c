#!/bin/bash
# Where the time of the headline kernel goes: the exp build (make experiments) with parts of the middle
# switched off (PVB_SKIP bits; RESULTS ARE WRONG in these runs, only the time matters):
#   1 whole middle   2 forward + inverse pass 2   4 second sub-step (read-add-store)   8 zero fill
#   16 first sub-step stores   32 stale-bin reconstruction
cd "$(dirname "$0")/.."
B="python bench.py --no-cpu-baseline --no-other-configs --no-e2e --no-batched --steps 600 --warmup 20"
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'us/step', round(1e3*d['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4))"; }
for pf in 0.8 1.2; do
for sk in 0 4 8 12 16 28 32 60 1 2 3; do
  PVB_LIBRARY=phaze_b200/libphaze_b200_exp.so PVB_SKIP=$sk $B --pitch $pf 2>/dev/null | tail -1 | line "pf=$pf skip=$sk guard=default"
done
PVB_LIBRARY=phaze_b200/libphaze_b200_exp.so PVB_SKIP=0 $B --pitch $pf --peak-guard 1 2>/dev/null | tail -1 | line "pf=$pf skip=0 guard=off"
done
