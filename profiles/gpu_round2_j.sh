#!/bin/bash
# ncu --set full capture of the headline kernel, summarised per phase on the box
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none $NCUARGS"
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --no-batched $BARGS"
timeout 300 $NCU --set full --import-source on -k regex:pv_process_ring -s 60 -c 1 -f -o $O/j_ncu python bench.py $B > $O/j_ncu.out 2>&1
(python profiles/ncu_summary.py $O/j_ncu.ncu-rep 2.0 | head -40; python profiles/ncu_phases.py $O/j_ncu.ncu-rep) > $O/j_ncu.txt 2>&1
rm -f $O/j_ncu.ncu-rep
cat $O/j_ncu.txt; tail -n 3 $O/j_ncu.out
