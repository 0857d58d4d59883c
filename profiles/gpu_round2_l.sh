#!/bin/bash
# ncu --set full capture of one ring-kernel instance, summarised per phase: $1 frame, $2 NBLK, $3 bench args
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none $NCUARGS"
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --no-batched $3"
timeout 300 $NCU --set full --import-source on -k regex:pv_process_ring -s 60 -c 1 -f -o $O/l_ncu python bench.py $B > $O/l_ncu.out 2>&1
(python profiles/ncu_summary.py $O/l_ncu.ncu-rep 2.0 | head -36; python profiles/ncu_phases.py $O/l_ncu.ncu-rep phaze_b200/csrc/build/ring_$1.o "ILi$1ELi$2ELb0ELb0ELi0EE") > $O/l_ncu_$1.txt 2>&1
rm -f $O/l_ncu.ncu-rep
cat $O/l_ncu_$1.txt; tail -n 2 $O/l_ncu.out | cut -c1-300
