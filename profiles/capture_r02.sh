#!/bin/bash
# Round-2 evidence on one B200: GPU tests, smoke, both bench arms, the ncu launch list of the bench
# command and one `ncu --set full` capture per frame size of the ring-order kernel.  Outputs go to
# gpurun_out/ (summaries are produced on the box with ncu_summary.py / ncu_regions.py and committed
# under profiles/ as r02_*).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
[ -n "$SKIP_TESTS" ] || (timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $O/r02_tests.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $O/r02_smoke.log
timeout 600 python bench.py 2>$O/r02_bench_1gpu.err | tail -1 > $O/r02_bench_1gpu.json
timeout 300 python bench.py --steps 20 --warmup 5 2>>$O/r02_bench_1gpu.err | tail -1 > $O/r02_bench_1gpu_driver_cmd.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>$O/r02_bench_ref.err | tail -1 > $O/r02_bench_reference_arm.json
NCU="ncu --clock-control none"
B="--steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --no-batched"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches.csv python bench.py $B > $O/r02_launches.out 2>&1
cap() { # name, bench args: capture, summarise on the box (the reports are 13 MB each), keep the text
  timeout 300 $NCU --set full --import-source on -k regex:pv_process_ring -s 60 -c 1 -f -o $O/$1 python bench.py $B $2 > $O/$1.out 2>&1
  (python profiles/ncu_summary.py $O/$1.ncu-rep 2.0; python profiles/ncu_regions.py $O/$1.ncu-rep 300; [ -n "$3" ] && python profiles/ncu_phases.py $O/$1.ncu-rep $3) > $O/$1.txt 2>&1
  rm -f $O/$1.ncu-rep
}
cap r02_ncu_ring_1024 "" "phaze_b200/csrc/build/ring_1024.o ILi1024ELi2ELb0ELb0ELi0EE"
cap r02_ncu_ring_2048 "--frame 2048 --channels 2048 --pitch 1.5" "phaze_b200/csrc/build/ring_2048.o ILi2048ELi4ELb0ELb0ELi0EE"
cap r02_ncu_ring_1024_deep "--pitch 0.6" "phaze_b200/csrc/build/ring_1024.o ILi1024ELi2ELb0ELb0ELi1EE"
cap r02_ncu_ring_512 "--frame 512 --channels 8192 --pitch 1.2"
cap r02_ncu_ring_256 "--frame 256 --channels 8192 --pitch 1.2"
cap r02_ncu_ring_4096 "--frame 4096 --channels 8192 --pitch 1.2"
ls -la $O | tail -20
cat $O/r02_tests.log $O/r02_smoke.log; head -c 3000 $O/r02_bench_1gpu.json; echo; cat $O/r02_bench_1gpu_driver_cmd.json | head -c 1500; echo; cat $O/r02_bench_reference_arm.json; tail -3 $O/r02_bench_1gpu.err
