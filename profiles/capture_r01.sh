#!/bin/bash
# Round-end evidence on one B200: GPU tests, smoke, both bench arms, the config table, the ncu launch
# list of the bench command and one `ncu --set full` capture per frame size of the ring-order kernel.
# Outputs go to gpurun_out/ (summaries are produced from the .ncu-rep files with ncu_summary.py /
# ncu_regions.py and committed under profiles/).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
[ -n "$SKIP_TESTS" ] || (timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $O/final_tests.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $O/final_smoke.log
timeout 300 python bench.py 2>$O/bench_1gpu.err | tail -1 > $O/bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>$O/bench_ref.err | tail -1 > $O/bench_ref.json
[ -n "$SKIP_TESTS" ] || timeout 300 bash profiles/sweep_configs.sh > $O/sweep_configs.log 2>&1
NCU="ncu --clock-control none"
B="--steps 40 --warmup 10 --no-cpu-baseline --no-e2e --no-other-configs"
timeout 200 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches.csv python bench.py $B > $O/launches.out 2>&1
cap() { # name, bench args: capture, summarise on the box (the reports are 13 MB each), keep the text
  timeout 200 $NCU --set full --import-source on -k regex:pv_process_ring -s 30 -c 1 -f -o $O/$1 python bench.py $B $2 > $O/$1.out 2>&1
  (python profiles/ncu_summary.py $O/$1.ncu-rep 2.0; python profiles/ncu_regions.py $O/$1.ncu-rep 300) > $O/$1.txt 2>&1
  [ "$1" = ncu_ring_1024 ] || rm -f $O/$1.ncu-rep
}
cap ncu_ring_1024 ""
cap ncu_ring_2048 "--frame 2048 --channels 2048 --pitch 1.5"
cap ncu_ring_512 "--frame 512 --channels 8192 --pitch 1.2"
cap ncu_ring_256 "--frame 256 --channels 8192 --pitch 1.2"
cap ncu_ring_4096 "--frame 4096 --channels 8192 --pitch 1.2"
ls -la $O | tail -20
cat $O/final_tests.log $O/final_smoke.log $O/bench_1gpu.json $O/bench_ref.json $O/sweep_configs.log 2>/dev/null
