#!/bin/bash
# compute-sanitizer over the ring-order kernel family on profiles/sanitize_workload.py:
# memcheck, synccheck (sub-warp __syncwarp masks, bar.sync with thread counts), initcheck, and racecheck
# (shared-memory hazards between the sub-steps of the shift, the exchange buffers, the named barriers).
# racecheck runs twice: on the product build, whose only hazards are write-after-write on the write-only
# "dump" word that out-of-range destinations are clamped to (see PVB_RING_NO_DUMP in pv_kernel_ring.cuh),
# and on the -DPVB_RING_NO_DUMP=1 build (make nodump), which predicates those stores off and must be clean.
# Output: gpurun_out/sanitizer.txt (copied to profiles/r02_sanitizer.txt).
cd "$(dirname "$0")/.."
O=gpurun_out/sanitizer.txt
mkdir -p gpurun_out
: > $O
for tool in memcheck synccheck initcheck; do
  echo "===== compute-sanitizer --tool $tool (product build) =====" >> $O
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_workload.py 2>&1 \
    | grep -v "^frame " | tail -8 >> $O
done
race() { # label, library
  echo "===== compute-sanitizer --tool racecheck ($1) =====" >> $O
  PVB_LIBRARY=$2 timeout 2400 compute-sanitizer --tool racecheck --print-limit 100000 python profiles/sanitize_workload.py \
    > gpurun_out/racecheck_raw.txt 2>&1
  echo "hazards by kind and source line (all reports, not just the first):" >> $O
  grep -E "Race reported|and .* access at" gpurun_out/racecheck_raw.txt \
    | sed -E 's/=========//; s/\+0x[0-9a-f]+//; s/bool pvb::ring_one_call<[^>]*>\([^)]*\)/ring_one_call<..>/; s/\[[0-9]+ hazards\]//' \
    | sort | uniq -c | sort -rn | head -40 >> $O
  grep -E "RACECHECK SUMMARY|WORKLOAD OK|worst rms" gpurun_out/racecheck_raw.txt >> $O
  rm -f gpurun_out/racecheck_raw.txt
}
race "product build" phaze_b200/libphaze_b200.so
race "-DPVB_RING_NO_DUMP=1 build" phaze_b200/libphaze_b200_nodump.so
cat $O
