#!/bin/bash
# compute-sanitizer over the DEEP instances of the ring-order kernel only (pitch factors in [0.5, 0.75)):
# profiles/sanitize_workload.py --deep-only.  racecheck on the product build checks the claim the three coloured
# sub-steps rest on: within one sub-step no two threads touch the same word of the shifted spectrum.
# Output: gpurun_out/sanitizer_deep.txt (copied to profiles/r02_sanitizer_deep.txt).
cd "$(dirname "$0")/.."
O=gpurun_out/sanitizer_deep.txt
mkdir -p gpurun_out
: > $O
for tool in memcheck synccheck racecheck; do
  echo "===== compute-sanitizer --tool $tool (product build, DEEP instances) =====" >> $O
  timeout 1200 compute-sanitizer --tool $tool --print-limit 200 python profiles/sanitize_workload.py --deep-only \
    > gpurun_out/san_raw.txt 2>&1
  grep -E "Race reported|and .* access at" gpurun_out/san_raw.txt \
    | sed -E 's/=========//; s/\+0x[0-9a-f]+//; s/bool pvb::ring_one_call<[^>]*>\([^)]*\)/ring_one_call<..>/; s/\[[0-9]+ hazards\]//' \
    | sort | uniq -c | sort -rn | head -20 >> $O
  grep -E "SUMMARY|WORKLOAD OK|worst rms|Error|error" gpurun_out/san_raw.txt | grep -v "Race reported" | head -12 >> $O
  rm -f gpurun_out/san_raw.txt
done
cat $O
