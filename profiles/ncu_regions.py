#!/usr/bin/env python
"""Per-code-region view of an .ncu-rep (SASS level): executed instructions, stall samples by
reason, in buckets of consecutive SASS instructions.  Usage: ncu_regions.py rep [bucket]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 250
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
isrc, iex, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
names = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_not_selected",
         "stall_selected", "stall_branch_resolving", "stall_no_inst", "stall_dispatch", "stall_lg", "stall_barrier"]
idx = [hdr.index(n) for n in names]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        data.append((r[isrc], int(r[iex]), int(r[isamp]), [int(r[i]) for i in idx]))
    except ValueError:
        pass
te = sum(d[1] for d in data); ts = sum(d[2] for d in data)
warps = max(d[1] for d in data[:50]) or 1
print(f"static SASS instructions {len(data)}; executed warp-instr {te} (per warp {te / warps:.0f}); samples {ts}")
print("stall totals:", {n[6:]: sum(d[3][k] for d in data) for k, n in enumerate(names)})
print("bucket  exec%  samples  " + " ".join(n[6:10] for n in names))
for b in range(0, len(data), bucket):
    seg = data[b:b + bucket]
    st = [sum(d[3][k] for d in seg) for k in range(len(names))]
    print(f"{b:6d} {100 * sum(d[1] for d in seg) / te:5.1f}% {sum(d[2] for d in seg):7d}  " + " ".join(f"{v:4d}" for v in st))
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        for i, d in enumerate(data):
            f.write(f"{i} {d[1]:8d} {d[2]:4d} {d[0]}\n")
