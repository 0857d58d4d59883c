#!/usr/bin/env python
"""Stall samples and executed instructions of one captured ring-kernel launch per PHASE of ring_one_call.
The ncu source page (SASS) gives samples per instruction in program order; profiles/sass_phases.py gives the
phase of every instruction of the same function from the line tables of the object the library was linked
from.  Usage: python profiles/ncu_phases.py rep.ncu-rep [object] [mangled-name substring]"""
import collections, csv, io, os, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
rep = sys.argv[1]
obj = sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, "..", "phaze_b200/csrc/build/ring_1024.o")
want = sys.argv[3] if len(sys.argv) > 3 else "ILi1024ELi2ELb0ELb0ELi0EE"
with tempfile.NamedTemporaryFile("r", suffix=".seq") as tf:
    subprocess.run([sys.executable, os.path.join(HERE, "sass_phases.py"), obj, want], check=True,
                   env=dict(os.environ, SASS_PHASES_SEQ=tf.name), stdout=subprocess.DEVNULL)
    seq = [l.rstrip("\n").split("\t") for l in open(tf.name)]
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
if len(data) != len(seq):
    print(f"warning: {len(data)} instructions in the report, {len(seq)} in the object: phases by position may be off")
warps = max(d[1] for d in data[:50]) or 1
agg = collections.OrderedDict()
for i, d in enumerate(data):
    ph = seq[i][0] if i < len(seq) else "?"
    a = agg.setdefault(ph, [0, 0, [0] * len(names)])
    a[0] += d[1]; a[1] += d[2]
    for k in range(len(names)):
        a[2][k] += d[3][k]
te = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"executed warp-instructions per warp {te / warps:.0f}; samples {ts}")
print(f"{'phase':22s} {'inst/warp':>9s} {'inst%':>6s} {'samp%':>6s}  " + " ".join(n[6:10] for n in names))
for ph, a in agg.items():
    if a[0] == 0 and a[1] == 0:
        continue
    print(f"{ph:22s} {a[0] / warps:9.0f} {100 * a[0] / te:6.1f} {100 * a[1] / ts:6.1f}  " + " ".join(f"{v:4d}" for v in a[2]))
