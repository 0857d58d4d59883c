#!/usr/bin/env python
"""Static SASS instruction count of one ring-kernel instance per PHASE of ring_one_call (pv_kernel_ring.cuh),
from `nvdisasm --print-line-info` of the object (no GPU needed).  Every instruction is attributed to the
outermost pv_kernel_ring.cuh line of its inline chain; phases are line ranges found by their marker comments.

    python profiles/sass_phases.py [object] [mangled-name substring]
"""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "phaze_b200/csrc/build/ring_1024.o")
want = sys.argv[2] if len(sys.argv) > 2 else "ILi1024ELi2ELb0ELb0ELi0EE"
src = os.path.join(ROOT, "phaze_b200/csrc/pv_kernel_ring.cuh")

MARKS = [("tables+flags", "CTA-shared tables: asynchronous"), ("frame loads", "---- frame loads"),
         ("window+fwd1", "---- Hann window"), ("fwd2", "---- forward pass 2"), ("fwd3", "---- forward pass 3"),
         ("split", "---- real split in registers"), ("middle", "---- peaks, regions of influence"),
         ("unsplit", "---- Hermitian C2R pre-pass in registers"), ("inv1", "---- inverse pass 1"),
         ("acc loads", "accumulator values (L2 hits"), ("inv2", "---- inverse pass 2"), ("inv3+ola", "---- inverse pass 3"),
         ("epilogue", "#undef PVB_NOT_DUMP")]
lines = open(src).read().split("\n")
bounds = []
for name, mark in MARKS:
    for i, l in enumerate(lines):
        if mark in l:
            bounds.append((i + 1, name))
            break
    else:
        raise SystemExit(f"marker not found: {mark}")
sub = [("mid:mags+masks", "contracting shifts (pitch factor < 1) read stale"), ("mid:keys+scan", "int dst0[16], dst1[16];"),
       ("mid:sources+stale", "sources into registers: own run"), ("mid:zero", "(PVB_RING_EXACT: while contracting every bin"),
       ("mid:scatter", "first sub-step: plain stores"), ("mid:rmw", "second sub-step: left halves add on top")]
for name, mark in sub:
    for i, l in enumerate(lines):
        if mark in l:
            bounds.append((i + 1, name))
            break
for name, mark in [("g:split+mags+stale", "---- A: real split in registers -> X planes"), ("g:mags+masks", "---- B: squared magnitudes of the run"),
                   ("g:descriptors", "---- C: descriptors"), ("g:gather+unsplit", "---- D: gather + Hermitian C2R pre-pass in registers")]:
    for i, l in enumerate(lines):
        if mark in l:
            bounds.append((i + 1, name))
            break
bounds.sort()
BODY = (next(i + 1 for i, l in enumerate(lines) if "__device__ __forceinline__ bool ring_one_call" in l),
        next(i + 1 for i, l in enumerate(lines) if "#undef PVB_NOT_DUMP" in l))

def phase_of(line):
    name = "prologue"
    for b, n in bounds:
        if line >= b:
            name = n
    return name

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(td, cub)], capture_output=True, text=True).stdout
    if not sass:
        sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cub)], capture_output=True, text=True).stdout

cur_fn, cur_line, in_fn = None, 0, False
block, block_open = [], False
cnt = collections.defaultdict(collections.Counter)
seq = []                       # (phase, opcode) of every instruction of the function, in SASS order
pending = None
for l in sass.split("\n"):
    m = re.match(r"\.text\.(\S+):", l)
    if m:
        in_fn = want in m.group(1) and "pv_process_ring_kernel" in m.group(1)
        cur_line = 0
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if not block_open:
            block, block_open = [], True
        block.append((m.group(1), int(m.group(2))))
        block += [(ff, int(nn)) for ff, nn in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        continue
    if block_open:
        # a block of location lines (innermost first) precedes the instructions it covers
        block_open = False
        cur_line = 0
        for ff, nn in block:
            if ff.endswith("pv_kernel_ring.cuh") and BODY[0] <= nn <= BODY[1]:
                cur_line = nn
                break
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        op = m.group(2)
        base = op.split(".")[0]
        if base in ("LDS", "STS", "LDG", "STG", "LDGSTS", "ATOMS", "ATOMG", "RED", "LDSM"):
            w = re.search(r"\.(64|128)", op)
            base = base + (w.group(1) if w else "32")
        cnt[phase_of(cur_line)][base] += 1
        seq.append((phase_of(cur_line), op))

if os.environ.get("SASS_PHASES_SEQ"):
    with open(os.environ["SASS_PHASES_SEQ"], "w") as f:
        for ph, op in seq:
            f.write(f"{ph}\t{op}\n")
CLASSES = [("mem", lambda o: o[:3] in ("LDS", "STS", "LDG", "STG", "ATO", "RED")), ("fp", lambda o: o[0] == "F" or o in ("HADD2", "HFMA2", "MUFU", "DADD", "DMUL", "DFMA")),
           ("shfl/vote", lambda o: o in ("SHFL", "VOTE", "VOTEU", "REDUX", "MATCH", "BAR", "WARPSYNC", "NANOSLEEP", "BSYNC", "BSSY")),]
tot = collections.Counter()
print(f"{'phase':20s} {'inst':>6s} {'fp':>6s} {'int/other':>9s} {'LDS+STS':>8s}  smem instructions by width / others")
order = ["prologue"] + [n for _, n in bounds]
for ph in order:
    c = cnt.get(ph)
    if not c:
        continue
    n = sum(c.values())
    fp = sum(v for o, v in c.items() if CLASSES[1][1](o))
    mem = sum(v for o, v in c.items() if CLASSES[0][1](o))
    sm = {o: v for o, v in c.items() if o[:3] in ("LDS", "STS", "LDG", "STG")}
    top = ", ".join(f"{o} {v}" for o, v in sorted(sm.items()))
    oth = ", ".join(f"{o} {v}" for o, v in c.most_common(9) if o not in sm)
    print(f"{ph:20s} {n:6d} {fp:6d} {n - fp - mem:9d} {mem:8d}  {top} | {oth}")
    tot.update(c)
print("total", sum(tot.values()))
