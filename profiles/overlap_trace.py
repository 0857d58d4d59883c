#!/usr/bin/env python
"""Direct evidence of how consecutive launches of the headline kernel overlap.

Needs the experiment build (make -C phaze_b200/csrc experiments), which stamps every ring-order launch
with %globaltimer at its earliest CTA start and latest CTA end:

    PVB_LIBRARY=phaze_b200/libphaze_b200_exp.so PVB_TRACE=400 python profiles/overlap_trace.py [mode]

mode 0 (default): programmatic dependent launch + per-pair completion flags; 1: whole-grid wait; 2: plain
launches.  Same workload as bench.py (4096 channels, 1024 / 256, pitch factor 0.8, 7 handles rotated).
Prints, per launch, start and end relative to the first start, its duration, the gap between consecutive
starts (== the per-step time of the bench) and how long it ran concurrently with its predecessor."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import phaze_b200                                   # noqa: E402
from phaze_b200 import signals                      # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
N, hop, Cn, rotate, launches = 1024, 256, 4096, 7, 60
lib = phaze_b200.load_library()
lib.pvb_trace_dump.restype = C.c_int32
lib.pvb_trace_dump.argtypes = [C.c_void_p, C.c_int32]
x = signals.channels(0, Cn, 8 * hop)
blocks = torch.from_numpy(np.ascontiguousarray(x.reshape(Cn, 8, hop).transpose(1, 0, 2))).cuda()
outs = [torch.empty((Cn, hop), dtype=torch.float32, device="cuda") for _ in range(rotate)]
procs = [phaze_b200.BatchedPhaseVocoder(Cn, N, hop, inputs_ready=1, launch_mode=mode) for _ in range(rotate)]
st = torch.cuda.Stream()
warm = rotate * (N // hop) + 40
torch.cuda.synchronize()
torch.cuda._sleep(int(3e6))
for i in range(warm + launches):
    procs[i % rotate].process_device(blocks[i % 8].data_ptr(), outs[i % rotate].data_ptr(), np.float32(0.8), st.cuda_stream)
torch.cuda.synchronize()
buf = (C.c_ulonglong * (2 * (warm + launches)))()
n = lib.pvb_trace_dump(buf, warm + launches)
t = np.array(buf[:2 * n], dtype=np.int64).reshape(n, 2)[warm:]
t0 = t[0, 0]
print(f"launch mode {mode}: {len(t)} consecutive launches of pv_process_ring_kernel<1024, 2> (4096 channels each)")
print("launch  start_us    end_us  duration_us  start_to_start_us  overlap_with_previous_us")
for i, (a, b) in enumerate(t):
    gap = (a - t[i - 1, 0]) * 1e-3 if i else float("nan")
    ov = max(0.0, (t[i - 1, 1] - a) * 1e-3) if i else float("nan")
    print(f"{i:6d} {(a - t0) * 1e-3:9.2f} {(b - t0) * 1e-3:9.2f} {(b - a) * 1e-3:12.2f} {gap:18.2f} {ov:25.2f}")
d = (t[:, 1] - t[:, 0]) * 1e-3
g = np.diff(t[:, 0]) * 1e-3
print(f"mean duration of one launch {d.mean():.2f} us; mean start-to-start {g.mean():.2f} us "
      f"(= time per step); mean overlap with the previous launch {np.maximum(0, (t[:-1, 1] - t[1:, 0]) * 1e-3).mean():.2f} us")
