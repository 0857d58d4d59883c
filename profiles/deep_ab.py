"""Pitch factors in [0.5, 0.75): the ring-order kernel's DEEP instances against the generic kernel
(PVB_OPT_KERNEL = generic), device-resident, measured like bench.py's other_configs.
    python profiles/deep_ab.py [channels [frame]]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
peak, _ = bench.measured_peak()
ONLY = int(sys.argv[2]) if len(sys.argv) > 2 else 0        # optional: one frame size
for frame, hop in ((1024, 256), (2048, 512), (2048, 128), (512, 128), (4096, 1024)):
    if ONLY and frame != ONLY:
        continue
    ch = C if frame <= 1024 else C * 1024 // frame
    for pf in (0.34, 0.4, 0.45, 0.5, 0.6, 0.7, 0.74, 0.8):
        row = {}
        for name, opts in (("ring", {}), ("generic", {"kernel": "generic"})):
            r = bench.quick_config(0, frame, hop, ch, pf, peak, steps=100, warm=20, **opts)
            row[name] = r
        print(f"frame {frame} hop {hop} ch {ch} pf {pf}: {row['ring']['kernel']} {row['ring']['ms_per_step'] * 1e3:.1f} us "
              f"(frac {row['ring']['roofline_frac']:.3f}) | generic {row['generic']['ms_per_step'] * 1e3:.1f} us "
              f"(frac {row['generic']['roofline_frac']:.3f}) | speed-up {row['generic']['ms_per_step'] / row['ring']['ms_per_step']:.2f}x", flush=True)
