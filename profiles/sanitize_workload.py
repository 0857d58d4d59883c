#!/usr/bin/env python
"""Small workload for compute-sanitizer (profiles/sanitize.sh): every frame size of the ring-order
kernel, flag-mode chains on a stream, two handles chained through a buffer, per-channel pitch factors,
several calls per launch, the peak guard's exact path, paused input, odd channel counts.  Checks every
result against the CPU oracle so that a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_lib                       # noqa: E402
from phaze_b200 import BatchedPhaseVocoder, signals   # noqa: E402


def rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64)))))


def deep(worst):
    """pitch factors in [0.5, 0.75): the DEEP instances (three coloured sub-steps of the shift, sub-transform sums
    from the per-pair scratch), scalar and per channel"""
    for N, hop in [(256, 64), (512, 128), (1024, 256), (2048, 128), (2048, 512), (4096, 1024)]:
        C, calls = 7, 2 * (N // hop) + 3
        x = signals.channels(50, C, calls * hop)
        for pf in (0.36, 0.5, 0.62, 0.74):
            ref = oracle_lib.OracleProcessor(N, hop, C).run(x, np.float32(pf))
            with BatchedPhaseVocoder(C, N, hop) as pv:
                assert "(deep" in pv.kernel_name(np.float32(pf))
                worst = max(worst, rms(pv.run(x, np.float32(pf)) - ref))
        for lo in (0.5, 0.34):
            pfs = np.linspace(lo, 1.3, C).astype(np.float32)
            ref = np.concatenate([oracle_lib.OracleProcessor(N, hop, 1).run(x[c:c + 1], pfs[c]) for c in range(C)])
            with BatchedPhaseVocoder(C, N, hop) as pv:
                worst = max(worst, rms(pv.run_pf(x, pfs) - ref))
        print(f"deep: frame {N} hop {hop}: worst rms error so far {worst:.3e}", flush=True)
    return worst


def main():
    worst = 0.0
    if "--deep-only" in sys.argv:
        worst = deep(worst)
        print(f"worst rms error vs oracle: {worst:.3e}")
        assert worst <= 2e-6
        print("WORKLOAD OK")
        return
    worst = deep(worst)
    for N, hop in [(256, 64), (512, 128), (1024, 256), (2048, 128), (2048, 512), (4096, 1024)]:
        C, calls = 9, 2 * (N // hop) + 3
        for pf in (0.8, 1.25):
            x = signals.channels(0, C, calls * hop)
            ref = oracle_lib.OracleProcessor(N, hop, C).run(x, np.float32(pf))
            # host entry point, one launch per call / calls sharing launches / exact peak decisions everywhere
            for opts in ({}, {"many_mode": 1}, {"peak_guard": 2}):
                with BatchedPhaseVocoder(C, N, hop, **opts) as pv:
                    worst = max(worst, rms(pv.run(x, np.float32(pf)) - ref))
            # device entry point: flag-mode chain on a stream, a paused call in the middle
            xin = torch.from_numpy(np.ascontiguousarray(x.reshape(C, calls, hop).transpose(1, 0, 2))).cuda()
            out = torch.empty_like(xin)
            st = torch.cuda.Stream()
            with BatchedPhaseVocoder(C, N, hop, inputs_ready=1) as pv:
                torch.cuda.synchronize()
                for k in range(calls):
                    pv.process_device(xin[k].data_ptr(), out[k].data_ptr(), np.float32(pf), st.cuda_stream)
                st.synchronize()
                assert pv.ring_stuck_count == 0
            worst = max(worst, rms(out.cpu().numpy().transpose(1, 0, 2).reshape(C, -1) - ref))
        # per-channel pitch factors
        pfs = np.linspace(0.76, 1.9, C).astype(np.float32)
        x = signals.channels(20, C, calls * hop)
        ref = np.concatenate([oracle_lib.OracleProcessor(N, hop, 1).run(x[c:c + 1], pfs[c]) for c in range(C)])
        with BatchedPhaseVocoder(C, N, hop) as pv:
            worst = max(worst, rms(pv.run_pf(x, pfs) - ref))
        # clean tones: the guard's exact path decides
        x = np.stack([signals.channel(30 + c, calls * hop, noise=0.0) for c in range(4)])
        ref = oracle_lib.OracleProcessor(N, hop, 4).run(x, np.float32(0.8))
        with BatchedPhaseVocoder(4, N, hop) as pv:
            worst = max(worst, rms(pv.run(x, np.float32(0.8)) - ref))
        print(f"frame {N} hop {hop}: worst rms error so far {worst:.3e}", flush=True)
    # two handles chained through one buffer (the library must order the aliasing launches)
    N, hop, C, calls = 1024, 256, 12, 10
    x = signals.channels(7, C, calls * hop)
    ra = oracle_lib.OracleProcessor(N, hop, C).run(x, np.float32(0.8))
    rb = oracle_lib.OracleProcessor(N, hop, C).run(ra, np.float32(1.25))
    xin = torch.from_numpy(np.ascontiguousarray(x.reshape(C, calls, hop).transpose(1, 0, 2))).cuda()
    mid = torch.empty((C, hop), dtype=torch.float32, device="cuda")
    out = torch.empty_like(xin)
    st = torch.cuda.Stream()
    with BatchedPhaseVocoder(C, N, hop, inputs_ready=1) as a, BatchedPhaseVocoder(C, N, hop, inputs_ready=1) as b:
        torch.cuda.synchronize()
        for k in range(calls):
            a.process_device(xin[k].data_ptr(), mid.data_ptr(), np.float32(0.8), st.cuda_stream)
            b.process_device(mid.data_ptr(), out[k].data_ptr(), np.float32(1.25), st.cuda_stream)
        st.synchronize()
    worst = max(worst, rms(out.cpu().numpy().transpose(1, 0, 2).reshape(C, -1) - rb))
    print(f"worst rms error vs oracle: {worst:.3e}")
    assert worst <= 2e-6
    print("WORKLOAD OK")


if __name__ == "__main__":
    main()
