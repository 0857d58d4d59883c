"""Several calls per launch (PVB_OPT_MANY_MODE = 1) against one launch per call at SMALL channel counts, where one
call does not fill the GPU and the chain of launches is bound by the launch rate.  Device-resident, 1024/256,
pitch 0.8, K = 64 calls per submission.
    python profiles/many_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phaze_b200 import BatchedPhaseVocoder, signals  # noqa: E402

N, hop, K, pf = 1024, 256, 64, np.float32(0.8)
for C in (32, 128, 256, 512, 1024, 2048, 4096):
    x = torch.from_numpy(np.ascontiguousarray(signals.channels(0, C, K * hop).reshape(C, K, hop).transpose(1, 0, 2))).cuda()
    y = torch.empty_like(x)
    st = torch.cuda.Stream()
    row = []
    for mode in (0, 1):
        with BatchedPhaseVocoder(C, N, hop, many_mode=mode, inputs_ready=1) as pv:
            torch.cuda.synchronize()
            for _ in range(3):
                pv.process_device(x.data_ptr(), y.data_ptr(), pf, st.cuda_stream, num_calls=K)
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record(st)
            for _ in range(reps):
                pv.process_device(x.data_ptr(), y.data_ptr(), pf, st.cuda_stream, num_calls=K)
            e1.record(st)
            st.synchronize()
            row.append(e0.elapsed_time(e1) * 1e3 / (reps * K))
    print(f"channels {C:5d}: one launch per call {row[0]:6.2f} us/call | shared launches {row[1]:6.2f} us/call | ratio {row[0] / row[1]:.2f}x", flush=True)
