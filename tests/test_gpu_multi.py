"""Multi-GPU (NCCL) test of the root scatter / gather mode; skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_root_scatter_gather_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_root_scatter.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit_identical=True" in r.stdout


@pytest.mark.parametrize("N,hop,C,pf", [(1024, 256, 70, 1.25), (2048, 128, 9, 0.8), (256, 64, 33, 1.2)])
def test_multi_device_handle_is_bit_identical(N, hop, C, pf):
    """pvb_multi_* (one process, several devices, SURVEY 8(b) device_ids[]): host entry points and the
    single-root entry point (slabs scattered / gathered with peer copies) against one handle with all the
    channels.  On a one-GPU box the shards are logical (the same device listed several times); with more
    GPUs they sit on different devices."""
    import numpy as np
    import torch
    from phaze_b200 import BatchedPhaseVocoder, MultiDevicePhaseVocoder, signals
    ngpu = torch.cuda.device_count()
    calls = 2 * (N // hop) + 3
    x = signals.channels(300, C, calls * hop)
    blocks = np.ascontiguousarray(x.reshape(C, calls, hop).transpose(1, 0, 2))
    with BatchedPhaseVocoder(C, N, hop, device=0) as pv:
        want = pv.process_many(blocks, np.float32(pf))
    for shards in (2, 3, 8):
        devices = [i % ngpu for i in range(shards)]
        with MultiDevicePhaseVocoder(C, N, hop, devices=devices) as mv:
            bounds = mv.shards
            assert bounds[0][0] == 0 and sum(n for _, n in bounds) == C
            assert all(lo % 2 == 0 for lo, _ in bounds)
            got = np.concatenate([mv.process_many(blocks[:4], np.float32(pf)),
                                  np.stack([mv.process(blocks[k], np.float32(pf)) for k in range(4, calls)])])
        assert np.array_equal(got, want), f"{shards} shards, host entry points"
        with MultiDevicePhaseVocoder(C, N, hop, devices=devices) as mv:
            torch.cuda.set_device(devices[0])
            din = torch.from_numpy(blocks).cuda()
            dout = torch.zeros_like(din)
            torch.cuda.synchronize()
            mv.process_root(din.data_ptr(), dout.data_ptr(), np.float32(pf), num_calls=5)
            for k in range(5, calls):
                mv.process_root(din[k].data_ptr(), dout[k].data_ptr(), np.float32(pf))
            got = dout.cpu().numpy()
        assert np.array_equal(got, want), f"{shards} shards, root entry point"


def test_multi_device_handle_paused_and_errors():
    import numpy as np
    from phaze_b200 import BatchedPhaseVocoder, MultiDevicePhaseVocoder, PhazeError, signals
    N, hop, C = 1024, 256, 6
    x = signals.channels(310, C, 8 * hop)
    blocks = np.ascontiguousarray(x.reshape(C, 8, hop).transpose(1, 0, 2))
    with BatchedPhaseVocoder(C, N, hop) as pv, MultiDevicePhaseVocoder(C, N, hop, devices=[0, 0, 0, 0]) as mv:
        for k in range(8):
            blk = None if k in (3, 4) else blocks[k]
            assert np.array_equal(mv.process(blk, np.float32(0.8)), pv.process(blk, np.float32(0.8)))
    with pytest.raises(PhazeError):
        MultiDevicePhaseVocoder(4, 1000, 250, devices=[0])          # not a power of two
    with pytest.raises(PhazeError):
        MultiDevicePhaseVocoder(4, 1024, 256, devices=[99])         # no such device
