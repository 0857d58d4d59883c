"""torchrun helper (NCCL, one process per GPU): single-root scatter -> process -> gather through
phaze_b200.sharded must equal the unsharded single-GPU run bit for bit.  Launched by
tests/test_gpu_multi.py when the box has >= 2 GPUs:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_root_scatter.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from phaze_b200 import BatchedPhaseVocoder, signals          # noqa: E402
from phaze_b200.sharded import ShardedPhaseVocoder            # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    C, N, hop, calls, pf = 70, 1024, 256, 10, np.float32(1.25)      # config 4 arithmetic, uneven shards
    sh = ShardedPhaseVocoder(C, N, hop)
    x = signals.channels(0, C, calls * hop)
    outs = []
    for t in range(calls):
        blk = torch.from_numpy(np.ascontiguousarray(x[:, t * hop:(t + 1) * hop])).cuda() if rank == 0 else None
        res = sh.process_from_root(blk, pf)
        if rank == 0:
            outs.append(res.cpu().numpy())
    # streamed root mode: K calls per message, scatter / gather overlapped with the kernels
    K = 2
    sh2 = ShardedPhaseVocoder(C, N, hop)
    bufs = [torch.from_numpy(np.ascontiguousarray(x[:, i * K * hop:(i + 1) * K * hop])).cuda()
            for i in range(calls // K)] if rank == 0 else None
    res = sh2.process_stream_from_root(bufs, pf, K, calls // K)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        with BatchedPhaseVocoder(C, N, hop, device=local) as pv:
            want = pv.run(x, pf)
        got = np.stack(outs).transpose(1, 0, 2).reshape(C, calls * hop)
        got_stream = np.concatenate([r.cpu().numpy() for r in res], axis=1)
        ok_stream = bool(np.array_equal(got_stream, want))
        ok = bool(np.array_equal(got, want)) and ok_stream
        print(f"ROOT_SCATTER world={world} bit_identical={ok} streamed_bit_identical={ok_stream}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
