"""Multi-process (gloo, world_size 2, CPU) test of the channel-sharding host logic.

The per-shard engine is the CPU oracle here (tests may use it); what is under test is
phaze_b200.sharded: the partition, the scatter / gather exchange and the claim that
sharding does not change a single bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phaze_b200.sharded import ShardedPhaseVocoder, shard_bounds


def test_shard_bounds_cover_and_even():
    for C in (1, 2, 7, 64, 4095, 4096, 32768):
        for G in (1, 2, 3, 4, 8):
            b = shard_bounds(C, G)
            assert b[0][0] == 0 and b[-1][1] == C
            for (lo, hi), (lo2, _) in zip(b, b[1:]):
                assert hi == lo2 and lo <= hi
                assert hi % 2 == 0 or hi == C           # pair boundaries
    assert shard_bounds(32768, 8)[3] == (12288, 16384)


class _OracleShard:
    def __init__(self, n, frame, hop):
        from oracle import oracle_lib
        self.p = oracle_lib.OracleProcessor(frame, hop, n) if n > 0 else None
        self.n, self.hop = n, hop

    def process(self, block, pf):
        if self.n == 0:
            return torch.empty((0, self.hop), dtype=torch.float32)
        out = self.p.process_packed(None if block is None else block.numpy(), pf)
        return torch.from_numpy(out)


def _worker(rank, world, port, C, frame, hop, calls, pf, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from phaze_b200 import signals
        sh = ShardedPhaseVocoder(C, frame, hop, processor_factory=lambda n: _OracleShard(n, frame, hop),
                                 device=torch.device("cpu"))
        x = signals.channels(0, C, calls * hop)
        outs_root, outs_local = [], []
        for t in range(calls):
            blk = torch.from_numpy(np.ascontiguousarray(x[:, t * hop:(t + 1) * hop]))
            if t == 3:
                blk = None                        # paused on the root (ola:93-100): every rank must follow
            res = sh.process_from_root(blk if rank == 0 else None, pf)
            if rank == 0:
                outs_root.append(res.numpy().copy())
        # second pass: shard-resident mode on fresh engines must give the same slab
        sh2 = ShardedPhaseVocoder(C, frame, hop, processor_factory=lambda n: _OracleShard(n, frame, hop),
                                  device=torch.device("cpu"))
        for t in range(calls):
            blk = torch.from_numpy(np.ascontiguousarray(x[sh2.first:sh2.last, t * hop:(t + 1) * hop]))
            outs_local.append(sh2.process_local(None if t == 3 else blk, pf).numpy().copy())
        # third pass: the streamed root mode (K calls per message, channel-major audio buffers)
        sh3 = ShardedPhaseVocoder(C, frame, hop, processor_factory=lambda n: _OracleShard(n, frame, hop),
                                  device=torch.device("cpu"))
        K = 2
        bufs = [torch.from_numpy(np.ascontiguousarray(x[:, i * K * hop:(i + 1) * K * hop]))
                for i in range(calls // K)] if rank == 0 else None
        res = sh3.process_stream_from_root(bufs, pf, K, calls // K)
        streamed = np.concatenate([r.numpy() for r in res], axis=1) if rank == 0 else None
        assert rank == 0 or res == []
        q.put((rank, (sh.first, sh.last), np.stack(outs_root) if rank == 0 else None, np.stack(outs_local),
               streamed))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_sharded_equals_unsharded_bit_for_bit():
    from oracle import oracle_lib
    from phaze_b200 import signals
    oracle_lib.build()
    C, frame, hop, calls, pf = 7, 256, 64, 10, np.float32(0.8)     # odd C: uneven shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, C, frame, hop, calls, pf, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    x = signals.channels(0, C, calls * hop)
    ref = oracle_lib.OracleProcessor(frame, hop, C).run(x, pf)               # unsharded
    xp = x.copy()
    xp[:, 3 * hop:4 * hop] = 0                                               # call 3 is paused in passes 1 and 2
    ref_p = oracle_lib.OracleProcessor(frame, hop, C).run(xp, pf)
    ref_calls = ref_p.reshape(C, calls, hop).transpose(1, 0, 2)
    by_rank = {r[0]: r for r in results}
    assert np.array_equal(by_rank[0][2], ref_calls), "root-gathered output differs from unsharded"
    assert np.array_equal(by_rank[0][4], ref), "streamed root mode differs from unsharded"
    for rank, (lo, hi), _, local, _ in results:
        assert np.array_equal(local, ref_calls[:, lo:hi]), f"rank {rank} shard-resident output differs"
