"""Channel sharding of one logical processor across the GPUs of a node (one process per GPU).

Channels are fully independent on this path (src/phase-vocoder.js:49-53 loops over them
serially; the only shared datum is timeCursor, phase-vocoder.js:71, which every shard
advances identically), so the path shards with NO data-path collective: rank r owns a
contiguous block of channels, its history / overlap-add rings live in its own HBM, and
`process_local` touches nothing but local memory.

`process_from_root` adds the exchange step the north-star names: rank 0 holds the full
[C][hop] input block, scatters each rank's slab (grouped point-to-point sends ==
ncclSend/ncclRecv over NVLink under the "nccl" backend), every rank processes its slab,
and the output slabs are gathered back on rank 0.  `process_stream_from_root` is the same
exchange for a stream of audio buffers ([C][K*hop], K process() calls per message): slabs are
sent without a staging copy, and the scatter of buffer i+1 and the gather of buffer i-1 run
under the kernels of buffer i.

The per-shard engine is injected (`processor_factory`) so the partition / exchange logic is
testable on CPU with the gloo backend; the default factory is the CUDA engine.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_channels: int, world_size: int) -> List[Tuple[int, int]]:
    """[(first, last+1)] per rank: contiguous blocks, boundaries on even channels (the CUDA
    kernels process channels in pairs; pairing must not depend on the number of shards)."""
    pairs = (num_channels + 1) // 2
    out = []
    for r in range(world_size):
        lo = min(2 * ((pairs * r) // world_size), num_channels)
        hi = min(2 * ((pairs * (r + 1)) // world_size), num_channels)
        out.append((lo, hi))
    return out


class _CudaShard:
    """Default per-shard engine: the C-ABI handle on this rank's GPU, device tensors in/out."""

    def __init__(self, num_channels: int, frame_size: int, hop_size: int, device: torch.device):
        from .processor import BatchedPhaseVocoder
        self.pv = BatchedPhaseVocoder(num_channels, frame_size, hop_size, device=device.index or 0)
        self.device = device
        self.hop = hop_size
        self.num_channels = num_channels
        self.side = torch.cuda.Stream(device)      # the kernels run here, ordered after the caller's stream

    def process(self, block: Optional[torch.Tensor], pitch_factor: float) -> torch.Tensor:
        out = torch.empty((self.num_channels, self.hop), dtype=torch.float32, device=self.device)
        if self.num_channels == 0:
            self.pv.process_device(None, 0, pitch_factor)      # keeps timeCursor in step
            return out
        cur = torch.cuda.current_stream(self.device)
        self.side.wait_stream(cur)                 # inputs (e.g. a finished NCCL recv) are ready
        in_ptr = None
        if block is not None:
            assert block.is_cuda and block.dtype == torch.float32 and block.is_contiguous()
            in_ptr = block.data_ptr()
            block.record_stream(self.side)
        out.record_stream(self.side)
        self.pv.process_device(in_ptr, out.data_ptr(), pitch_factor, self.side.cuda_stream)
        cur.wait_stream(self.side)                 # whoever consumes `out` on the caller's stream waits
        return out


    def process_many(self, blocks: torch.Tensor, pitch_factor: float) -> torch.Tensor:
        """blocks [K][num_channels][hop] on the device: K consecutive process() calls in one
        submission (pvb_process_many_device); returns [K][num_channels][hop]."""
        K = blocks.shape[0]
        out = torch.empty((K, self.num_channels, self.hop), dtype=torch.float32, device=self.device)
        if self.num_channels == 0:
            self.pv.process_device(None, 0, pitch_factor, num_calls=K)
            return out
        assert blocks.is_cuda and blocks.dtype == torch.float32 and blocks.is_contiguous()
        cur = torch.cuda.current_stream(self.device)
        self.side.wait_stream(cur)
        blocks.record_stream(self.side)
        out.record_stream(self.side)
        self.pv.process_device(blocks.data_ptr(), out.data_ptr(), pitch_factor, self.side.cuda_stream,
                               num_calls=K)
        cur.wait_stream(self.side)
        return out


class ShardedPhaseVocoder:
    def __init__(self, num_channels: int, frame_size: int = 2048, hop_size: int = 128,
                 group: Optional[dist.ProcessGroup] = None,
                 processor_factory: Optional[Callable[[int], object]] = None,
                 device: Optional[torch.device] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.num_channels, self.frame_size, self.hop_size = num_channels, frame_size, hop_size
        self.bounds = shard_bounds(num_channels, self.world)
        self.first, self.last = self.bounds[self.rank]
        self.local_channels = self.last - self.first
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device("cpu")
        self.device = device
        if processor_factory is None:
            processor_factory = lambda n: _CudaShard(n, frame_size, hop_size, device)   # noqa: E731
        self.engine = processor_factory(self.local_channels)
        self._gather_group = None       # second communicator: results travel while the next slabs arrive

    # -- shard-resident steady state: no collective --------------------------------------------
    def process_local(self, local_block: Optional[torch.Tensor], pitch_factor: float) -> torch.Tensor:
        """local_block: this rank's [local_channels][hop] slab (None == paused)."""
        return self.engine.process(local_block, pitch_factor)

    # -- single-root mode: scatter, process, gather ----------------------------------------------
    def process_from_root(self, full_block: Optional[torch.Tensor], pitch_factor: float,
                          root: int = 0) -> Optional[torch.Tensor]:
        """full_block [C][hop] on `root` (ignored elsewhere); None on root == paused input
        (ola-processor.js:93-100): the root then scatters blocks of zeros, which is exactly what the
        reference feeds its frames, so no extra message is needed to tell the other ranks.
        Returns the [C][hop] output on root."""
        hop = self.hop_size
        if full_block is None and self.rank == root and self.world > 1:
            full_block = torch.zeros((self.num_channels, hop), dtype=torch.float32, device=self.device)
        local = torch.empty((self.local_channels, hop), dtype=torch.float32, device=self.device)
        if self.world == 1:
            local = full_block
        else:
            ops = []
            if self.rank == root:
                for r, (lo, hi) in enumerate(self.bounds):
                    if r == root:
                        local = full_block[lo:hi].contiguous()
                    elif hi > lo:
                        ops.append(dist.P2POp(dist.isend, full_block[lo:hi].contiguous(), r, self.group))
            elif self.local_channels > 0:
                ops.append(dist.P2POp(dist.irecv, local, root, self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        out_local = self.engine.process(local, pitch_factor)
        if self.world == 1:
            return out_local
        result = None
        ops = []
        if self.rank == root:
            result = torch.empty((self.num_channels, hop), dtype=torch.float32, device=self.device)
            for r, (lo, hi) in enumerate(self.bounds):
                if r == root:
                    result[lo:hi] = out_local
                elif hi > lo:
                    ops.append(dist.P2POp(dist.irecv, result[lo:hi], r, self.group))
        elif self.local_channels > 0:
            ops.append(dist.P2POp(dist.isend, out_local.contiguous(), root, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return result

    # -- single-root mode for a stream of audio buffers: K calls per message, pipelined -------------
    def _engine_many(self, blocks: torch.Tensor, pitch_factor: float) -> torch.Tensor:
        if hasattr(self.engine, "process_many"):
            return self.engine.process_many(blocks, pitch_factor)
        return torch.stack([self.engine.process(blocks[k], pitch_factor) for k in range(blocks.shape[0])])

    def process_stream_from_root(self, buffers, pitch_factor: float, calls_per_buffer: int,
                                 num_buffers: int, root: int = 0) -> list:
        """`buffers`: on `root`, a sequence of `num_buffers` tensors [C][K*hop] (channel-major audio,
        K = calls_per_buffer consecutive process() calls per channel); ignored elsewhere.  Returns
        the [C][K*hop] outputs on root (an empty list elsewhere).

        A rank's slab of a buffer is contiguous, so it is sent and received in place (one message
        of K hops per peer and direction); the [channel][call] <-> [call][channel] re-ordering the
        kernels need happens on the owning rank.  Messages are posted one buffer ahead: the scatter
        of buffer i+1 and the gather of buffer i-1 overlap the kernels of buffer i; the gathers use
        a communicator of their own (their own NCCL stream), so both NVLink directions are busy at
        once.  The result is bit-identical to `num_buffers * K` process_from_root calls."""
        hop, K = self.hop_size, calls_per_buffer
        Cl, lo, hi = self.local_channels, self.first, self.last
        is_root = self.rank == root
        results: list = []
        if self.world == 1:
            for i in range(num_buffers):
                blocks = buffers[i].view(Cl, K, hop).transpose(0, 1).contiguous()
                out = self._engine_many(blocks, pitch_factor)
                results.append(out.transpose(0, 1).contiguous().view(Cl, K * hop))
            return results

        if self._gather_group is None:
            self._gather_group = dist.new_group(ranks=list(range(self.world))) if self.group is None \
                else dist.new_group(ranks=dist.get_process_group_ranks(self.group))
        ggroup = self._gather_group

        def post_scatter(i):
            """-> (works, this rank's [Cl][K*hop] slab of buffer i)"""
            ops = []
            if is_root:
                buf = buffers[i]
                assert buf.shape == (self.num_channels, K * hop) and buf.is_contiguous()
                for r, (a, b) in enumerate(self.bounds):
                    if r != root and b > a:
                        ops.append(dist.P2POp(dist.isend, buf[a:b], r, self.group))
                slab = buf[lo:hi]
            else:
                slab = torch.empty((Cl, K * hop), dtype=torch.float32, device=self.device)
                if Cl > 0:
                    ops.append(dist.P2POp(dist.irecv, slab, root, self.group))
            return (dist.batch_isend_irecv(ops) if ops else []), slab

        def post_gather(out_slab):
            """-> (works, the [C][K*hop] result on root / None)"""
            ops, res = [], None
            if is_root:
                res = torch.empty((self.num_channels, K * hop), dtype=torch.float32, device=self.device)
                res[lo:hi] = out_slab
                for r, (a, b) in enumerate(self.bounds):
                    if r != root and b > a:
                        ops.append(dist.P2POp(dist.irecv, res[a:b], r, ggroup))
            elif Cl > 0:
                ops.append(dist.P2POp(dist.isend, out_slab, root, ggroup))
            return (dist.batch_isend_irecv(ops) if ops else []), res

        pending_in = post_scatter(0) if num_buffers > 0 else None
        pending_out = []                                   # [(works, result, keep-alive)]
        for i in range(num_buffers):
            works, slab = pending_in
            for w in works:
                w.wait()                                   # buffer i has arrived
            if i + 1 < num_buffers:
                pending_in = post_scatter(i + 1)           # travels under the kernels of buffer i
            blocks = slab.view(Cl, K, hop).transpose(0, 1).contiguous()
            out = self._engine_many(blocks, pitch_factor)
            out_slab = out.transpose(0, 1).contiguous().view(Cl, K * hop)
            gw, res = post_gather(out_slab)                # travels under the kernels of buffer i+1
            pending_out.append((gw, res, out_slab))
            while len(pending_out) > 2:                    # bound the buffers in flight
                gw0, res0, _ = pending_out.pop(0)
                for w in gw0:
                    w.wait()
                if is_root:
                    results.append(res0)
        for gw0, res0, _ in pending_out:
            for w in gw0:
                w.wait()
            if is_root:
                results.append(res0)
        return results
