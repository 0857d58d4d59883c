"""GPU tests of the drop-in surface: golden vectors through the CUDA path, the reference-style
process(inputs, outputs, parameters) mirror, pause / channel change, checkpointing, batching,
sharding and full-size (BASELINE.json) properties.  Everything goes through the C ABI."""
import glob
import os

import numpy as np
import pytest

from phaze_b200 import signals

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "scenario" not in p)
RMS_BAR = 1e-4            # north_star tolerance (float32 RMS)
RMS_EXPECTED = 2e-6


def _rms(a):
    return float(np.sqrt(np.mean(np.square(np.asarray(a, np.float64)))))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_cuda_matches_reference_golden(path):
    """CUDA output vs the output of the reference's own JavaScript (committed fixtures)."""
    from phaze_b200 import BatchedPhaseVocoder
    g = np.load(path)
    N, hop, pf = int(g["frame"]), int(g["hop"]), np.float32(g["pitch_factor"])
    x, want = g["input"], g["output"]
    with BatchedPhaseVocoder(x.shape[0], N, hop) as pv:
        got = pv.run(x, pf)
    err = _rms(got - want)
    print(f"{os.path.basename(path)}: rms err {err:.3e} (out rms {_rms(want):.3e})")
    assert err <= RMS_BAR and err <= RMS_EXPECTED


def test_mirror_process_pause_and_channel_change():
    """PhaseVocoderProcessor.process(inputs, outputs, parameters) against the reference run with
    paused (zero-length) blocks and a channel-count change (ola-processor.js:38-52,93-100)."""
    from phaze_b200 import PhaseVocoderProcessor
    g = np.load(os.path.join(GOLDEN, "scenario_pause_and_channel_change.npz"))
    hop, pf = int(g["hop"]), np.float32(g["pitch_factor"])
    x, want, layout = g["input"], g["output"], g["layout"]
    proc = PhaseVocoderProcessor({"numberOfInputs": 1, "numberOfOutputs": 1})
    assert (proc.blockSize, proc.hopSize, proc.nbOverlaps) == (2048, 128, 16)
    errs = []
    for t, nch in enumerate(layout):
        sl = slice(t * hop, (t + 1) * hop)
        if nch == 0:
            ins, n_out = [[np.zeros(0, np.float32)]], 1
        else:
            ins, n_out = [[x[c, sl].copy() for c in range(nch)]], int(nch)
        outs = [[np.full(hop, np.nan, np.float32) for _ in range(n_out)]]
        keep = [a.copy() for a in ins[0]]
        assert proc.process(ins, outs, {"pitchFactor": np.array([pf], np.float32)}) is True
        for a, b in zip(ins[0], keep):
            assert np.array_equal(a, b)                     # inputs are not modified
        for c in range(n_out):
            errs.append(outs[0][c] - want[t, c])
    assert proc.timeCursor == len(layout) * hop
    assert _rms(np.concatenate(errs)) <= RMS_EXPECTED
    proc.close()


def test_pitch_factor_takes_last_element():
    from phaze_b200 import PhaseVocoderProcessor
    x = signals.channels(0, 1, 8 * 128)
    outs = []
    for arr in (np.array([1.3], np.float32), np.r_[np.full(127, 0.5), 1.3].astype(np.float32)):
        proc = PhaseVocoderProcessor({"numberOfInputs": 1, "numberOfOutputs": 1})
        o = []
        for t in range(8):
            out = [[np.zeros(128, np.float32)]]
            proc.process([[x[0, t * 128:(t + 1) * 128]]], out, {"pitchFactor": arr})
            o.append(out[0][0])
        outs.append(np.concatenate(o))
        proc.close()
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("N,hop,pf", [(1024, 256, 0.8), (2048, 512, 1.5), (2048, 128, 0.8), (256, 64, 1.2),
                                      (512, 128, 0.9), (4096, 1024, 1.25), (1024, 64, 0.8), (1024, 256, 0.5), (1024, 256, 0.4)])
@pytest.mark.parametrize("K", [4, 16, 64, 70])
def test_process_many_is_bit_identical_to_single_calls(N, hop, pf, K):
    """K consecutive calls in one submission: with PVB_OPT_MANY_MODE = 1 they share
    kernel launches where the ring-order kernel applies (each pair loops over the hops, state stays in
    L1 / L2); elsewhere
    (hop 64 at frame 1024, pitch factor 0.5: the ring-order kernel's DEEP instances) it is K launches.  Both must equal K single calls bit for
    bit, through the host entry point (copies pipelined in groups) and the device entry point."""
    import torch
    from phaze_b200 import BatchedPhaseVocoder
    C = 7
    calls = K + 3
    x = signals.channels(0, C, calls * hop)
    blocks = np.ascontiguousarray(x.reshape(C, calls, hop).transpose(1, 0, 2))
    with BatchedPhaseVocoder(C, N, hop) as a, BatchedPhaseVocoder(C, N, hop, many_mode=1) as b, \
            BatchedPhaseVocoder(C, N, hop, many_mode=1) as c:
        one = np.stack([a.process(blocks[t], pf) for t in range(calls)])
        many = np.concatenate([b.process_many(blocks[:3], pf), b.process_many(blocks[3:], pf)])
        assert a.time_cursor == b.time_cursor == calls * hop
        kernel = b.kernel_name(np.float32(pf))
        ring = "ring" in kernel and "(deep" not in kernel
        assert b.kernel_launches < calls if ring else b.kernel_launches == calls
        din = torch.from_numpy(blocks).cuda()
        dout = torch.empty_like(din)
        torch.cuda.synchronize()
        c.process_device(din.data_ptr(), dout.data_ptr(), pf, None, num_calls=calls)
        c.sync()
        dev = dout.cpu().numpy()
    if "pv_process_kernel" in kernel or "atomics" in kernel:
        # the generic kernel (and the ring-order kernel below pitch factor 0.5) adds colliding regions with
        # shared-memory atomics: the order, and so the last bit, is not reproducible from run to run
        assert np.abs(one - many).max() <= 1e-6 and np.abs(one - dev).max() <= 1e-6
    else:
        assert np.array_equal(one, many)
        assert np.array_equal(one, dev)


@pytest.mark.parametrize("N,hop,pf", [(1024, 256, 0.8), (2048, 128, 1.2)])
def test_checkpoint_resume(N, hop, pf):
    from phaze_b200 import BatchedPhaseVocoder
    C, calls, cut = 5, 3 * (N // hop), N // hop + 3
    x = signals.channels(7, C, calls * hop)
    with BatchedPhaseVocoder(C, N, hop) as a:
        full = a.run(x, pf)
    with BatchedPhaseVocoder(C, N, hop) as a:
        first = a.run(x[:, :cut * hop], pf)
        state = a.get_state()
    assert state["input_history"].shape == (C, N) and state["time_cursor"] == cut * hop
    # the history is the newest N input samples in time order (inputBuffers[..][0..N) of the JS)
    assert np.array_equal(state["input_history"], x[:, cut * hop - N:cut * hop])
    assert not state["output_accumulator"][:, N - hop:].any()        # ola-processor.js:134
    with BatchedPhaseVocoder(C, N, hop) as b:
        b.run(x[:, :3 * hop], pf)                    # put the rings at a different phase first
        b.set_state(state)
        rest = b.run(x[:, cut * hop:], pf)
    assert np.array_equal(np.concatenate([first, rest], axis=1), full)


def test_resize_resets_state_and_keeps_cursor(oracle):
    from phaze_b200 import BatchedPhaseVocoder
    N, hop, pf = 1024, 256, np.float32(1.2)
    x = signals.channels(0, 4, 12 * hop)
    with BatchedPhaseVocoder(2, N, hop) as pv:
        pv.run(x[:2, :6 * hop], pf)
        pv.resize(4)
        assert pv.num_channels == 4 and pv.time_cursor == 6 * hop
        got = pv.run(x[:, 6 * hop:], pf)
    ref = oracle.OracleProcessor(N, hop, 4)
    ref.time_cursor = 6 * hop
    want = ref.run(x[:, 6 * hop:], pf)
    assert _rms(got - want) <= RMS_EXPECTED


def test_paused_input_is_a_block_of_zeros(oracle):
    from phaze_b200 import BatchedPhaseVocoder
    N, hop, pf = 1024, 256, np.float32(0.8)
    x = signals.channels(0, 3, 10 * hop)
    ref = oracle.OracleProcessor(N, hop, 3)
    with BatchedPhaseVocoder(3, N, hop) as pv:
        for t in range(10):
            blk = None if t in (4, 5) else x[:, t * hop:(t + 1) * hop]
            assert _rms(pv.process(blk, pf) - ref.process_packed(blk, pf)) <= RMS_EXPECTED


def test_silence_gives_exact_zeros_and_channels_are_independent():
    from phaze_b200 import BatchedPhaseVocoder
    N, hop, pf = 1024, 256, np.float32(0.8)
    x = signals.channels(0, 6, 12 * hop)
    x[2:4] = 0.0                                       # one silent channel PAIR
    with BatchedPhaseVocoder(6, N, hop) as pv:
        y = pv.run(x, pf)
    assert not y[2:4].any()
    with BatchedPhaseVocoder(2, N, hop) as pv:         # channels 4,5 alone: same bits
        y45 = pv.run(x[4:6], pf)
    assert np.array_equal(y45, y[4:6])


def test_logical_sharding_is_bit_identical():
    """G logical shards on one GPU == the unsharded handle, bit for bit."""
    from phaze_b200 import BatchedPhaseVocoder
    from phaze_b200.sharded import shard_bounds
    N, hop, pf, C, calls = 1024, 256, np.float32(1.25), 22, 9
    x = signals.channels(100, C, calls * hop)
    with BatchedPhaseVocoder(C, N, hop) as pv:
        whole = pv.run(x, pf)
    for G in (2, 4, 8):
        parts = []
        for lo, hi in shard_bounds(C, G):
            with BatchedPhaseVocoder(hi - lo, N, hop) as pv:
                parts.append(pv.run(x[lo:hi], pf))
        assert np.array_equal(np.concatenate(parts), whole), f"G={G}"


def test_full_size_config2_properties(oracle):
    """BASELINE config 2 at full size: 4096 channels, 1024 / 256.  Size-independent checks:
    unity pitch is the 0.375-scaled delay; at pitchFactor 0.8 a spread of channels (both ends of
    the range and of every 8-way shard) matches the oracle."""
    from phaze_b200 import BatchedPhaseVocoder
    from phaze_b200.sharded import shard_bounds
    N, hop, C, calls = 1024, 256, 4096, 12
    rng = np.random.default_rng(0)
    x = (0.5 * rng.standard_normal((C, calls * hop))).astype(np.float32).clip(-1, 1)
    with BatchedPhaseVocoder(C, N, hop) as pv:
        y1 = pv.run(x, 1.0)
    d = N - hop
    assert _rms(y1[:, d:] - 0.375 * x[:, :-d]) <= 2e-7
    with BatchedPhaseVocoder(C, N, hop) as pv:
        y = pv.run(x, np.float32(0.8))
        assert 0 < pv.kernel_launches <= calls
    picks = sorted({c for lo, hi in shard_bounds(C, 8) for c in (lo, lo + 1, hi - 2, hi - 1)} | {777, 2049})
    want = oracle.OracleProcessor(N, hop, len(picks)).run(x[picks], np.float32(0.8))
    err = _rms(y[picks] - want)
    print(f"config 2 full size: {len(picks)} channels vs oracle rms err {err:.3e}")
    assert err <= RMS_EXPECTED


def test_full_size_config3_properties(oracle):
    """BASELINE config 3: 1024 stereo streams (2048 channels), 2048 / 512, pitchFactor 1.5"""
    from phaze_b200 import BatchedPhaseVocoder
    N, hop, C, calls = 2048, 512, 2048, 9
    x = signals.uniform_noise(0, 8, calls * hop, amp=0.5)
    big = np.tile(x, (C // 8, 1))
    with BatchedPhaseVocoder(C, N, hop) as pv:
        y = pv.run(big, np.float32(1.5))
    want = oracle.OracleProcessor(N, hop, 8).run(x, np.float32(1.5))
    assert _rms(y[:8] - want) <= RMS_EXPECTED
    assert np.array_equal(y[:8], y[C - 8:])            # identical channels -> identical bits everywhere


@pytest.mark.parametrize("N,hop,C,pf", [
    (1024, 256, 32768, 1.25),      # BASELINE config 4: all 32768 channels (4681 CTAs of 7 pairs at frame 1024)
    (256, 64, 8192, 1.2),          # BASELINE config 5: 8192 channels at every frame size of the sweep
    (512, 128, 8192, 0.8),
    (1024, 256, 8192, 1.2),
    (2048, 512, 8192, 0.8),
    (4096, 1024, 8192, 1.2),       # 2048 CTAs of 2 pairs
    (1024, 256, 4096, 0.6),        # the headline channel count at the low end of the demo's pitch range: DEEP instances
    (2048, 128, 2048, 0.5),        # the reference's own frame / hop at the bottom of that range
])
def test_full_size_configs_4_and_5(oracle, N, hop, C, pf):
    """BASELINE configs 4 and 5 at their full channel counts (the grid-size edges the small parity cases do
    not reach).  (a) A spread of channels -- both ends of the range, both sides of every 8-way shard
    boundary, every 1021st channel -- matches the CPU oracle; (b) size-independent: the same channels
    processed alone in a small handle give the same bits (channels are independent, pv:49-53), and a
    silent channel pair in the middle of the batch stays exactly zero."""
    from phaze_b200 import BatchedPhaseVocoder
    from phaze_b200.sharded import shard_bounds
    calls = N // hop + 6
    rng = np.random.default_rng(N + C)
    base = signals.channels(200, 128, calls * hop)
    gain = rng.uniform(0.5, 1.0, C).astype(np.float32)
    x = np.ascontiguousarray(base[np.arange(C) % 128] * gain[:, None])       # every channel its own roundings
    quiet = 2 * (C // 4)
    x[quiet:quiet + 2] = 0.0
    picks = sorted({c for lo, hi in shard_bounds(C, 8) for c in (lo, lo + 1, hi - 2, hi - 1)} | set(range(5, C, 1021)))
    with BatchedPhaseVocoder(C, N, hop) as pv:
        y = pv.run(x, np.float32(pf))
        assert pv.kernel_launches == calls and pv.ring_stuck_count == 0
    assert not y[quiet:quiet + 2].any()
    want = oracle.OracleProcessor(N, hop, len(picks)).run(x[picks], np.float32(pf))
    err = _rms(y[picks] - want)
    # The default peak-guard policy leaves a single natural near-tie of a broadband frame to the float32
    # decision (pv_kernel_ring.cuh, "Peak guard"); with 2049 bins per frame one of these few hundred frames
    # can contain a tie that float32 decides the other way (4e-6 RMS over the sample, 25x under the bar).
    # The strict policy re-decides those frames and must be at the float32 noise floor.
    with BatchedPhaseVocoder(len(picks), N, hop, peak_guard=3) as pv:
        strict = pv.run(x[picks], np.float32(pf))
    err_strict = _rms(strict - want)
    print(f"N={N} C={C} pf={pf}: {len(picks)} channels vs oracle rms err {err:.3e} (default policy), "
          f"{err_strict:.3e} (strict)")
    assert err <= 1e-5
    assert err_strict <= RMS_EXPECTED
    pairs = sorted({c & ~1 for c in picks})[:24]
    idx = [c for p in pairs for c in (p, p + 1)]
    with BatchedPhaseVocoder(len(idx), N, hop) as pv:
        small = pv.run(x[idx], np.float32(pf))
    assert np.array_equal(small, y[idx])
