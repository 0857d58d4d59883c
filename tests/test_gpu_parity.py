"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): RMS error <= 1e-4 in float32."""
import numpy as np
import pytest

from phaze_b200 import signals

pytestmark = pytest.mark.gpu

RMS_BAR = 1e-4          # north_star: "within 1e-4 RMS (float32)"
RMS_EXPECTED = 2e-6     # what a correct float32 pipeline actually achieves on these inputs


def _run_both(oracle, N, hop, C, pf, calls, first_channel=0, sig=None, **options):
    from phaze_b200 import BatchedPhaseVocoder
    x = signals.channels(first_channel, C, calls * hop) if sig is None else sig
    ref = oracle.OracleProcessor(N, hop, C).run(x, pf)
    # one launch per call (default), then consecutive calls sharing launches (PVB_OPT_MANY_MODE = 1: up to
    # 16 per launch here, where the ring-order kernel applies): both paths must give the same bits
    with BatchedPhaseVocoder(C, N, hop, **options) as pv:
        got = pv.run(x, pf)
        assert pv.kernel_launches == calls
    with BatchedPhaseVocoder(C, N, hop, many_mode=1, **options) as pv:
        got_many = pv.run(x, pf)
        assert 0 < pv.kernel_launches <= calls
        kernel = pv.kernel_name(pf)
    if "pv_process_kernel" in kernel or "atomics" in kernel:
        # the generic kernel (and the ring-order kernel below pitch factor 0.5) adds colliding regions with
        # shared-memory atomics: the last bit depends on their order
        assert np.abs(got - got_many).max() <= 1e-6
    else:
        assert np.array_equal(got, got_many), "calls sharing a launch differ from one launch per call"
    return x, ref, got


def _rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64)))))


@pytest.mark.parametrize("N,hop,C,pf", [
    (1024, 256, 1, 1.2),      # BASELINE config 1 (mono)
    (1024, 256, 8, 0.8),      # config 2 arithmetic (stale upper bins, colliding regions)
    (2048, 512, 6, 1.5),      # config 3 arithmetic
    (1024, 256, 7, 1.25),     # config 4 arithmetic, odd channel count
    (2048, 128, 4, 1.2),      # the reference's native 2048 / 128
    (2048, 128, 4, 0.8),
    (1024, 256, 4, 1.0),
])
def test_parity_configs(oracle, N, hop, C, pf):
    calls = 3 * (N // hop) + 5
    x, ref, got = _run_both(oracle, N, hop, C, np.float32(pf), calls)
    err = _rms(got - ref)
    print(f"N={N} hop={hop} C={C} pf={pf}: rms err {err:.3e}, out rms {_rms(ref):.3e}, "
          f"max abs {np.abs(got - ref).max():.3e}")
    assert _rms(ref) > 1e-2
    assert err <= RMS_BAR
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("N", [256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("pf", [1.2, 0.8])
def test_parity_frame_size_sweep(oracle, N, pf):
    hop = N // 4
    x, ref, got = _run_both(oracle, N, hop, 4, np.float32(pf), 14)
    err = _rms(got - ref)
    print(f"N={N} pf={pf}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("pf", [0.8, 0.67, 0.7, 0.75, 0.9, 1.0, 1.2, 1.5, 2.0, 3.0])
@pytest.mark.parametrize("kernel", ["ring", "warp", "cta", "generic"])
def test_parity_1024_all_kernels(oracle, pf, kernel):
    """frame 1024 has four CUDA paths: the ring-order kernel (default for pitch factors in
    [0.75, 64] and hop % 128 == 0), one warp per channel pair in frame order, the CTA kernel and the
    generic kernel (pvb_set_option PVB_OPT_KERNEL); all must match."""
    x, ref, got = _run_both(oracle, 1024, 256, 5, np.float32(pf), 17, kernel=kernel)
    err = _rms(got - ref)
    print(f"pf={pf} kernel={kernel}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("pf", [0.5, 0.6, 0.34, 0.3, 0.2])
def test_parity_deep_stale(oracle, pf):
    """pitch factors below 0.75 read past the first level of stale upper bins."""
    x, ref, got = _run_both(oracle, 1024, 256, 4, np.float32(pf), 14)
    err = _rms(got - ref)
    print(f"pf={pf}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("N,hop", [(256, 64), (256, 128), (512, 128), (1024, 256), (1024, 128), (1024, 512), (2048, 512), (2048, 128), (4096, 1024), (4096, 256)])
@pytest.mark.parametrize("pf", [0.5, 0.55, 0.62, 0.7, 0.7499])
def test_parity_ring_deep(oracle, N, hop, pf):
    """pitch factors in [0.5, 0.75) on the ring-order kernel's DEEP instances: stale slots up to N/2 + N/4 - 1
    (sub-transforms of the radix-4 recursion, bundle:394-438) rebuilt from the windowed frame, any number of
    regions per bin."""
    from phaze_b200 import BatchedPhaseVocoder
    with BatchedPhaseVocoder(3, N, hop) as pv:
        assert "(deep)" in pv.kernel_name(np.float32(pf))
    x, ref, got = _run_both(oracle, N, hop, 3, np.float32(pf), 2 * (N // hop) + 6)
    err = _rms(got - ref)
    print(f"N={N} hop={hop} pf={pf}: rms err {err:.3e}, out rms {_rms(ref):.3e}")
    assert _rms(ref) > 1e-3
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("N,hop", [(256, 64), (512, 128), (1024, 256), (1024, 128), (2048, 512), (2048, 128), (4096, 1024)])
@pytest.mark.parametrize("pf", [0.33, 0.33333334, 0.4, 0.45, 0.4999])
def test_parity_ring_deep_below_one_half(oracle, N, hop, pf):
    """the rest of the demo's reachable range (pitch slider / speed slider, main.js:82,93: down to 0.5 / 1.5): the
    last region reads into the quarter 3N/4 + o of the realTransform output (DFT_{N/4}(xw[4m+3])), and any number
    of regions can land on one bin (shared-memory atomics)"""
    from phaze_b200 import BatchedPhaseVocoder
    with BatchedPhaseVocoder(3, N, hop) as pv:
        assert "(deep, atomics)" in pv.kernel_name(np.float32(pf))
    x, ref, got = _run_both(oracle, N, hop, 3, np.float32(pf), 2 * (N // hop) + 6)
    err = _rms(got - ref)
    print(f"N={N} hop={hop} pf={pf}: rms err {err:.3e}, out rms {_rms(ref):.3e}")
    assert _rms(ref) > 1e-3
    assert err <= RMS_EXPECTED


def test_parity_hop128_warp_kernel(oracle):
    x, ref, got = _run_both(oracle, 1024, 128, 3, np.float32(0.8), 30)
    assert _rms(got - ref) <= RMS_EXPECTED


@pytest.mark.parametrize("N", [256, 512, 2048, 4096])
@pytest.mark.parametrize("pf", [0.75, 0.9, 1.0, 1.3, 2.5])
@pytest.mark.parametrize("kernel", ["auto", "cta", "generic"])
def test_parity_other_frame_sizes_both_kernels(oracle, N, pf, kernel):
    """frame sizes other than 1024: the default kernel (ring-order), the CTA kernel with the in-place
    shift (pitch factors in [0.75, 64]) and the fully generic kernel must all match the oracle."""
    hop = N // 4
    x, ref, got = _run_both(oracle, N, hop, 5, np.float32(pf), 13, kernel=kernel)
    err = _rms(got - ref)
    print(f"N={N} pf={pf} kernel={kernel}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


def test_parity_native_2048_128_r16(oracle):
    """the reference's own sizes: R = 16, rotations are not quarter turns"""
    for pf in (0.8, 1.3):
        x, ref, got = _run_both(oracle, 2048, 128, 3, np.float32(pf), 40)
        assert _rms(got - ref) <= RMS_EXPECTED


@pytest.mark.parametrize("hop", [32, 64, 128, 256, 512, 1024])
@pytest.mark.parametrize("pf", [0.8, 1.3])
def test_parity_1024_hop_sweep(oracle, hop, pf):
    """warp kernel at every hop (R = 32 ... 1): hop % 128 == 0 takes the specialisation with the
    rotation folded into the twiddles, the others the general addressing."""
    calls = 2 * (1024 // hop) + 7
    x, ref, got = _run_both(oracle, 1024, hop, 3, np.float32(pf), calls, kernel="warp")
    err = _rms(got - ref)
    print(f"hop={hop} pf={pf}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("hop", [128, 256, 512])
@pytest.mark.parametrize("pf", [0.75, 0.8, 0.97, 1.0, 1.3, 2.0, 7.5])
def test_parity_ring_kernel_hops(oracle, hop, pf):
    """the ring-order kernel (default at frame 1024) at every hop it accepts, odd channel count"""
    from phaze_b200 import BatchedPhaseVocoder
    with BatchedPhaseVocoder(3, 1024, hop) as pv:
        assert "ring" in pv.kernel_name(np.float32(pf))
    calls = 3 * (1024 // hop) + 5
    x, ref, got = _run_both(oracle, 1024, hop, 3, np.float32(pf), calls)
    err = _rms(got - ref)
    print(f"hop={hop} pf={pf}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("hop", [128, 256, 512, 1024])
@pytest.mark.parametrize("pf", [0.75, 0.8, 1.0, 1.3, 2.0])
def test_parity_ring_kernel_2048(oracle, hop, pf):
    """frame 2048 (the reference's own size): two warps per pair, radix-16 first pass"""
    from phaze_b200 import BatchedPhaseVocoder
    with BatchedPhaseVocoder(5, 2048, hop) as pv:
        assert "ring" in pv.kernel_name(np.float32(pf))
    calls = 2 * (2048 // hop) + 5
    x, ref, got = _run_both(oracle, 2048, hop, 5, np.float32(pf), calls)
    err = _rms(got - ref)
    print(f"N=2048 hop={hop} pf={pf}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("hop", [128, 256])
@pytest.mark.parametrize("pf", [0.75, 0.8, 1.0, 1.3, 2.0])
def test_parity_ring_kernel_512(oracle, hop, pf):
    """frame 512: half a warp per pair (two pairs share a warp), radix-4 first pass"""
    from phaze_b200 import BatchedPhaseVocoder
    with BatchedPhaseVocoder(5, 512, hop) as pv:
        assert "ring" in pv.kernel_name(np.float32(pf))
    calls = 3 * (512 // hop) + 5
    x, ref, got = _run_both(oracle, 512, hop, 5, np.float32(pf), calls)
    err = _rms(got - ref)
    print(f"N=512 hop={hop} pf={pf}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("C,pf", [(45, 0.8), (70, 1.25), (2, 0.9)])
def test_parity_ring_kernel_512_many_channels(oracle, C, pf):
    """several CTAs of 16 pairs, a warp whose second half-warp has no pair, an odd last channel"""
    x, ref, got = _run_both(oracle, 512, 128, C, np.float32(pf), 13)
    per_channel = np.sqrt(np.mean(np.square((got - ref).astype(np.float64)), axis=1))
    assert per_channel.max() <= RMS_EXPECTED


@pytest.mark.parametrize("hop", [64, 128])
@pytest.mark.parametrize("pf", [0.75, 0.8, 1.0, 1.3, 2.0])
@pytest.mark.parametrize("C", [5, 70])
def test_parity_ring_kernel_256(oracle, hop, pf, C):
    """frame 256: a quarter of a warp per pair (four pairs share a warp, a partly filled last warp
    and a second CTA at 70 channels); hop 64 rotates the rings by half a 128-sample block on odd calls"""
    from phaze_b200 import BatchedPhaseVocoder
    with BatchedPhaseVocoder(C, 256, hop) as pv:
        assert "ring" in pv.kernel_name(np.float32(pf))
    calls = 3 * (256 // hop) + 6
    x, ref, got = _run_both(oracle, 256, hop, C, np.float32(pf), calls)
    per_channel = np.sqrt(np.mean(np.square((got - ref).astype(np.float64)), axis=1))
    print(f"N=256 hop={hop} pf={pf} C={C}: worst channel rms err {per_channel.max():.3e}")
    assert per_channel.max() <= RMS_EXPECTED


@pytest.mark.parametrize("hop", [256, 512, 1024, 2048])
@pytest.mark.parametrize("pf", [0.75, 0.8, 1.0, 1.3, 2.0])
def test_parity_ring_kernel_4096(oracle, hop, pf):
    """frame 4096: four warps per pair, each thread holds the frame blocks of one parity (radix-16
    in registers, the radix-2 step folded into pass 2)"""
    from phaze_b200 import BatchedPhaseVocoder
    with BatchedPhaseVocoder(5, 4096, hop) as pv:
        assert "ring" in pv.kernel_name(np.float32(pf))
    calls = 2 * (4096 // hop) + 3
    x, ref, got = _run_both(oracle, 4096, hop, 5, np.float32(pf), calls)
    err = _rms(got - ref)
    print(f"N=4096 hop={hop} pf={pf}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


def test_parity_ring_kernel_4096_many_channels(oracle):
    """more pairs than one CTA holds (4 per CTA), odd last channel; hop 128 stays on the CTA kernel"""
    from phaze_b200 import BatchedPhaseVocoder
    x, ref, got = _run_both(oracle, 4096, 1024, 19, np.float32(0.8), 7)
    per_channel = np.sqrt(np.mean(np.square((got - ref).astype(np.float64)), axis=1))
    assert per_channel.max() <= RMS_EXPECTED
    with BatchedPhaseVocoder(2, 4096, 128) as pv:
        assert "cta" in pv.kernel_name(np.float32(0.8))


def test_parity_ring_kernel_2048_many_channels(oracle):
    x, ref, got = _run_both(oracle, 2048, 512, 23, np.float32(0.8), 9)
    per_channel = np.sqrt(np.mean(np.square((got - ref).astype(np.float64)), axis=1))
    assert per_channel.max() <= RMS_EXPECTED


def test_parity_ring_kernel_many_channels(oracle):
    """more pairs than one CTA holds, last pair half empty, several CTAs per SM slot"""
    x, ref, got = _run_both(oracle, 1024, 256, 45, np.float32(0.8), 11)
    assert _rms(got - ref) <= RMS_EXPECTED
    per_channel = np.sqrt(np.mean(np.square((got - ref).astype(np.float64)), axis=1))
    assert per_channel.max() <= RMS_EXPECTED


@pytest.mark.parametrize("N,hop", [(1024, 256), (256, 64), (512, 128), (2048, 512), (4096, 1024)])
def test_layout_changes_mid_stream(oracle, N, hop):
    """the pitch factor moves between the ring-order kernel (paired state, aligned to the time
    cursor) and the generic kernel (planar state): the state is re-laid on the device each time;
    a paused block, a checkpoint round trip and a time-cursor jump happen in between."""
    from phaze_b200 import BatchedPhaseVocoder
    C = 5
    plan = [(0.8, 5), (0.3, 3), (1.2, 4), (0.6, 2), (0.25, 2), (0.9, 6)]     # 0.6: the ring-order kernel's DEEP instances; 0.3, 0.25: generic
    total = sum(n for _, n in plan)
    x = signals.channels(3, C, total * hop)
    ref_p = oracle.OracleProcessor(N, hop, C)
    ref = np.empty_like(x)
    got = np.empty_like(x)
    with BatchedPhaseVocoder(C, N, hop) as pv:
        k = 0
        for pf, n in plan:
            for _ in range(n):
                s = slice(k * hop, (k + 1) * hop)
                blk = None if k == 7 else x[:, s]
                ref[:, s] = ref_p.process_packed(blk, np.float32(pf))
                got[:, s] = pv.process(blk, np.float32(pf))
                k += 1
                if k == 9:
                    st = pv.get_state()
                    pv.reset()
                    pv.set_state(st)
                if k == 12:
                    ref_p.time_cursor = 40 * hop
                    pv.time_cursor = 40 * hop
    err = _rms(got - ref)
    print(f"layout changes: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


@pytest.mark.parametrize("N,hop", [(1024, 256), (256, 64), (512, 128), (2048, 512), (4096, 1024)])
def test_flag_mode_chained_handles_and_back_to_back(oracle, N, hop):
    """Consecutive launches of the frame-1024 kernel overlap (per-pair completion flags instead of
    a grid-wide wait).  (a) one handle called back to back on a stream with device buffers: each
    pair waits for its own previous call; (b) two handles chained through ONE device buffer
    (A's output is B's input, and is overwritten by A's next call): the library must notice the
    aliasing and order those launches; (c) no flag was ever lost."""
    import torch
    from phaze_b200 import BatchedPhaseVocoder
    C, calls = 37, 24
    pfa, pfb = np.float32(0.8), np.float32(1.25)
    x = signals.channels(7, C, calls * hop)
    oa, ob = oracle.OracleProcessor(N, hop, C), oracle.OracleProcessor(N, hop, C)
    ref_a = np.empty_like(x)
    ref_b = np.empty_like(x)
    for k in range(calls):
        s = slice(k * hop, (k + 1) * hop)
        ref_a[:, s] = oa.process_packed(x[:, s], pfa)
        ref_b[:, s] = ob.process_packed(ref_a[:, s], pfb)
    stream = torch.cuda.Stream()
    xin = torch.from_numpy(np.ascontiguousarray(x.reshape(C, calls, hop).transpose(1, 0, 2))).cuda()
    mid = torch.empty((C, hop), dtype=torch.float32, device="cuda")        # shared by both handles
    out_a = torch.empty((calls, C, hop), dtype=torch.float32, device="cuda")
    out_b = torch.empty((calls, C, hop), dtype=torch.float32, device="cuda")
    # inputs_ready: xin is resident before the first submission, so launches may run in flag mode
    with BatchedPhaseVocoder(C, N, hop, inputs_ready=1) as a, BatchedPhaseVocoder(C, N, hop, inputs_ready=1) as b, \
            BatchedPhaseVocoder(C, N, hop, inputs_ready=1) as solo:
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            for k in range(calls):
                a.process_device(xin[k].data_ptr(), mid.data_ptr(), pfa, stream.cuda_stream)
                b.process_device(mid.data_ptr(), out_b[k].data_ptr(), pfb, stream.cuda_stream)
                solo.process_device(xin[k].data_ptr(), out_a[k].data_ptr(), pfa, stream.cuda_stream)
        stream.synchronize()
        assert a.ring_stuck_count == 0 and b.ring_stuck_count == 0 and solo.ring_stuck_count == 0
    got_a = out_a.cpu().numpy().transpose(1, 0, 2).reshape(C, calls * hop)
    got_b = out_b.cpu().numpy().transpose(1, 0, 2).reshape(C, calls * hop)
    assert _rms(got_a - ref_a) <= RMS_EXPECTED
    assert _rms(got_b - ref_b) <= RMS_EXPECTED


@pytest.mark.parametrize("N,hop", [(1024, 256), (256, 64), (2048, 128), (4096, 1024)])
def test_flag_mode_chain_across_kernel_instances(oracle, N, hop):
    """One handle called back to back on a stream while the pitch factor wanders over the ring-order kernel's
    three instance kinds (>= 0.75; [0.5, 0.75): ordered sub-steps; < 0.5: atomics) -- they use different CTA shapes
    (7 / 6 pairs per CTA at frame 1024), and consecutive launches overlap: each pair must still wait for ITS previous
    call's flag, whichever instance wrote it."""
    import torch
    from phaze_b200 import BatchedPhaseVocoder
    C, calls = 45, 30
    plan = [0.8, 0.8, 0.6, 0.62, 0.4, 0.8, 0.5, 0.36, 1.3, 0.74, 0.75, 0.49]
    x = signals.channels(11, C, calls * hop)
    orc = oracle.OracleProcessor(N, hop, C)
    ref = np.empty_like(x)
    for k in range(calls):
        s = slice(k * hop, (k + 1) * hop)
        ref[:, s] = orc.process_packed(x[:, s], np.float32(plan[k % len(plan)]))
    stream = torch.cuda.Stream()
    xin = torch.from_numpy(np.ascontiguousarray(x.reshape(C, calls, hop).transpose(1, 0, 2))).cuda()
    out = torch.empty((calls, C, hop), dtype=torch.float32, device="cuda")
    with BatchedPhaseVocoder(C, N, hop, inputs_ready=1) as pv:
        torch.cuda.synchronize()
        for k in range(calls):
            pv.process_device(xin[k].data_ptr(), out[k].data_ptr(), np.float32(plan[k % len(plan)]), stream.cuda_stream)
        stream.synchronize()
        assert pv.ring_stuck_count == 0 and pv.kernel_launches == calls
    got = out.cpu().numpy().transpose(1, 0, 2).reshape(C, calls * hop)
    err = _rms(got - ref)
    print(f"N={N} hop={hop}: rms err {err:.3e}")
    assert err <= RMS_EXPECTED


def test_flag_mode_chained_through_deep_buffer(oracle):
    """Two handles chained through a [K][C][hop] buffer, K = 24 calls per submission, few channels
    (tiny grids: many launches are co-resident) and a faster second handle: B's call j reads what
    A's call j wrote 24 launches earlier.  The library must order them (its alias window covers the
    real overlap depth, not just the last few launches)."""
    import torch
    from phaze_b200 import BatchedPhaseVocoder
    C, K, rounds = 6, 24, 3
    Na, Nb, hop = 2048, 256, 128
    pfa, pfb = np.float32(0.8), np.float32(1.25)
    x = signals.channels(11, C, rounds * K * hop)
    ref_a = oracle.OracleProcessor(Na, hop, C).run(x, pfa)
    ref_b = oracle.OracleProcessor(Nb, hop, C).run(ref_a, pfb)
    stream = torch.cuda.Stream()
    xin = torch.from_numpy(np.ascontiguousarray(x.reshape(C, rounds * K, hop).transpose(1, 0, 2))).cuda()
    mid = torch.empty((K, C, hop), dtype=torch.float32, device="cuda")
    out_b = torch.empty((rounds * K, C, hop), dtype=torch.float32, device="cuda")
    with BatchedPhaseVocoder(C, Na, hop, inputs_ready=1) as a, BatchedPhaseVocoder(C, Nb, hop, inputs_ready=1) as b:
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            for r in range(rounds):
                a.process_device(xin[r * K].data_ptr(), mid.data_ptr(), pfa, stream.cuda_stream, num_calls=K)
                b.process_device(mid.data_ptr(), out_b[r * K].data_ptr(), pfb, stream.cuda_stream, num_calls=K)
        stream.synchronize()
        assert a.ring_stuck_count == 0 and b.ring_stuck_count == 0
    got_b = out_b.cpu().numpy().transpose(1, 0, 2).reshape(C, rounds * K * hop)
    assert _rms(got_b - ref_b) <= RMS_EXPECTED


def test_strict_stream_order_after_foreign_producer(oracle):
    """Default (PVB_OPT_INPUTS_READY = 0): the input of a device submission may be produced by a
    kernel the library knows nothing about, enqueued on the same stream just before the call."""
    import torch
    from phaze_b200 import BatchedPhaseVocoder
    C, N, hop, calls = 64, 1024, 256, 12
    pf = np.float32(0.8)
    x = signals.channels(5, C, calls * hop)
    ref = oracle.OracleProcessor(N, hop, C).run(x, pf)
    stream = torch.cuda.Stream()
    src = torch.from_numpy(np.ascontiguousarray(x.reshape(C, calls, hop).transpose(1, 0, 2))).cuda()
    buf = torch.zeros((C, hop), dtype=torch.float32, device="cuda")
    big = torch.zeros(32 << 20, dtype=torch.float32, device="cuda")
    out = torch.empty((calls, C, hop), dtype=torch.float32, device="cuda")
    with BatchedPhaseVocoder(C, N, hop) as pv:
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            for k in range(calls):
                big.add_(1.0)                    # keeps the stream busy in front of the producer
                buf.copy_(src[k] * 2.0).mul_(0.5)   # foreign kernels write the input just before the call
                pv.process_device(buf.data_ptr(), out[k].data_ptr(), pf, stream.cuda_stream)
        stream.synchronize()
    got = out.cpu().numpy().transpose(1, 0, 2).reshape(C, calls * hop)
    assert _rms(got - ref) <= RMS_EXPECTED
