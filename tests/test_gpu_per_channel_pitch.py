"""Per-channel pitch factors (pvb_process_pf): the reference takes one scalar per
processor and call (phase-vocoder.js:47); here every channel of a handle has its own.  Channel c must
get exactly what a reference processor with the scalar pitch_factors[c] computes."""
import numpy as np
import pytest

from phaze_b200 import signals

pytestmark = pytest.mark.gpu

RMS_EXPECTED = 2e-6


def _rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64)))))


def _oracle_per_channel(oracle, N, hop, x, pf):
    """C independent reference processors, one scalar each; pf: [C] or [T][C]"""
    C, total = x.shape
    T = total // hop
    ref = np.empty_like(x)
    for c in range(C):
        p = oracle.OracleProcessor(N, hop, 1)
        for k in range(T):
            f = pf[k, c] if pf.ndim == 2 else pf[c]
            ref[c:c + 1, k * hop:(k + 1) * hop] = p.process_packed(x[c:c + 1, k * hop:(k + 1) * hop], np.float32(f))
    return ref


@pytest.mark.parametrize("N,hop,C", [(1024, 256, 9), (2048, 128, 6), (2048, 512, 7), (512, 128, 37), (256, 64, 70),
                                     (4096, 1024, 5)])
def test_per_channel_pitch_in_ring_range(oracle, N, hop, C):
    """every channel its own factor inside the ring-order kernel's range (both sides of 1 in one pair)"""
    from phaze_b200 import BatchedPhaseVocoder
    calls = 2 * (N // hop) + 5
    rng = np.random.default_rng(N + hop)
    pf = rng.uniform(0.75, 2.0, C).astype(np.float32)
    pf[0], pf[1] = np.float32(0.8), np.float32(1.25)
    if C > 4:
        pf[2], pf[3], pf[4] = np.float32(1.0), np.float32(0.75), np.float32(3.0)
    x = signals.channels(100, C, calls * hop)
    ref = _oracle_per_channel(oracle, N, hop, x, pf)
    with BatchedPhaseVocoder(C, N, hop) as pv:
        got = pv.run_pf(x, pf)
        assert pv.kernel_launches == calls
    per_channel = np.sqrt(np.mean(np.square((got - ref).astype(np.float64)), axis=1))
    print(f"N={N} hop={hop} C={C}: worst channel rms err {per_channel.max():.3e}")
    assert per_channel.max() <= RMS_EXPECTED


@pytest.mark.parametrize("N,hop,C", [(1024, 256, 9), (2048, 128, 6), (2048, 512, 7), (512, 128, 37), (256, 64, 70), (4096, 1024, 5)])
def test_per_channel_pitch_down_to_one_half(oracle, N, hop, C):
    """factors in [0.5, 0.75) for some channels, above for others (both kinds in one pair): the ring-order
    kernel's DEEP instances with per-pair key tables, one launch per call, no state re-layout"""
    from phaze_b200 import BatchedPhaseVocoder
    calls = 2 * (N // hop) + 5
    rng = np.random.default_rng(N + 3 * hop)
    pf = rng.uniform(0.34, 1.5, C).astype(np.float32)
    pf[0], pf[1], pf[2], pf[3], pf[4] = (np.float32(v) for v in (0.5, 1.25, 0.62, 0.36, 0.9))
    x = signals.channels(160, C, calls * hop)
    ref = _oracle_per_channel(oracle, N, hop, x, pf)
    with BatchedPhaseVocoder(C, N, hop) as pv:
        got = pv.run_pf(x, pf)
        assert pv.kernel_launches == calls
        st = pv.get_state()
    per_channel = np.sqrt(np.mean(np.square((got - ref).astype(np.float64)), axis=1))
    print(f"N={N} hop={hop} C={C}: worst channel rms err {per_channel.max():.3e}")
    assert per_channel.max() <= RMS_EXPECTED


def test_uniform_array_in_deep_range_equals_scalar_call_bitwise():
    """pitch_factors[c] == f in [0.5, 0.75) for every c must give the bits of pvb_process(..., f)"""
    from phaze_b200 import BatchedPhaseVocoder
    N, hop, C, calls = 1024, 256, 11, 9
    x = signals.channels(141, C, calls * hop)
    for f in (0.5, 0.66):
        with BatchedPhaseVocoder(C, N, hop) as a, BatchedPhaseVocoder(C, N, hop) as b:
            want = a.run(x, np.float32(f))
            got = b.run_pf(x, np.full(C, f, np.float32))
        assert np.array_equal(got, want)


def test_per_channel_pitch_changes_every_call_and_leaves_the_range(oracle):
    """factors that change from call to call (k-rate automation) and wander outside [0.75, 64] for some
    channels and calls: those calls run on the generic kernel, state is re-laid in between"""
    from phaze_b200 import BatchedPhaseVocoder
    N, hop, C, calls = 1024, 256, 5, 18
    rng = np.random.default_rng(7)
    pf = rng.uniform(0.8, 1.6, (calls, C)).astype(np.float32)
    pf[6:9, 1] = np.float32(0.5)
    pf[12, 3] = np.float32(0.3)
    x = signals.channels(120, C, calls * hop)
    ref = _oracle_per_channel(oracle, N, hop, x, pf)
    with BatchedPhaseVocoder(C, N, hop) as pv:
        got = pv.run_pf(x, pf)
    per_channel = np.sqrt(np.mean(np.square((got - ref).astype(np.float64)), axis=1))
    print(f"worst channel rms err {per_channel.max():.3e}")
    assert per_channel.max() <= RMS_EXPECTED


def test_uniform_array_equals_scalar_call_bitwise():
    """pitch_factors[c] == f for every c must give the bits of pvb_process(..., f)"""
    from phaze_b200 import BatchedPhaseVocoder
    N, hop, C, calls = 1024, 256, 11, 9
    x = signals.channels(140, C, calls * hop)
    for f in (0.8, 1.25):
        with BatchedPhaseVocoder(C, N, hop) as a, BatchedPhaseVocoder(C, N, hop) as b:
            want = a.run(x, np.float32(f))
            got = b.run_pf(x, np.full(C, f, np.float32))
        assert np.array_equal(got, want)


def test_per_channel_pitch_clean_tones(oracle):
    """the peak guard works the same in the per-channel kernel"""
    from phaze_b200 import BatchedPhaseVocoder
    N, hop, C, calls = 1024, 256, 4, 14
    x = np.stack([signals.channel(30 + c, calls * hop, noise=0.0) for c in range(C)])
    pf = np.array([0.8, 1.2, 0.9, 0.75], np.float32)
    ref = _oracle_per_channel(oracle, N, hop, x, pf)
    with BatchedPhaseVocoder(C, N, hop) as pv:
        got = pv.run_pf(x, pf)
        assert pv.peak_guard_count > 0
    assert _rms(got - ref) <= RMS_EXPECTED
