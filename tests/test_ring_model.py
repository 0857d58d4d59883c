"""CPU check of the ring-order kernel's design (tests/ring_kernel_model.py restates
phaze_b200/csrc/pv_kernel_ring.cuh lane by lane in numpy): with both rings aligned to the
time cursor the per-region rotation of shiftPeaks disappears, and the run-based region scan,
the two ordered shift sub-steps and the rebuilt stale bins reproduce the oracle."""
import numpy as np
import pytest

from phaze_b200 import signals

import ring_kernel_model as model


def _rms(a):
    return float(np.sqrt(np.mean(np.square(a.astype(np.float64)))))


@pytest.mark.parametrize("frame,hop,pf,calls,start", [
    (1024, 256, 0.8, 10, 0), (1024, 256, 1.2, 9, 0), (1024, 256, 1.25, 9, 3), (1024, 128, 0.75, 20, 5),
    (1024, 512, 1.5, 6, 1), (1024, 256, 1.0, 9, 0), (1024, 256, 3.0, 8, 2),
    (2048, 512, 1.5, 7, 0), (2048, 512, 0.8, 7, 2), (2048, 128, 0.8, 20, 3), (2048, 128, 1.2, 18, 0),
    (2048, 1024, 2.0, 4, 1),
    (512, 128, 0.8, 12, 0), (512, 128, 1.2, 12, 3), (512, 256, 1.5, 6, 1), (512, 128, 0.75, 12, 2),
    (4096, 1024, 1.2, 6, 0), (4096, 1024, 0.8, 6, 1), (4096, 512, 0.8, 10, 3), (4096, 256, 1.3, 18, 5),
    (4096, 2048, 1.5, 3, 1),
    (256, 64, 1.2, 14, 0), (256, 64, 0.8, 14, 1), (256, 64, 0.75, 14, 2), (256, 128, 1.3, 8, 1), (512, 64, 0.8, 20, 3),
])
def test_model_matches_oracle(oracle, frame, hop, pf, calls, start):
    x = signals.channels(11, 2, calls * hop)
    ref_p = oracle.OracleProcessor(frame, hop, 2)
    ref_p.time_cursor = start * hop
    ref = ref_p.run(x, np.float32(pf))
    got = model.run(x, pf, hop, start_calls=start, frame=frame)
    assert _rms(ref) > 1e-2
    assert _rms(got - ref) <= 2e-8


@pytest.mark.parametrize("frame,hop,pf,calls,start", [
    (1024, 256, 0.6, 9, 0), (1024, 256, 0.5, 9, 3), (1024, 128, 0.7, 18, 1), (1024, 256, 0.74, 8, 0),
    (1024, 256, 0.4, 9, 2), (1024, 256, 0.34, 9, 0), (1024, 512, 0.45, 6, 1),
    (2048, 128, 0.55, 20, 3), (2048, 512, 0.36, 7, 0), (512, 128, 0.62, 12, 1), (512, 128, 0.4, 12, 0),
    (4096, 1024, 0.5, 5, 1), (4096, 1024, 0.38, 5, 0), (256, 64, 0.6, 14, 1), (256, 64, 0.35, 14, 2),
])
def test_model_deep_instances_match_oracle(oracle, frame, hop, pf, calls, start):
    """pitch factors in [0.33, 0.75): the stale slots the last region reaches (first level and the quarter 3N/4 + o
    from four spectrum terms, second level from sixteen, deeper levels from N/64 or fewer windowed frame samples)
    and the scatter in three coloured sub-steps (the model asserts that no sub-step has two writers for a bin)"""
    x = signals.channels(21, 2, calls * hop)
    ref_p = oracle.OracleProcessor(frame, hop, 2)
    ref_p.time_cursor = start * hop
    ref = ref_p.run(x, np.float32(pf))
    got = model.run(x, pf, hop, start_calls=start, frame=frame)
    assert _rms(ref) > 1e-3
    assert _rms(got - ref) <= 5e-8


def test_model_shared_memory_patterns_are_conflict_free(oracle):
    """every 128-bit exchange pattern of the three radix-8 passes and the run reads take the
    minimum number of wavefronts; the split stores at most twice that (lane 0 is special)"""
    cf = model.Conflicts()
    x = signals.channels(0, 2, 6 * 256)
    model.run(x, 0.8, 256, cf)
    model.run(signals.channels(0, 2, 5 * 512), 0.8, 512, cf, frame=2048)
    for name in ("p1_st", "p2_ld", "p3_ldA", "p3_ldB", "run_ld", "stale_ld"):
        assert cf.worst[name] == 1.0, (name, cf.worst)
    assert max(cf.worst.values()) <= 2.0, cf.worst
    # frame 512: two pairs per warp, PAIR_BYTES apart; padded exchange rows (stride 74, groups of 9)
    cf = model.Conflicts()
    cf.pair_bytes = model.Geo(512).PAIR_BYTES
    assert cf.pair_bytes == 4800
    model.run(signals.channels(0, 2, 6 * 128), 0.8, 128, cf, frame=512)
    for name in ("p1_st", "p2_ld", "p3_ldA", "run_ld", "stale_ld"):
        assert cf.worst[name] == 1.0, (name, cf.worst)
    assert max(cf.worst.values()) <= 2.0, cf.worst


@pytest.mark.parametrize("frame,hop,pf,calls", [
    (1024, 256, 0.8, 9), (1024, 256, 0.75, 9), (1024, 256, 1.25, 8), (1024, 256, 3.0, 7), (256, 64, 0.8, 12),
    (2048, 512, 1.5, 6), (4096, 1024, 0.85, 4),
])
@pytest.mark.parametrize("variant", ["reference", "runs", "rows"])
def test_gather_middle_prototype_matches_oracle(oracle, monkeypatch, frame, hop, pf, calls, variant):
    """the destination-order gather planned as the next middle: one descriptor per
    peak, every destination bin reads its one or two sources; same output as the scatter in two
    ordered sub-steps.  "reference": whole-array restatement; "runs" / "rows": thread by thread with
    32-bit descriptors, destinations owned as runs of 16 or in row order (lane l: bins TP r + l)"""
    monkeypatch.setattr(model, "MIDDLE", "gather" if variant == "reference" else "gather_lanes")
    monkeypatch.setattr(model, "ROWS", variant == "rows")
    x = signals.channels(5, 2, calls * hop)
    ref = oracle.OracleProcessor(frame, hop, 2).run(x, np.float32(pf))
    got = model.run(x, pf, hop, frame=frame)
    assert _rms(ref) > 1e-2
    assert _rms(got - ref) <= 2e-8


@pytest.mark.parametrize("frame,hop,pf,calls", [
    (1024, 256, 0.8, 9), (1024, 256, 0.75, 9), (1024, 256, 1.25, 8), (1024, 256, 3.0, 7), (1024, 256, 1.0, 6),
    (1024, 128, 17.0, 10), (1024, 512, 2.0, 5), (256, 64, 0.8, 12), (512, 128, 1.3, 8), (2048, 512, 1.5, 6),
    (2048, 512, 0.77, 6),
])
def test_gather_rows_blueprint_matches_oracle(oracle, monkeypatch, frame, hop, pf, calls):
    """the gather middle of pv_kernel_ring.cuh (PVB_RING_GATHER) restated thread by thread: 32-bit region
    descriptors in destination space (T' | overlap | -delta), copies at the first two bins of every aligned
    group of lanes, destinations taken in the order of the Hermitian pre-pass (model.gather_rows_kernel)"""
    monkeypatch.setattr(model, "MIDDLE", "gather_rows2")
    x = signals.channels(5, 2, calls * hop)
    ref = oracle.OracleProcessor(frame, hop, 2).run(x, np.float32(pf))
    got = model.run(x, pf, hop, frame=frame)
    assert _rms(ref) > 1e-2
    assert _rms(got - ref) <= 2e-8


@pytest.mark.parametrize("pf", [0.75, 0.8, 0.93, 1.0, 1.07, 1.5, 2.0, 5.0, 40.0])
@pytest.mark.parametrize("kind", ["tones", "silence_then_tone", "sine", "silence", "impulse"])
def test_gather_rows_blueprint_equals_scatter_model(monkeypatch, pf, kind):
    """few peaks, long regions, regions that start below bin 0 or end beyond nb, no peaks at all: the gather
    blueprint and the scatter model (same float32 spectrum) give bit-identical output"""
    hop, calls = 256, 7
    n = calls * hop
    if kind == "tones":
        x = signals.channels(3, 2, n, noise=0.0)
    elif kind == "silence_then_tone":
        x = np.stack([signals.silence_then_tone(1, n, 700), signals.silence_then_tone(2, n, 1300)])
    elif kind == "sine":
        x = np.stack([signals.bin_centred_sine(n, 1024, 37), signals.bin_centred_sine(n, 1024, 300, amp=0.2)])
    elif kind == "silence":
        x = np.zeros((2, n), np.float32)
        x[1] = signals.channel(9, n)
    else:
        x = np.zeros((2, n), np.float32)
        x[0, 900] = 1.0
        x[1, 100::333] = 0.5
    monkeypatch.setattr(model, "MIDDLE", "scatter")
    want = model.run(x, pf, hop)
    monkeypatch.setattr(model, "MIDDLE", "gather_rows2")
    got = model.run(x, pf, hop)
    assert np.array_equal(want, got)
