"""Pins the CPU oracle (oracle/phaze_oracle.c) against golden vectors produced by RUNNING
THE REFERENCE'S OWN JAVASCRIPT (tests/golden/generate_golden.py, via oracle/jsmini.py).
The oracle does the same float64 / float32 arithmetic in the same order, so the match is
required to be bit-exact for the float32 outputs."""
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "scenario" not in p)


def test_fixtures_present():
    assert len(CASES) >= 10


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_matches_reference_output(oracle, path):
    g = np.load(path)
    N, hop, pf = int(g["frame"]), int(g["hop"]), np.float32(g["pitch_factor"])
    x, want = g["input"], g["output"]
    got = oracle.OracleProcessor(N, hop, x.shape[0]).run(x, pf)
    assert np.array_equal(got, want), f"max abs diff {np.abs(got - want).max():.3e}"


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_matches_reference_intermediates(oracle, path):
    """last call, last channel: the raw realTransform buffer (INCLUDING the stale bins above
    N/2 that shiftPeaks reads for pitchFactor < 1), |X|^2 in float32, and the peak list."""
    g = np.load(path)
    N, hop, pf = int(g["frame"]), int(g["hop"]), np.float32(g["pitch_factor"])
    x = g["input"][-1]
    calls = x.size // hop
    hist = np.concatenate([np.zeros(N, np.float32), x])[-N:] if x.size < N else x[-N:]
    r = oracle.frame(hist, pf, float((calls - 1) * hop))
    want_spec = g["last_spectrum"][0::2] + 1j * g["last_spectrum"][1::2]
    assert np.array_equal(r["spectrum"], want_spec)
    assert np.array_equal(r["magnitudes"], g["last_magnitudes"])
    assert np.array_equal(r["peaks"], g["last_peaks"])
    # the stale region is not the Hermitian mirror: this is what SURVEY.md F4 is about
    mirror = np.conj(want_spec[1:N // 2][::-1])
    assert not np.allclose(want_spec[N // 2 + 1:], mirror)


def test_oracle_matches_reference_pause_and_channel_change(oracle):
    g = np.load(os.path.join(GOLDEN, "scenario_pause_and_channel_change.npz"))
    N, hop, pf = int(g["frame"]), int(g["hop"]), np.float32(g["pitch_factor"])
    x, want, layout = g["input"], g["output"], g["layout"]
    p = oracle.OracleProcessor(N, hop, 1)
    for t, nch in enumerate(layout):
        sl = slice(t * hop, (t + 1) * hop)
        if nch == 0:                                   # paused: zero-length blocks
            out = p.process_packed(None, pf)
            assert np.array_equal(out[0], want[t, 0])
            continue
        if nch != p.num_channels:                      # reallocateChannelsIfNeeded: state -> 0
            cursor = p.time_cursor
            p.resize(int(nch))
            assert p.time_cursor == cursor
        out = p.process_packed(x[:nch, sl], pf)
        assert np.array_equal(out, want[t, :nch]), f"call {t}"
    assert p.time_cursor == len(layout) * hop
