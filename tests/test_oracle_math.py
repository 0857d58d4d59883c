"""Known-answer properties of the oracle: independent of the golden vectors."""
import numpy as np
import pytest


@pytest.mark.parametrize("n", [16, 64, 256, 512, 1024, 2048, 4096])
def test_forward_fft_valid_half_is_rfft(oracle, n):
    x = np.random.default_rng(n).uniform(-1, 1, n).astype(np.float32)
    X = oracle.real_transform(x)
    ref = np.fft.rfft(x.astype(np.float64))
    assert np.abs(X[: n // 2 + 1] - ref).max() <= 1e-12 * n


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096])
def test_stale_upper_bins_are_sub_transforms(oracle, n):
    """bins N/2+1.. hold DFT_{N/4}(x[4m+2])[1..N/8]; 3N/4.. hold DFT_{N/4}(x[4m+3])[0..N/8];
    N/2+N/8+1.. hold DFT_{N/16}(x[16m+10])[1..N/32]  (bundle:394-438 write pattern)."""
    x = np.random.default_rng(7 * n).uniform(-1, 1, n).astype(np.float32)
    X = oracle.real_transform(x)
    xd = x.astype(np.float64)
    f = np.fft.fft
    assert np.abs(X[n // 2 + 1: n // 2 + n // 8 + 1] - f(xd[2::4])[1: n // 8 + 1]).max() < 1e-11
    assert np.abs(X[3 * n // 4: 3 * n // 4 + n // 8 + 1] - f(xd[3::4])[: n // 8 + 1]).max() < 1e-11
    a = n // 2 + n // 8 + 1
    assert np.abs(X[a: a + n // 32] - f(xd[10::16])[1: n // 32 + 1]).max() < 1e-11


def test_stale_bins_from_valid_half_identity(oracle):
    """the identity the CUDA kernels use: slot N/2+q == 1/4 W^{-2q} (X[q] - X[N/4+q] + conj X[N/2-q] - conj X[N/4-q])"""
    n = 1024
    x = np.random.default_rng(3).uniform(-1, 1, n).astype(np.float32)
    X = oracle.real_transform(x)
    q = np.arange(1, n // 8 + 1)
    w = np.exp(2j * np.pi * 2 * q / n)
    rebuilt = 0.25 * w * (X[q] - X[n // 4 + q] + np.conj(X[n // 2 - q]) - np.conj(X[n // 4 - q]))
    assert np.abs(rebuilt - X[n // 2 + q]).max() < 1e-11


def test_inverse_round_trip(oracle):
    n = 1024
    x = np.random.default_rng(5).uniform(-1, 1, n)
    spec = np.fft.fft(x)
    assert np.abs(oracle.inverse_transform(spec).real - x).max() < 1e-12


@pytest.mark.parametrize("n,hop", [(1024, 256), (2048, 128), (2048, 512), (256, 64)])
def test_unity_pitch_is_a_scaled_delay(oracle, n, hop):
    """pitchFactor 1: y[t] = (sum w^2 / R) x[t - (N - hop)] = 0.375 x[t - (N - hop)]"""
    calls = 3 * n // hop
    x = np.random.default_rng(11).uniform(-1, 1, (2, calls * hop)).astype(np.float32)
    y = oracle.OracleProcessor(n, hop, 2).run(x, 1.0)
    d = n - hop
    err = y[:, d:] - 0.375 * x[:, :-d]
    assert np.sqrt(np.mean(err ** 2)) < 5e-8


def test_silence_gives_exact_zeros(oracle):
    y = oracle.OracleProcessor(1024, 256, 2).run(np.zeros((2, 20 * 256), np.float32), 1.3)
    assert not y.any()


def test_bin_centred_sine_moves_to_scaled_bin(oracle):
    n, hop, k, pf = 1024, 256, 40, 1.5
    t = np.arange(40 * hop)
    x = (0.5 * np.sin(2 * np.pi * k * t / n))[None].astype(np.float32)
    y = oracle.OracleProcessor(n, hop, 1).run(x, pf)[0]
    seg = y[20 * hop: 20 * hop + n] * np.hanning(n + 1)[:n]
    assert np.argmax(np.abs(np.fft.rfft(seg))) == round(k * pf)


def test_highest_source_bin_read(oracle):
    """how far above N/2 shiftPeaks reads"""
    x = np.random.default_rng(2).uniform(-1, 1, (1, 24 * 256)).astype(np.float32)
    seen = {}
    for pf in (0.8, 0.75, 0.5):
        p = oracle.OracleProcessor(1024, 256, 1)
        p.run(x, pf)
        seen[pf] = p.max_source_bin
    assert 600 <= seen[0.8] <= 616 and 630 <= seen[0.75] <= 640 and 740 <= seen[0.5] <= 768


def test_bad_sizes_rejected(oracle):
    with pytest.raises(ValueError):
        oracle.OracleProcessor(1000, 250, 1)
    with pytest.raises(ValueError):
        oracle.real_transform(np.zeros(48, np.float32))


def _walk_block_tree(n, pos):
    """(L, r, s, o): slot `pos` of the realTransform output holds DFT_L(x[r m + s])[o] (bundle:394-438 writes
    only outputs 0 .. L/2 of every length-L block of the radix-4 recursion); pv_kernel_ring.cuh, DEEP instances"""
    l0 = 4 if int(np.log2(n)) % 2 == 0 else 2
    L, r, s, o = n, 1, 0, pos
    while L > l0 and o > L // 2:
        q = L // 4
        sb = o // q
        o -= sb * q
        s += r * sb
        r *= 4
        L = q
    return L, r, s, o


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
def test_deep_stale_slots_are_sums_of_frame_samples(oracle, n):
    """what the ring-order kernel's DEEP instances compute for pitch factors in [0.5, 0.75): every slot
    N/2 + q, q < N/4, is a sum of L samples x[r m + s]; beyond the first level all of them have
    n = 10 or 14 (mod 16) and sit in the kernel's scratch at 2 (n >> 4) + ((n >> 2) & 1), stride r / 8"""
    x = np.random.default_rng(5 * n).uniform(-1, 1, n).astype(np.float32)
    X = oracle.real_transform(x)
    xd = x.astype(np.float64)
    keep = np.flatnonzero((np.arange(n) % 16 == 10) | (np.arange(n) % 16 == 14))
    scratch = np.zeros(n // 8)
    scratch[2 * (keep >> 4) + ((keep >> 2) & 1)] = xd[keep]
    worst = 0.0
    for q in range(1, n // 4):
        L, r, s, o = _walk_block_tree(n, n // 2 + q)
        m = np.arange(L)
        if q <= n // 8:
            assert (L, r, s) == (n // 4, 4, 2)                      # first level: rebuilt from the valid half instead
            samples = xd[r * m + s]
        else:
            assert r >= 16 and s % 16 in (10, 14)
            samples = scratch[2 * (s >> 4) + ((s >> 2) & 1) + (r // 8) * m]
        worst = max(worst, abs(np.sum(samples * np.exp(-2j * np.pi * o * m / L)) - X[n // 2 + q]))
    assert worst < 1e-10


def test_three_colours_suffice_for_pitch_factors_from_one_half():
    """DEEP instances add colliding regions in three ordered sub-steps (ordinal of the owning peak mod 3): with
    peaks at least 3 bins apart (pv:95-116) and pitch factors >= 0.5 the images of regions i and i + 3 never
    overlap (pv:119-147), so at most three regions land on one bin"""
    rng = np.random.default_rng(11)
    n, nb = 1024, 513
    for trial in range(3000):
        peaks, p = [], 2 + int(rng.integers(0, 4))
        while p < n // 2 - 2:
            peaks.append(p)
            p += 3 + (int(rng.geometric(0.5)) - 1 if rng.random() < 0.8 else int(rng.integers(0, 40)))
        pf = np.float32(rng.choice([0.5, 0.5000001, 0.51, 0.55, 0.6, 0.66, 0.7, 0.7499, rng.uniform(0.5, 0.75)]))
        delta = [int(np.floor(p * float(pf) + 0.5)) - p for p in peaks]
        img = []
        for i, p in enumerate(peaks):
            s = 0 if i == 0 else p - (p - peaks[i - 1]) // 2
            e = n if i == len(peaks) - 1 else p + -(-(peaks[i + 1] - p) // 2)
            img.append((max(s + delta[i], 0), min(e + delta[i], nb)))
        for i in range(len(peaks) - 3):
            assert img[i + 3][0] >= img[i][1]


@pytest.mark.parametrize("n", [256, 1024, 4096])
def test_second_stale_level_from_the_spectrum(oracle, n):
    """DEEP instances: slots N/2 + q of the second level (block-tree walk stops at L = N/16, r = 16) are
    1/16 sum_u W_N^{-s (o + u L)} X[o + u L] over the Hermitian-extended spectrum: 16 terms instead of N/16 samples"""
    x = np.random.default_rng(9 * n).uniform(-1, 1, n).astype(np.float32)
    X = oracle.real_transform(x)
    m = n // 2
    full = np.concatenate([X[:m + 1], np.conj(X[1:m][::-1])])           # X[k] for k < N from the valid half
    L = n // 16
    worst, seen = 0.0, 0
    for q in range(n // 8 + 1, n // 4):
        Lw, r, s, o = _walk_block_tree(n, m + q)
        if r != 16:
            continue
        assert Lw == L and s in (10, 14) and o <= n // 32
        idx = o + L * np.arange(16)
        rebuilt = np.sum(full[idx] * np.exp(2j * np.pi * s * idx / n)) / 16
        worst = max(worst, abs(rebuilt - X[m + q]))
        seen += 1
    assert seen == 2 * (n // 32) + 1 and worst < 1e-10


@pytest.mark.parametrize("n", [256, 1024, 4096])
def test_quarter_three_slots_from_the_spectrum(oracle, n):
    """DEEP instances below pitch factor 0.5: slot 3N/4 + o, o <= N/8, holds DFT_{N/4}(x[4m + 3])[o] =
    1/4 W_N^{-3 o} (X[o] - j X[N/4 + o] - conj X[N/2 - o] + j conj X[N/4 - o])"""
    x = np.random.default_rng(13 * n).uniform(-1, 1, n).astype(np.float32)
    X = oracle.real_transform(x)
    m = n // 2
    o = np.arange(0, n // 8 + 1)
    rebuilt = 0.25 * np.exp(2j * np.pi * 3 * o / n) * (X[o] - 1j * X[n // 4 + o] - np.conj(X[m - o]) + 1j * np.conj(X[n // 4 - o]))
    assert np.abs(rebuilt - X[3 * n // 4 + o]).max() < 1e-11
    assert np.abs(X[3 * n // 4 + o] - np.fft.fft(x.astype(np.float64)[3::4])[: n // 8 + 1]).max() < 1e-10
