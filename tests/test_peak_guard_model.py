"""CPU check of the peak guard's error model (pv_kernel_ring.cuh, "Peak guard").

The lane-level numpy model of the ring-order kernel (tests/ring_kernel_model.py) gives the float32
spectrum the kernel works on; the CPU oracle gives the reference's float64 spectrum and peak list.
Checked here: (1) the float32 transform obeys |dX_k| <= a |X_k| + c ||X||_2 with the constants the
kernel uses; (2) the kernel's uncertainty test flags EVERY frame whose float32 peak set differs from
the reference's; (3) on the benchmark's broadband input it flags few frames."""
import re
import os

import numpy as np
import pytest

import ring_kernel_model as rk
from phaze_b200 import signals

HERE = os.path.dirname(os.path.abspath(__file__))


def _kernel_constants():
    src = open(os.path.join(HERE, "..", "phaze_b200", "csrc", "pv_kernel_ring.cuh")).read()
    a = float(re.search(r"#define PVB_GUARD_A ([0-9.eE+-]+)f", src).group(1))
    c = float(re.search(r"#define PVB_GUARD_C ([0-9.eE+-]+)f", src).group(1))
    return a, c


def _frames(oracle, N, hop, noise, nch, calls, pf=0.8):
    g = rk.Geo(N)
    k = np.arange(N // 2 + 1)
    for c0 in range(0, nch, 2):
        x = np.stack([signals.channel(c0 + i, calls * hop, noise=noise) for i in range(2)])
        hist2 = np.zeros((N, 2), np.float32)
        acc2 = np.zeros((N, 2), np.float32)
        hist = np.zeros((2, N), np.float32)
        for m in range(calls):
            blk = x[:, m * hop:(m + 1) * hop]
            cap = {}
            rk.step(g, hist2, acc2, blk, m * hop, np.float32(pf), hop, capture=cap)
            hist = np.concatenate([hist[:, hop:], blk], axis=1)
            t = (m * hop) % N
            for ch in range(2):
                fr = oracle.frame(hist[ch], np.float32(pf), 0.0)
                # undo the ring alignment (U[k] = X[k] e^{-j 2 pi k t / N}) and the 2x scale of the split
                x32 = cap["X"][ch].astype(np.complex128) / 2.0 * np.exp(2j * np.pi * k * t / N)
                yield x32, fr["spectrum"][:N // 2 + 1], fr["peaks"]


def _peaks(m):
    c = m[2:-2]
    pk = (c > m[:-4]) & (c > m[1:-3]) & (c > m[3:-1]) & (c > m[4:])
    return np.nonzero(pk)[0] + 2


def _uncertain(m, a, cc):
    """ring_peak_masks_guarded: D^2 <= q (8 a^2 q + 16 c^2 S) for some candidate bin"""
    f = np.float32
    S = f(np.sum(m[:-1], dtype=np.float32))
    kappa = f(16.0 * cc * cc) * S
    rho = f(8.0 * a * a)
    c = m[2:-2]
    nbm = np.maximum(np.maximum(m[:-4], m[1:-3]), np.maximum(m[3:-1], m[4:]))
    D, q = c - nbm, c + nbm
    u = D * D - q * (rho * q + kappa)
    return bool((u < 0).any())


@pytest.mark.parametrize("N,hop", [(256, 64), (1024, 256), (2048, 512), (4096, 1024)])
def test_error_model_and_completeness(oracle, N, hop):
    a, cc = _kernel_constants()
    worst_c, worst_a, differing, missed, total = 0.0, 0.0, 0, 0, 0
    for noise in (0.1, 1e-4, 0.0):
        for x32, x64, peaks64 in _frames(oracle, N, hop, noise, 4, 7):
            S = float(np.sum(np.abs(x64) ** 2))
            if S == 0.0:
                continue
            total += 1
            d, A = np.abs(x32 - x64), np.abs(x64)
            worst_c = max(worst_c, float(np.max((d - a * A) / np.sqrt(S))))
            strong = A > 3.0 * np.sqrt(S / A.size)
            if strong.any():
                worst_a = max(worst_a, float(np.max(d[strong] / A[strong])))
            m32 = (x32.real.astype(np.float32) ** 2 + x32.imag.astype(np.float32) ** 2).astype(np.float32)
            if not np.array_equal(_peaks(m32), peaks64):
                differing += 1
                missed += not _uncertain(m32, a, cc)
    print(f"N={N}: {total} frames, c needed {worst_c:.3e} (kernel uses {cc:.1e}), relative error of strong bins "
          f"{worst_a:.3e} (kernel uses a = {a:.1e}); {differing} frames with a different float32 peak set, "
          f"{missed} of them not flagged")
    assert worst_c <= cc and worst_a <= a
    assert differing > 0, "the clean-tone frames must exercise the guard"
    assert missed == 0


def test_guard_is_rare_on_benchmark_input():
    """numpy's float32 FFT stands in for the kernel's (same error level): fraction of steady-state
    frames of the benchmark signal (0.1 broadband floor) the criterion sends to the exact path"""
    a, cc = _kernel_constants()
    N, hop = 1024, 256
    win = (0.5 * (1 - np.cos(2 * np.pi * np.arange(N) / N))).astype(np.float32)
    flagged = total = 0
    for c in range(24):
        x = signals.channel(c, 40 * hop)
        for m in range(4, 40):
            xw = (x[m * hop - N + hop:(m + 1) * hop] * win).astype(np.float32)
            X = np.fft.rfft(xw)
            assert X.dtype == np.complex64
            mm = (X.real ** 2 + X.imag ** 2).astype(np.float32)
            flagged += _uncertain(mm, a, cc)
            total += 1
    print(f"{flagged} of {total} frames flagged ({100.0 * flagged / total:.2f} %)")
    assert flagged <= 0.02 * total
