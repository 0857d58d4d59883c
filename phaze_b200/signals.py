"""Synthetic multichannel audio for tests and benchmarks.

Channel c, sample n:  x = 0.2 * sum_{k<4} a_k sin(2 pi f_k n / 48000 + phi_k) + 0.1 * u[n]
with a_k ~ U(0.2, 1), f_k ~ logU(60, 12000) Hz, phi_k ~ U(0, 2 pi), u ~ U(-1, 1), all drawn
from a counter-based Philox stream keyed by 0xB2000000 + c, so any shard of channels can be
generated independently and reproducibly.  The broadband floor keeps the reference's peak
picking well conditioned (below ~1e-3 its output depends on float64 round-off patterns).
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0xB2000000
SAMPLE_RATE = 48000.0


def channel(c: int, num_samples: int, noise: float = 0.1, tone: float = 0.2) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(key=SEED_BASE + int(c)))
    a = rng.uniform(0.2, 1.0, 4)
    f = np.exp(rng.uniform(np.log(60.0), np.log(12000.0), 4))
    phi = rng.uniform(0.0, 2 * np.pi, 4)
    u = rng.uniform(-1.0, 1.0, num_samples)
    n = np.arange(num_samples, dtype=np.float64)
    x = np.zeros(num_samples)
    for k in range(4):
        x += a[k] * np.sin(2 * np.pi * f[k] * n / SAMPLE_RATE + phi[k])
    return (tone * x + noise * u).astype(np.float32)


def channels(first: int, count: int, num_samples: int, **kw) -> np.ndarray:
    """[count][num_samples] float32 for channels first .. first+count-1."""
    out = np.empty((count, num_samples), np.float32)
    for i in range(count):
        out[i] = channel(first + i, num_samples, **kw)
    return out


def uniform_noise(first: int, count: int, num_samples: int, amp: float = 1.0) -> np.ndarray:
    out = np.empty((count, num_samples), np.float32)
    for i in range(count):
        rng = np.random.Generator(np.random.Philox(key=SEED_BASE + 0x10000000 + first + i))
        out[i] = (amp * rng.uniform(-1.0, 1.0, num_samples)).astype(np.float32)
    return out


def bin_centred_sine(num_samples: int, frame: int, bin_index: int, amp: float = 0.5, phase: float = 0.3) -> np.ndarray:
    """A sine whose period divides the frame: every bin but one is exactly zero in exact arithmetic,
    so the reference's peak picking runs on its own float64 round-off pattern."""
    n = np.arange(num_samples, dtype=np.float64)
    return (amp * np.sin(2 * np.pi * bin_index * n / frame + phase)).astype(np.float32)


def silence_then_tone(c: int, num_samples: int, start: int) -> np.ndarray:
    """Digital silence, then noise-free tones (frames that are partly exact zeros)."""
    x = channel(c, num_samples, noise=0.0)
    x[:start] = 0.0
    return x
