"""ctypes view of the CPU oracle (oracle/phaze_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  phaze_b200/ never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libphaze_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Returns the path of the .so."""
    src = os.path.join(_HERE, "phaze_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    f32p, f64p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.pvo_create.restype = C.c_void_p
    L.pvo_create.argtypes = [C.c_int, C.c_int, C.c_int]
    L.pvo_destroy.argtypes = [C.c_void_p]
    L.pvo_process.restype = C.c_int
    L.pvo_process.argtypes = [C.c_void_p, f32p, f32p, C.c_float]
    L.pvo_resize.argtypes = [C.c_void_p, C.c_int]
    L.pvo_time_cursor.restype = C.c_double
    L.pvo_time_cursor.argtypes = [C.c_void_p]
    L.pvo_set_time_cursor.argtypes = [C.c_void_p, C.c_double]
    L.pvo_max_source_bin.argtypes = [C.c_void_p]
    L.pvo_fft_real_transform.argtypes = [C.c_int, f32p, f64p]
    L.pvo_fft_inverse_transform.argtypes = [C.c_int, f64p, f64p]
    L.pvo_fft_complete_spectrum.argtypes = [C.c_int, f64p]
    L.pvo_frame.argtypes = [C.c_int, f32p, C.c_float, C.c_double, f32p, f64p, f32p, i32p, i32p, f64p]
    _lib = L
    return L


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class OracleProcessor:
    """The reference's PhaseVocoderProcessor, restated on the CPU.

    process_packed(in[C][hop]) -> out[C][hop]; one call == one process() of the
    reference (ola-processor.js:159-171)."""

    def __init__(self, frame_size: int = 2048, hop_size: int = 128, num_channels: int = 1):
        self._L = lib()
        self._h = self._L.pvo_create(frame_size, hop_size, num_channels)
        if not self._h:
            raise ValueError("FFT size must be a power of two and bigger than 1")
        self.frame_size, self.hop_size, self.num_channels = frame_size, hop_size, num_channels

    def close(self):
        if self._h:
            self._L.pvo_destroy(self._h)
            self._h = None

    __del__ = close

    def process_packed(self, block: np.ndarray | None, pitch_factor: float) -> np.ndarray:
        out = np.empty((self.num_channels, self.hop_size), np.float32)
        if block is None:
            self._L.pvo_process(self._h, None, _p(out, C.c_float), np.float32(pitch_factor))
            return out
        block = np.ascontiguousarray(block, np.float32)
        assert block.shape == (self.num_channels, self.hop_size), block.shape
        self._L.pvo_process(self._h, _p(block, C.c_float), _p(out, C.c_float), np.float32(pitch_factor))
        return out

    def run(self, signal: np.ndarray, pitch_factor: float) -> np.ndarray:
        """signal [C][T*hop] -> output [C][T*hop] (T consecutive process() calls)."""
        signal = np.ascontiguousarray(signal, np.float32)
        Cn, total = signal.shape
        assert Cn == self.num_channels and total % self.hop_size == 0
        out = np.empty_like(signal)
        for t in range(total // self.hop_size):
            s = slice(t * self.hop_size, (t + 1) * self.hop_size)
            out[:, s] = self.process_packed(signal[:, s], pitch_factor)
        return out

    def resize(self, num_channels: int):
        self._L.pvo_resize(self._h, num_channels)
        self.num_channels = num_channels

    @property
    def time_cursor(self) -> float:
        return self._L.pvo_time_cursor(self._h)

    @time_cursor.setter
    def time_cursor(self, t: float):
        self._L.pvo_set_time_cursor(self._h, float(t))

    @property
    def max_source_bin(self) -> int:
        return self._L.pvo_max_source_bin(self._h)


def real_transform(x: np.ndarray) -> np.ndarray:
    """fft.realTransform: f32[N] -> complex128[N], bins > N/2 as the reference leaves them."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(2 * x.size, np.float64)
    if lib().pvo_fft_real_transform(x.size, _p(x, C.c_float), _p(out, C.c_double)) != 0:
        raise ValueError("FFT size must be a power of two and bigger than 1")
    return out[0::2] + 1j * out[1::2]


def inverse_transform(spec: np.ndarray) -> np.ndarray:
    spec = np.asarray(spec, np.complex128)
    data = np.empty(2 * spec.size, np.float64)
    data[0::2], data[1::2] = spec.real, spec.imag
    out = np.empty_like(data)
    if lib().pvo_fft_inverse_transform(spec.size, _p(data, C.c_double), _p(out, C.c_double)) != 0:
        raise ValueError("FFT size must be a power of two and bigger than 1")
    return out[0::2] + 1j * out[1::2]


def frame(x: np.ndarray, pitch_factor: float, time_cursor: float) -> dict:
    """One channel body of processOLA on an un-windowed frame; returns all intermediates."""
    x = np.ascontiguousarray(x, np.float32)
    n = x.size
    nb = n // 2 + 1
    out = np.empty(n, np.float32)
    spec = np.empty(2 * n, np.float64)
    mag = np.empty(nb, np.float32)
    peaks = np.empty(nb, np.int32)
    npk = C.c_int32(0)
    sh = np.empty(2 * n, np.float64)
    rc = lib().pvo_frame(n, _p(x, C.c_float), np.float32(pitch_factor), float(time_cursor),
                         _p(out, C.c_float), _p(spec, C.c_double), _p(mag, C.c_float),
                         _p(peaks, C.c_int32), C.byref(npk), _p(sh, C.c_double))
    if rc != 0:
        raise ValueError("FFT size must be a power of two and bigger than 1")
    return {"out": out, "spectrum": spec[0::2] + 1j * spec[1::2], "magnitudes": mag,
            "peaks": peaks[: npk.value].copy(), "shifted": sh[0::2] + 1j * sh[1::2]}
