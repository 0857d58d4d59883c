"""Host-side mirror of the reference's processor interface, over the C ABI.

The reference's public surface for this path is one class,
``PhaseVocoderProcessor extends OLAProcessor`` (src/phase-vocoder.js:16-176,
src/ola-processor.js:6-176), registered as "phase-vocoder-processor", with

    static get parameterDescriptors()        -> [{name: 'pitchFactor', defaultValue: 1.0}]
    constructor(options)                     options.numberOfInputs / numberOfOutputs
    process(inputs, outputs, parameters)     -> true

Node is not available in this image, so this module is the host layer the tests drive;
``addon/`` holds the N-API shim + JS class a Node host would load instead (same C ABI).
Names and argument meaning follow the JS; arrays are numpy float32 instead of Float32Array.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib

PROCESSOR_NAME = "phase-vocoder-processor"       # phase-vocoder.js:176
BUFFERED_BLOCK_SIZE = 2048                       # phase-vocoder.js:6
WEBAUDIO_BLOCK_SIZE = 128                        # ola-processor.js:3


class BatchedPhaseVocoder:
    """One C-ABI handle: `num_channels` independent mono streams, packed [C][hop] blocks.

    This is the fast path a host with many streams uses (pvb_process / pvb_process_device).
    """

    def __init__(self, num_channels: int, frame_size: int = BUFFERED_BLOCK_SIZE,
                 hop_size: int = WEBAUDIO_BLOCK_SIZE, device: int = -1, **options):
        self._lib = _lib.load()
        cfg = _lib.PvbConfig(frame_size, hop_size, num_channels, device)
        h = C.c_void_p()
        rc = self._lib.pvb_create(C.byref(cfg), C.byref(h))
        if rc != _lib.PVB_OK:
            _lib.check(None, rc)
        self._h = h
        self.frame_size = self._lib.pvb_frame_size(h)
        self.hop_size = self._lib.pvb_hop_size(h)
        for name, value in options.items():
            self.set_option(name, value)

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.pvb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- properties -------------------------------------------------------------------
    @property
    def num_channels(self) -> int:
        return self._lib.pvb_num_channels(self._h)

    @property
    def time_cursor(self) -> float:
        return self._lib.pvb_time_cursor(self._h)

    @time_cursor.setter
    def time_cursor(self, samples: float):
        _lib.check(self._h, self._lib.pvb_set_time_cursor(self._h, float(samples)))

    @property
    def kernel_launches(self) -> int:
        return self._lib.pvb_kernel_launches(self._h)

    @property
    def ring_stuck_count(self) -> int:
        """channel pairs whose completion flag never arrived (0 on a healthy handle)"""
        return self._lib.pvb_ring_stuck_count(self._h)

    @property
    def peak_guard_count(self) -> int:
        """channel frames whose peak set was re-decided with the float64 transform so far"""
        return self._lib.pvb_peak_guard_count(self._h)

    _OPTIONS = {"kernel": _lib.PVB_OPT_KERNEL, "launch_mode": _lib.PVB_OPT_LAUNCH_MODE,
                "inputs_ready": _lib.PVB_OPT_INPUTS_READY, "peak_guard": _lib.PVB_OPT_PEAK_GUARD,
                "many_mode": _lib.PVB_OPT_MANY_MODE}
    _KERNELS = {"auto": 0, "ring": 1, "warp": 2, "cta": 3, "generic": 4}

    def set_option(self, name: str, value) -> None:
        """pvb_set_option: kernel ("auto" | "ring" | "warp" | "cta" | "generic"), launch_mode (0 flags,
        1 grid-wide wait, 2 plain launches), inputs_ready (0 | 1), peak_guard (0 auto, 1 off, 2 always,
        3 strict), many_mode (0: one launch per call, 1: consecutive calls of a submission share launches)."""
        if name == "kernel" and isinstance(value, str):
            value = self._KERNELS[value]
        _lib.check(self._h, self._lib.pvb_set_option(self._h, self._OPTIONS[name], int(value)))

    def get_option(self, name: str) -> int:
        return int(self._lib.pvb_get_option(self._h, self._OPTIONS[name]))

    def kernel_name(self, pitch_factor: float) -> str:
        return self._lib.pvb_kernel_name(self._h, np.float32(pitch_factor)).decode()

    # -- process ----------------------------------------------------------------------
    def process(self, block: np.ndarray | None, pitch_factor: float,
                out: np.ndarray | None = None) -> np.ndarray:
        """One process() call on host arrays: block [C][hop] float32 (None == paused)."""
        Cn, hop = self.num_channels, self.hop_size
        if out is None:
            out = np.empty((Cn, hop), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (Cn, hop)
        inp = None
        if block is not None:
            block = np.ascontiguousarray(block, np.float32)
            if block.shape != (Cn, hop):
                raise ValueError(f"expected block of shape {(Cn, hop)}, got {block.shape}")
            inp = block.ctypes.data
        rc = self._lib.pvb_process(self._h, inp, out.ctypes.data, np.float32(pitch_factor))
        _lib.check(self._h, rc)
        return out

    def process_pf(self, block: np.ndarray | None, pitch_factors: np.ndarray,
                   out: np.ndarray | None = None) -> np.ndarray:
        """One process() call with one pitch factor per channel (pvb_process_pf): channel c is
        processed exactly as a processor with the scalar pitch_factors[c] would."""
        Cn, hop = self.num_channels, self.hop_size
        pf = np.ascontiguousarray(pitch_factors, np.float32)
        if pf.shape != (Cn,):
            raise ValueError(f"expected {Cn} pitch factors, got shape {pf.shape}")
        if out is None:
            out = np.empty((Cn, hop), np.float32)
        inp = None
        if block is not None:
            block = np.ascontiguousarray(block, np.float32)
            if block.shape != (Cn, hop):
                raise ValueError(f"expected block of shape {(Cn, hop)}, got {block.shape}")
            inp = block.ctypes.data
        rc = self._lib.pvb_process_pf(self._h, inp, out.ctypes.data, pf.ctypes.data)
        _lib.check(self._h, rc)
        return out

    def process_pf_device(self, in_ptr: int | None, out_ptr: int, pitch_factors: np.ndarray,
                          stream: int | None = None):
        """Asynchronous per-channel-pitch call on device pointers; pitch_factors is a host array."""
        pf = np.ascontiguousarray(pitch_factors, np.float32)
        assert pf.shape == (self.num_channels,)
        _lib.check(self._h, self._lib.pvb_process_pf_device(self._h, in_ptr, out_ptr, pf.ctypes.data, stream))

    def run_pf(self, signal: np.ndarray, pitch_factors: np.ndarray) -> np.ndarray:
        """signal [C][T*hop] -> [C][T*hop] with per-channel pitch factors ([C], or [T][C] per call)."""
        signal = np.ascontiguousarray(signal, np.float32)
        Cn, total = signal.shape
        hop = self.hop_size
        T = total // hop
        pf = np.asarray(pitch_factors, np.float32)
        out = np.empty_like(signal)
        for k in range(T):
            out[:, k * hop:(k + 1) * hop] = self.process_pf(signal[:, k * hop:(k + 1) * hop],
                                                            pf[k] if pf.ndim == 2 else pf)
        return out

    def process_many(self, blocks: np.ndarray, pitch_factor: float,
                     out: np.ndarray | None = None) -> np.ndarray:
        """K consecutive process() calls: blocks [K][C][hop] float32."""
        blocks = np.ascontiguousarray(blocks, np.float32)
        K = blocks.shape[0]
        if blocks.shape != (K, self.num_channels, self.hop_size):
            raise ValueError(f"expected [K][{self.num_channels}][{self.hop_size}], got {blocks.shape}")
        if out is None:
            out = np.empty_like(blocks)
        rc = self._lib.pvb_process_many(self._h, blocks.ctypes.data, out.ctypes.data, K,
                                        np.float32(pitch_factor))
        _lib.check(self._h, rc)
        return out

    def run(self, signal: np.ndarray, pitch_factor: float, calls_per_submit: int = 16) -> np.ndarray:
        """signal [C][T*hop] -> [C][T*hop], i.e. T consecutive process() calls."""
        signal = np.ascontiguousarray(signal, np.float32)
        Cn, total = signal.shape
        hop = self.hop_size
        assert Cn == self.num_channels and total % hop == 0
        T = total // hop
        blocks = np.ascontiguousarray(signal.reshape(Cn, T, hop).transpose(1, 0, 2))
        out = np.empty_like(blocks)
        for k in range(0, T, calls_per_submit):
            self.process_many(blocks[k:k + calls_per_submit], pitch_factor, out[k:k + calls_per_submit])
        return np.ascontiguousarray(out.transpose(1, 0, 2)).reshape(Cn, total)

    def process_device(self, in_ptr: int | None, out_ptr: int, pitch_factor: float,
                       stream: int | None = None, num_calls: int = 1):
        """Asynchronous call(s) on device pointers (e.g. torch tensor .data_ptr())."""
        rc = self._lib.pvb_process_many_device(self._h, in_ptr, out_ptr, num_calls,
                                               np.float32(pitch_factor), stream)
        _lib.check(self._h, rc)

    def sync(self):
        _lib.check(self._h, self._lib.pvb_sync(self._h))

    def resize(self, num_channels: int):
        _lib.check(self._h, self._lib.pvb_resize(self._h, num_channels))

    def reset(self):
        _lib.check(self._h, self._lib.pvb_reset(self._h))

    # -- checkpoint / resume ----------------------------------------------------------
    def get_state(self) -> dict:
        Cn, N = self.num_channels, self.frame_size
        blob = np.empty((2, Cn, N), np.float32)
        _lib.check(self._h, self._lib.pvb_get_state(self._h, blob.ctypes.data))
        return {"input_history": blob[0].copy(), "output_accumulator": blob[1].copy(),
                "time_cursor": self.time_cursor}

    def set_state(self, state: dict):
        Cn, N = self.num_channels, self.frame_size
        blob = np.empty((2, Cn, N), np.float32)
        blob[0] = state["input_history"]
        blob[1] = state["output_accumulator"]
        _lib.check(self._h, self._lib.pvb_set_state(self._h, blob.ctypes.data))
        self.time_cursor = state["time_cursor"]


class MultiDevicePhaseVocoder:
    """One logical processor sharded over several GPUs of a node inside one process (pvb_multi_*):
    `devices` lists the CUDA ordinals, shard i owns a contiguous pair-aligned block of channels on
    devices[i].  Bit-identical to a single BatchedPhaseVocoder with all the channels."""

    def __init__(self, num_channels: int, frame_size: int = BUFFERED_BLOCK_SIZE,
                 hop_size: int = WEBAUDIO_BLOCK_SIZE, devices: Sequence[int] = (0,), **options):
        self._lib = _lib.load()
        cfg = _lib.PvbConfig(frame_size, hop_size, num_channels, -1)
        ids = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = self._lib.pvb_multi_create(C.byref(cfg), ids, len(devices), C.byref(h))
        if rc != _lib.PVB_OK:
            raise _lib.PhazeError(rc, (self._lib.pvb_multi_last_error(None) or b"").decode())
        self._h = h
        self.num_channels, self.frame_size, self.hop_size = num_channels, frame_size, hop_size
        for name, value in options.items():
            if name == "kernel" and isinstance(value, str):
                value = BatchedPhaseVocoder._KERNELS[value]
            self._check(self._lib.pvb_multi_set_option(h, BatchedPhaseVocoder._OPTIONS[name], int(value)))

    def _check(self, rc):
        if rc != _lib.PVB_OK:
            raise _lib.PhazeError(rc, (self._lib.pvb_multi_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pvb_multi_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def shards(self):
        """[(first_channel, num_channels)] per device"""
        out = []
        for i in range(self._lib.pvb_multi_num_devices(self._h)):
            lo, n = C.c_int32(), C.c_int32()
            self._lib.pvb_multi_shard(self._h, i, C.byref(lo), C.byref(n))
            out.append((lo.value, n.value))
        return out

    def process_many(self, blocks: np.ndarray | None, pitch_factor: float, num_calls: int | None = None) -> np.ndarray:
        """blocks [K][C][hop] on the host (None == paused, give num_calls) -> [K][C][hop]"""
        K = num_calls if blocks is None else blocks.shape[0]
        inp = None
        if blocks is not None:
            blocks = np.ascontiguousarray(blocks, np.float32)
            assert blocks.shape == (K, self.num_channels, self.hop_size)
            inp = blocks.ctypes.data
        out = np.empty((K, self.num_channels, self.hop_size), np.float32)
        self._check(self._lib.pvb_multi_process_many(self._h, inp, out.ctypes.data, K, np.float32(pitch_factor)))
        return out

    def process(self, block: np.ndarray | None, pitch_factor: float) -> np.ndarray:
        return self.process_many(None if block is None else block[None], pitch_factor, 1)[0]

    def process_root(self, in_ptr: int | None, out_ptr: int, pitch_factor: float, num_calls: int = 1):
        """in / out: device pointers on devices[0], [num_calls][C][hop]; synchronous"""
        self._check(self._lib.pvb_multi_process_root(self._h, in_ptr, out_ptr, num_calls, np.float32(pitch_factor)))


class PhaseVocoderProcessor:
    """Drop-in mirror of the reference class (phase-vocoder.js:16-174).

    ``process(inputs, outputs, parameters)`` takes the Web Audio nesting:
    inputs[i][j] / outputs[i][j] are float32 arrays of one render quantum for input i,
    channel j; parameters["pitchFactor"] is an array of length 1 or hop whose LAST
    element is used (phase-vocoder.js:47).  Returns True (ola-processor.js:170).
    """

    @staticmethod
    def parameterDescriptors():
        return [{"name": "pitchFactor", "defaultValue": 1.0}]        # phase-vocoder.js:17-22

    def __init__(self, options: dict | None = None):
        options = dict(options or {})
        # The reference overwrites processorOptions with {blockSize: 2048} (phase-vocoder.js:25-27)
        # and fixes the hop at 128 (ola-processor.js:15).  frameSize / hopSize are an
        # extension; leaving them out reproduces the reference.
        po = options.get("processorOptions") or {}
        self.blockSize = int(po.get("frameSize", BUFFERED_BLOCK_SIZE))
        self.hopSize = int(po.get("hopSize", WEBAUDIO_BLOCK_SIZE))
        self.nbInputs = int(options.get("numberOfInputs", 1))          # ola-processor.js:10
        self.nbOutputs = int(options.get("numberOfOutputs", 1))        # ola-processor.js:11
        self.nbOverlaps = self.blockSize // self.hopSize               # ola-processor.js:17
        self.fftSize = self.blockSize
        self._device = int(options.get("device", -1))
        # one handle per input, 1 channel each until we know more (ola-processor.js:23-26)
        self._inputs = [BatchedPhaseVocoder(1, self.blockSize, self.hopSize, self._device)
                        for _ in range(self.nbInputs)]

    @property
    def timeCursor(self) -> float:
        return self._inputs[0].time_cursor if self._inputs else 0.0

    def close(self):
        for h in self._inputs:
            h.close()
        self._inputs = []

    def process(self, inputs: Sequence[Sequence[np.ndarray]], outputs: Sequence[Sequence[np.ndarray]],
                parameters: dict) -> bool:
        pf_arr = np.asarray(parameters["pitchFactor"], np.float32).reshape(-1)
        pitch_factor = pf_arr[-1]                                       # phase-vocoder.js:47
        # paused playback: zero-length blocks on the first input (ola-processor.js:93)
        paused = len(inputs[0]) > 0 and len(inputs[0][0]) == 0
        for i in range(self.nbInputs):
            chans = inputs[i]
            h = self._inputs[i]
            if len(chans) != h.num_channels:                            # ola-processor.js:38-45
                cursor = h.time_cursor
                h.resize(len(chans))
                h.time_cursor = cursor
            if len(chans) == 0:
                h.process(None, pitch_factor)                           # keeps timeCursor in step
                continue
            if len(outputs[i]) < len(chans):
                raise TypeError("outputs[%d] has fewer channels than inputs[%d]" % (i, i))
            block = None if paused else np.stack([np.asarray(c, np.float32) for c in chans])
            res = h.process(block, pitch_factor)
            for j in range(len(chans)):                                 # ola-processor.js:111-118
                outputs[i][j][...] = res[j]
        return True                                                     # ola-processor.js:170
