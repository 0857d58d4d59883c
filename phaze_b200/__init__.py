"""phaze_b200 — B200-native batched phase-vocoder pitch shifter.

A drop-in for one path of olvb/phaze: process(inputs, outputs, {pitchFactor})
(src/ola-processor.js:159-171, src/phase-vocoder.js:45-72).  All compute lives in
hand-written sm_100a CUDA kernels behind the C ABI of include/phaze_b200.h
(libphaze_b200.so); this package is the thin host layer over it.
"""
from ._lib import PhazeError, load as load_library  # noqa: F401
from .processor import (BatchedPhaseVocoder, MultiDevicePhaseVocoder, PhaseVocoderProcessor, PROCESSOR_NAME,  # noqa: F401
                        BUFFERED_BLOCK_SIZE, WEBAUDIO_BLOCK_SIZE)

__all__ = ["BatchedPhaseVocoder", "MultiDevicePhaseVocoder", "PhaseVocoderProcessor", "PhazeError", "PROCESSOR_NAME",
           "BUFFERED_BLOCK_SIZE", "WEBAUDIO_BLOCK_SIZE", "load_library"]
