"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol that
include/phaze_b200.h declares, validates its arguments, and fails loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "phaze_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"PVB_API[^;]*?\b(pvb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from phaze_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib._SIGNATURES) == names
    assert lib.pvb_version() >= 100


def test_error_strings_and_bad_config():
    from phaze_b200 import _lib
    lib = _lib.load()
    assert b"power of two" in lib.pvb_error_string(_lib.PVB_ERR_BAD_SIZE)
    h = C.c_void_p()
    for frame, hop in [(1000, 250), (1024, 300), (8192, 2048), (1024, 2)]:
        cfg = _lib.PvbConfig(frame, hop, 4, -1)
        assert lib.pvb_create(C.byref(cfg), C.byref(h)) == _lib.PVB_ERR_BAD_SIZE
        assert not h.value
    cfg = _lib.PvbConfig(1024, 256, -1, -1)
    assert lib.pvb_create(C.byref(cfg), C.byref(h)) == _lib.PVB_ERR_BAD_ARG
    assert lib.pvb_create(None, C.byref(h)) == _lib.PVB_ERR_BAD_ARG
    assert lib.pvb_process(None, None, None, 1.0) == _lib.PVB_ERR_BAD_ARG


def test_no_cpu_fallback():
    """without a CUDA device construction must fail with PVB_ERR_CUDA, never compute on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import phaze_b200
    with pytest.raises(phaze_b200.PhazeError) as e:
        phaze_b200.BatchedPhaseVocoder(4, 1024, 256)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under phaze_b200/ may reference it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "phaze_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("test infrastructure", ""), f


def test_mirror_surface_matches_reference_names():
    import phaze_b200
    P = phaze_b200.PhaseVocoderProcessor
    assert phaze_b200.PROCESSOR_NAME == "phase-vocoder-processor"
    assert P.parameterDescriptors() == [{"name": "pitchFactor", "defaultValue": 1.0}]
    assert (phaze_b200.BUFFERED_BLOCK_SIZE, phaze_b200.WEBAUDIO_BLOCK_SIZE) == (2048, 128)


def test_synthetic_signal_is_reproducible_and_shardable():
    from phaze_b200 import signals
    a = signals.channels(0, 6, 512)
    b = signals.channels(3, 3, 512)
    assert np.array_equal(a[3:], b) and a.dtype == np.float32
    assert np.abs(a).max() < 0.9 and 0.05 < a.std() < 0.5
