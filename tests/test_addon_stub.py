"""The N-API glue (addon/phaze_napi.c) compiled and driven without Node.

No Node toolchain exists in this image, so the shim is compiled against a stub node_api.h that declares
the subset of N-API it uses (addon/stub/node_api.h) and linked with a small stand-in runtime
(addon/stub/fake_napi.c); addon/stub/napi_driver.c plays the JavaScript of
addon/phase-vocoder-processor.js: construct, processPacked with Float32Arrays (one paused call), RangeError
on a wrong length, resize, timeCursor, close, finalisation."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "addon", "stub")


def _build():
    subprocess.run(["make", "-C", STUB, "-s"], check=True)
    return os.path.join(STUB, "napi_driver")


def test_addon_compiles_and_reports_library_errors_through_napi():
    """CPU part: the shim compiles warning-free against the stub header, exports NativeProcessor, turns the
    library's error codes into JavaScript exceptions (bad frame size; no CUDA device in this container)."""
    exe = _build()
    r = subprocess.run([exe, "1024", "256", "2", "4", "0.8", os.devnull, os.devnull], capture_output=True, text=True)
    print(r.stdout)
    assert "FAIL" not in r.stdout
    assert "ok: module exports the class NativeProcessor" in r.stdout
    assert "ok: constructor throws on a frame size that is not a power of two" in r.stdout
    assert r.returncode in (0, 2, 3)          # 3: no GPU (the CUDA error came through); 2: GPU box, no input file


@pytest.mark.gpu
@pytest.mark.parametrize("frame,hop,channels,pf", [(1024, 256, 5, 0.8), (2048, 128, 2, 1.2)])
def test_addon_process_matches_ctypes_binding(tmp_path, frame, hop, channels, pf):
    """GPU part: the same calls through the N-API shim and through ctypes give the same bits"""
    from phaze_b200 import BatchedPhaseVocoder, signals
    exe = _build()
    calls = 9
    x = signals.channels(400, channels, calls * hop)
    blocks = np.ascontiguousarray(x.reshape(channels, calls, hop).transpose(1, 0, 2))
    fin, fout = tmp_path / "in.f32", tmp_path / "out.f32"
    blocks.tofile(fin)
    r = subprocess.run([exe, str(frame), str(hop), str(channels), str(calls), repr(float(np.float32(pf))), str(fin), str(fout)],
                       capture_output=True, text=True)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "DRIVER OK" in r.stdout and "FAIL" not in r.stdout
    got = np.fromfile(fout, np.float32).reshape(calls, channels, hop)
    with BatchedPhaseVocoder(channels, frame, hop) as pv:
        want = np.stack([pv.process(None if k == 2 else blocks[k], np.float32(pf)) for k in range(calls)])
    assert np.array_equal(got, want)
