#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN JAVASCRIPT in this container.

    python tests/golden/generate_golden.py          (needs /root/reference; takes ~2 minutes)

The three source files on the hot path (src/ola-processor.js, src/phase-vocoder.js and
fft.js 4.0.3 from the browserify bundle) are executed unmodified by oracle/jsmini.py, a
small JavaScript interpreter written for this purpose (no JS engine exists in the image).
Frame / hop sizes other than the reference's 2048 / 128 are obtained by overriding the two
source constants BUFFERED_BLOCK_SIZE (phase-vocoder.js:6) and WEBAUDIO_BLOCK_SIZE
(ola-processor.js:3) at interpretation time; nothing else is touched.

Every fixture stores the float32 input, the float32 output of consecutive process() calls,
and for the last call the raw float64 spectrum buffer (freqComplexBuffer, including the
stale bins above N/2), the float32 magnitudes and the peak list.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import jsmini                      # noqa: E402
from phaze_b200 import signals                 # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name, frame, hop, channels, pitch factor, calls, first synthetic channel
CASES = [
    ("native_2048_128_pf1.2", None, None, 1, 1.2, 24, 0),       # the reference exactly as shipped
    ("native_2048_128_pf0.8", None, None, 1, 0.8, 20, 1),
    ("c1_1024_256_pf1.2_mono", 1024, 256, 1, 1.2, 12, 2),       # BASELINE config 1
    ("c2_1024_256_pf0.8", 1024, 256, 2, 0.8, 12, 3),            # BASELINE config 2 arithmetic
    ("c3_2048_512_pf1.5_stereo", 2048, 512, 2, 1.5, 10, 5),     # BASELINE config 3 arithmetic
    ("c4_1024_256_pf1.25", 1024, 256, 1, 1.25, 10, 7),          # BASELINE config 4 arithmetic
    ("deep_1024_256_pf0.5", 1024, 256, 1, 0.5, 8, 8),           # reads the deeper stale levels
    ("deep_1024_256_pf0.36", 1024, 256, 1, 0.36, 8, 14),        # ... and the quarter 3N/4 + o (the demo's low end: 0.5 / 1.5)
    ("deep_native_2048_128_pf0.6", None, None, 1, 0.6, 20, 15), # the reference as shipped, low end of its pitch slider
    ("deep_256_64_pf0.4", 256, 64, 2, 0.4, 12, 16),
    ("sweep_256_64_pf0.8", 256, 64, 2, 0.8, 12, 9),             # BASELINE config 5 ends
    ("sweep_512_128_pf1.2", 512, 128, 1, 1.2, 10, 11),
    ("sweep_4096_1024_pf0.8", 4096, 1024, 1, 0.8, 6, 12),
    ("unity_1024_256_pf1.0", 1024, 256, 1, 1.0, 10, 13),
]


# ill-conditioned input (round 2): name, frame, hop, pitch factor, calls, signal maker(num_samples, frame)
TONAL_CASES = [
    ("tonal_1024_256_pf0.8_clean", 1024, 256, 0.8, 12,
     lambda n, N: np.stack([signals.channel(30 + c, n, noise=0.0) for c in range(2)])),
    ("tonal_2048_128_pf0.8_clean", None, None, 0.8, 20,
     lambda n, N: np.stack([signals.channel(32, n, noise=0.0)])),
    ("tonal_1024_256_pf1.2_noise1e-5", 1024, 256, 1.2, 10,
     lambda n, N: np.stack([signals.channel(33, n, noise=1e-5)])),
    ("sine_1024_256_pf0.8_bin40", 1024, 256, 0.8, 10,
     lambda n, N: np.stack([signals.bin_centred_sine(n, N, 40)])),
    ("silence_then_tone_1024_256_pf0.8", 1024, 256, 0.8, 12,
     lambda n, N: np.stack([signals.silence_then_tone(70, n, 4 * 256 + 17)])),
]


def run_case(name, frame, hop, channels, pf, calls, first, make=None):
    ref = jsmini.ReferenceProcessor(frame, hop)
    N, H = ref.frame, ref.hop
    x = signals.channels(first, channels, calls * H) if make is None else make(calls * H, N)
    channels = x.shape[0]
    y = ref.run(x, np.float32(pf))
    obj = ref.obj
    spec = np.array(obj.get("freqComplexBuffer").items, dtype=np.float64)      # last channel, last call
    mags = obj.get("magnitudes").a.copy()
    npk = int(obj.get("nbPeaks"))
    peaks = obj.get("peakIndexes").a[:npk].copy()
    assert ref.time_cursor == calls * H
    np.savez_compressed(os.path.join(HERE, name + ".npz"), frame=N, hop=H, pitch_factor=np.float32(pf),
                        input=x, output=y, last_spectrum=spec, last_magnitudes=mags, last_peaks=peaks)
    print(f"{name}: N={N} hop={H} C={channels} pf={pf} calls={calls} out_rms={np.sqrt((y ** 2).mean()):.4f} "
          f"peaks={npk}")


def run_scenario():
    """paused input (zero-length blocks) and a channel-count change, reference native sizes"""
    ref = jsmini.ReferenceProcessor()
    H = ref.hop
    x = signals.channels(20, 2, 30 * H)
    outs = []
    layout = []
    for t in range(30):
        sl = slice(t * H, (t + 1) * H)
        if 10 <= t < 13:                       # paused: zero-length blocks (ola-processor.js:93)
            ins = [[np.zeros(0, np.float32)]]
            nch = 1
        elif t < 18:
            ins = [[x[0, sl].copy()]]
            nch = 1
        else:                                  # second channel appears (ola-processor.js:38-52)
            ins = [[x[0, sl].copy(), x[1, sl].copy()]]
            nch = 2
        o = [[np.zeros(H, np.float32) for _ in range(nch)]]
        ref.process(ins, o, np.float32(1.3))
        row = np.zeros((2, H), np.float32)
        for c in range(nch):
            row[c] = o[0][c]
        outs.append(row)
        layout.append(0 if 10 <= t < 13 else nch)
    np.savez_compressed(os.path.join(HERE, "scenario_pause_and_channel_change.npz"), frame=ref.frame, hop=H,
                        pitch_factor=np.float32(1.3), input=x, output=np.stack(outs),
                        layout=np.array(layout, np.int32))
    print("scenario: paused calls 10-12, 2 channels from call 18")


if __name__ == "__main__":
    only_tonal = len(sys.argv) > 1 and sys.argv[1] == "tonal"
    if len(sys.argv) > 2 and sys.argv[1] == "only":            # python generate_golden.py only <name prefix>
        for case in CASES:
            if case[0].startswith(sys.argv[2]):
                run_case(*case)
        sys.exit(0)
    if not only_tonal:
        for case in CASES:
            run_case(*case)
        run_scenario()
    for name, frame, hop, pf, calls, make in TONAL_CASES:
        run_case(name, frame, hop, 0, pf, calls, 0, make)
