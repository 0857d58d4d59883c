"""Keeps the pin live: runs the reference's own JavaScript (when /root/reference is mounted,
i.e. in the build container, not on the GPU box) for a few calls and requires the C oracle to
reproduce it bit for bit.  The committed fixtures in tests/golden/ come from the same path."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/src/phase-vocoder.js"),
                                reason="reference sources are only mounted in the build container")


def test_interpreter_semantics_spot_checks():
    from oracle import jsmini
    g = jsmini.make_globals()
    src = """
    var a = new Array(4); a.fill(0); a[-2] += 5; a[1] = 7;
    var f = new Float32Array(3); f[0] = 0.1; f[5] = 9;
    var sub = f.subarray(1); sub[0] = 2.5;
    var bits = ((5 >>> 0) & 3) << -1;
    module.exports = {len: a.length, a1: a[1], neg: a[-2], f0: f[0], f1: f[1], flen: f.length,
                      bits: bits, r1: Math.round(2.5), r2: Math.round(-2.5), r3: Math.round(0.49999999999999994),
                      pow: 3 ** 2, mod: -7 % 3, tern: (1 > 2 ? 1 : 2), und: a[9] === undefined};
    """
    out, _ = jsmini.run_module(src, g)
    p = out.props
    assert p["len"] == 4 and p["a1"] == 7 and p["neg"] != p["neg"]          # NaN: undefined + 5
    assert p["f0"] == float(np.float32(0.1)) and p["f1"] == 2.5 and p["flen"] == 3
    assert p["bits"] == float(-2147483648)                                    # 1 << 31 as int32
    assert p["r1"] == 3 and p["r2"] == -2 and p["r3"] == 0
    assert p["pow"] == 9 and p["mod"] == -1 and p["tern"] == 2 and p["und"] is True


def test_reference_source_matches_oracle_bitwise(oracle):
    from oracle import jsmini
    from phaze_b200 import signals
    ref = jsmini.ReferenceProcessor(1024, 256)
    x = signals.channels(40, 1, 6 * 256)
    want = ref.run(x, np.float32(0.8))
    got = oracle.OracleProcessor(1024, 256, 1).run(x, np.float32(0.8))
    assert np.array_equal(got, want)
    assert ref.time_cursor == 6 * 256


def test_reference_class_surface():
    from oracle import jsmini
    cls, _ = jsmini.load_reference()
    desc = cls.get("parameterDescriptors")
    assert desc.items[0].props == {"name": "pitchFactor", "defaultValue": 1.0}
    ref = jsmini.ReferenceProcessor()
    assert (ref.frame, ref.hop) == (2048, 128)       # phase-vocoder.js:6, ola-processor.js:3
