"use strict";
/*
 * examples/node/host_loop.js — the caller of the hot path in a Node deployment: what the browser's audio
 * render loop is for the reference (one process() call per quantum, src/ola-processor.js:159-171), for
 * thousands of independent streams at once.
 *
 *     node examples/node/host_loop.js [channels] [frameSize] [hopSize] [pitchFactor] [seconds]
 *
 * Needs the addon built against Node's own headers (INTEGRATION.md section 2).  Not runnable in the build
 * image (no Node there); the same sequence of calls is driven through the compiled shim by
 * addon/stub/napi_driver.c (tests/test_addon_stub.py).
 */
const PhaseVocoderProcessor = require('../../addon/phase-vocoder-processor.js');

const channels = parseInt(process.argv[2] || '4096', 10);
const frameSize = parseInt(process.argv[3] || '1024', 10);
const hopSize = parseInt(process.argv[4] || '256', 10);
const pitchFactor = parseFloat(process.argv[5] || '0.8');
const seconds = parseFloat(process.argv[6] || '2');
const sampleRate = 48000;

// one processor with one input; its channels are the streams (reference: inputs[i][j], ola:20-33)
const proc = new PhaseVocoderProcessor({
    numberOfInputs: 1, numberOfOutputs: 1, processorOptions: { frameSize, hopSize },
});

// (a) the reference's nested surface, a few channels: inputs[0][j] is one quantum of channel j
{
    const few = 4;
    const inputs = [[]], outputs = [[]];
    for (let j = 0; j < few; j++) {
        const q = new Float32Array(hopSize);
        for (let n = 0; n < hopSize; n++) q[n] = 0.3 * Math.sin(2 * Math.PI * 440 * (j + 1) * n / sampleRate);
        inputs[0].push(q);
        outputs[0].push(new Float32Array(hopSize));
    }
    const params = { pitchFactor: new Float32Array([pitchFactor]) };
    for (let k = 0; k < 2 * frameSize / hopSize; k++) proc.process(inputs, outputs, params);   // -> true
    console.log(`nested surface: ${few} channels, timeCursor ${proc.timeCursor}`);
}

// (b) the packed fast path for many streams: one [channels][hop] Float32Array per call, no per-channel copies
{
    const many = new PhaseVocoderProcessor({
        numberOfInputs: 1, numberOfOutputs: 1, processorOptions: { frameSize, hopSize },
    });
    many.allocate(0, channels);
    const input = new Float32Array(channels * hopSize), output = new Float32Array(channels * hopSize);
    for (let i = 0; i < input.length; i++) input[i] = 0.2 * (Math.random() * 2 - 1);
    const calls = Math.ceil(seconds * sampleRate / hopSize);
    const t0 = process.hrtime.bigint();
    for (let k = 0; k < calls; k++) many.processPacked(input, output, pitchFactor);
    const dt = Number(process.hrtime.bigint() - t0) * 1e-9;
    console.log(`${channels} streams x ${calls} calls: ${(channels * calls / dt).toExponential(3)} frames/s, ` +
                `${(channels * calls * hopSize / sampleRate / dt).toFixed(0)} x realtime`);
}
