"use strict";
/*
 * Node-side drop-in for the reference's PhaseVocoderProcessor (src/phase-vocoder.js:16-174,
 * which extends OLAProcessor, src/ola-processor.js:6-176), backed by the B200 CUDA library
 * through addon/phaze_napi.c.  Same class name, same static parameterDescriptors, same
 * process(inputs, outputs, parameters) -> true contract; every channel of every input is one
 * independent mono stream on the GPU.
 *
 * Not runnable in the build image (no Node there); exercised through the identical C ABI by the
 * Python host in phaze_b200/processor.py and tests/test_gpu_api.py.
 */
const { NativeProcessor } = require('./build/Release/phaze_b200.node');

const BUFFERED_BLOCK_SIZE = 2048;   // src/phase-vocoder.js:6
const WEBAUDIO_BLOCK_SIZE = 128;    // src/ola-processor.js:3

class PhaseVocoderProcessor {
    static get parameterDescriptors() {
        return [{ name: 'pitchFactor', defaultValue: 1.0 }];          // phase-vocoder.js:17-22
    }

    constructor(options) {
        // The reference overwrites processorOptions with {blockSize: 2048} and fixes the hop at
        // 128; frameSize / hopSize / device are extensions and default to exactly that.
        const po = (options && options.processorOptions) || {};
        this.blockSize = po.frameSize || BUFFERED_BLOCK_SIZE;
        this.hopSize = po.hopSize || WEBAUDIO_BLOCK_SIZE;
        this.nbInputs = options.numberOfInputs;                        // ola-processor.js:10
        this.nbOutputs = options.numberOfOutputs;                      // ola-processor.js:11
        this.nbOverlaps = this.blockSize / this.hopSize;               // ola-processor.js:17
        this.device = po.device === undefined ? -1 : po.device;
        // one native handle per input, 1 channel until we know more (ola-processor.js:23-26)
        this.native = [];
        this.staging = [];
        for (let i = 0; i < this.nbInputs; i++) this.allocate(i, 1);
    }

    allocate(i, nbChannels) {
        if (this.native[i]) this.native[i].resize(nbChannels);        // state -> 0, cursor kept (ola:38-52)
        else this.native[i] = new NativeProcessor(this.blockSize, this.hopSize, nbChannels, this.device);
        this.staging[i] = {
            in: new Float32Array(nbChannels * this.hopSize),
            out: new Float32Array(nbChannels * this.hopSize),
        };
    }

    get timeCursor() { return this.native.length ? this.native[0].timeCursor() : 0; }

    process(inputs, outputs, parameters) {
        const pf = parameters.pitchFactor[parameters.pitchFactor.length - 1];   // phase-vocoder.js:47
        const paused = inputs[0].length && inputs[0][0].length == 0;            // ola-processor.js:93
        for (let i = 0; i < this.nbInputs; i++) {
            const chans = inputs[i];
            if (chans.length != this.native[i].numChannels()) this.allocate(i, chans.length);
            const st = this.staging[i];
            if (!paused) for (let j = 0; j < chans.length; j++) st.in.set(chans[j], j * this.hopSize);
            this.native[i].processPacked(paused ? null : st.in, st.out, pf);
            for (let j = 0; j < chans.length; j++)                              // ola-processor.js:111-118
                outputs[i][j].set(st.out.subarray(j * this.hopSize, (j + 1) * this.hopSize));
        }
        return true;                                                            // ola-processor.js:170
    }

    /** fast path for hosts with thousands of streams: one packed [C][hop] block per call */
    processPacked(input, output, pitchFactor) {
        return this.native[0].processPacked(input, output, pitchFactor);
    }
}

module.exports = PhaseVocoderProcessor;
// in an AudioWorklet-like host: registerProcessor("phase-vocoder-processor", PhaseVocoderProcessor);
