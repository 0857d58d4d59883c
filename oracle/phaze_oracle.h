/*
 * phaze_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * reported CPU baseline.  The product path (phaze_b200/) never links or calls it.
 *
 * The oracle is a restatement in C of the reference's JavaScript hot path
 *   /root/reference/src/ola-processor.js      (OLAProcessor)
 *   /root/reference/src/phase-vocoder.js      (PhaseVocoderProcessor)
 *   fft.js 4.0.3 (npm dependency, package.json:33; only copy of its source is
 *   inside the committed bundle /root/reference/www/phase-vocoder.js:2-508)
 * with `double` wherever the JS holds a plain Array / Number and `float`
 * wherever it holds a Float32Array; frame size and hop size are parameters
 * instead of the two source constants (phase-vocoder.js:6, ola-processor.js:3).
 *
 * Parity pinning: the reference ships no tests, fixtures or golden vectors
 * (package.json:20).  The oracle is pinned against the reference SOURCE TEXT
 * executed in this container by oracle/jsmini.py (a small JavaScript
 * interpreter written for exactly this purpose; see tests/golden/README.md),
 * and against mathematical identities (numpy rfft, sub-FFT identities for the
 * stale upper bins, pitchFactor==1 delay identity).
 */
#ifndef PHAZE_ORACLE_H
#define PHAZE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pvo_processor pvo_processor;

/* new PhaseVocoderProcessor(options) with blockSize=frame_size, hop=hop_size and
 * num_channels flattened (inputs x channels) channels.  Returns NULL on a bad
 * size (fft.js throws for non powers of two, bundle:6-7). */
pvo_processor *pvo_create(int frame_size, int hop_size, int num_channels);
void pvo_destroy(pvo_processor *p);

/* process(inputs, outputs, {pitchFactor}) — ola-processor.js:159-171.
 * in/out: [num_channels][hop_size] packed float32.  in == NULL means the paused
 * case (zero-length blocks, ola-processor.js:93-100).  Always returns 1 ("true"). */
int pvo_process(pvo_processor *p, const float *in, float *out, float pitch_factor);

/* reallocateChannelsIfNeeded (ola-processor.js:38-52): state -> 0, timeCursor kept */
int pvo_resize(pvo_processor *p, int num_channels);

double pvo_time_cursor(const pvo_processor *p);
void pvo_set_time_cursor(pvo_processor *p, double t);

/* ---- white-box hooks for the tests ------------------------------------- */

/* fft.realTransform(out, data): data f32[N] (promoted), out f64[2N] interleaved;
 * bins above N/2 hold whatever _realTransform4 leaves there (bundle:306-442). */
int pvo_fft_real_transform(int n, const float *data, double *out);
/* fft.inverseTransform(out, data): full complex, both f64[2N] (bundle:102-114) */
int pvo_fft_inverse_transform(int n, const double *data, double *out);
/* fft.completeSpectrum(spectrum) in place (bundle:69-76) */
int pvo_fft_complete_spectrum(int n, double *spectrum);

/* One pass of the per-channel body of processOLA (phase-vocoder.js:52-67) on a
 * frame of N samples that is NOT yet windowed, at the given timeCursor.
 * frame_out: f32[N] (windowed synthesis frame, before the /nbOverlaps add).
 * Optional outputs (may be NULL): spectrum f64[2N], magnitudes f32[N/2+1],
 * peaks i32[N/2+1] (+ *nb_peaks), shifted f64[2N] (after completeSpectrum). */
int pvo_frame(int n, const float *frame_in, float pitch_factor, double time_cursor,
              float *frame_out, double *spectrum, float *magnitudes,
              int32_t *peaks, int32_t *nb_peaks, double *shifted);

/* last frame statistics of a processor (debug): highest source bin read by
 * shiftPeaks since creation */
int pvo_max_source_bin(const pvo_processor *p);

#ifdef __cplusplus
}
#endif
#endif
