/*
 * phaze_b200.h — C ABI of the B200-native batched phase-vocoder pitch shifter.
 *
 * This library replaces ONE path of olvb/phaze: the per-quantum
 *     process(inputs, outputs, {pitchFactor})
 * of its AudioWorklet processor, i.e.
 *     /root/reference/src/ola-processor.js:159-171   (OLAProcessor.process)
 *     /root/reference/src/phase-vocoder.js:45-72     (PhaseVocoderProcessor.processOLA)
 * including everything those call (Hann window, fft.js realTransform /
 * completeSpectrum / inverseTransform, peak picking, region shifting, the
 * overlap-add ring).  All compute runs in hand-written sm_100a CUDA kernels;
 * there is no CPU fallback: every entry point fails with PVB_ERR_CUDA when no
 * usable device is present.
 *
 * The signatures are plain C (pointers and sizes, no torch / CUDA types) so the
 * same shared object is bound by the Node N-API shim (addon/phaze_napi.c), by
 * ctypes (phaze_b200/_lib.py) and by C/C++ hosts.  One handle == one
 * PhaseVocoderProcessor instance with all of its channels flattened
 * (inputs x channels) into `num_channels` independent mono streams.
 *
 * Threading: like the reference (one audio render thread), a handle is NOT
 * thread-safe; use one caller thread per handle.
 */
#ifndef PHAZE_B200_H
#define PHAZE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVB_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define PVB_API __attribute__((visibility("default")))
#else
#define PVB_API
#endif

/* error codes (0 == success) */
enum {
    PVB_OK = 0,
    PVB_ERR_BAD_SIZE = -1,   /* frame size not a power of two in [256, 4096]  (fft.js ctor throw,
                                www/phase-vocoder.js:6-7), hop does not divide it or hop % 4 != 0 */
    PVB_ERR_BAD_ARG = -2,    /* NULL handle / pointer, negative channel count, bad state blob      */
    PVB_ERR_CUDA = -3,       /* CUDA runtime / driver error, or no sm_100 device                   */
    PVB_ERR_NOMEM = -4       /* host or device allocation failed                                   */
};

typedef struct pvb_processor pvb_processor;

/* Replaces the constructor options of PhaseVocoderProcessor / OLAProcessor
 * (phase-vocoder.js:24-43, ola-processor.js:7-34).  The reference hard-codes
 * frame 2048 (phase-vocoder.js:6) and hop 128 (ola-processor.js:3,15); both are
 * parameters here and frame_size == 0 / hop_size == 0 select those defaults. */
typedef struct pvb_config {
    int32_t frame_size;    /* BUFFERED_BLOCK_SIZE; power of two, 256..4096; 0 -> 2048         */
    int32_t hop_size;      /* WEBAUDIO_BLOCK_SIZE == samples per process() call; 0 -> 128     */
    int32_t num_channels;  /* flattened inputs x channels handled by this handle (>= 0)       */
    int32_t device;        /* CUDA device ordinal; -1 -> current device                       */
} pvb_config;

PVB_API int32_t pvb_version(void);
PVB_API const char *pvb_error_string(int32_t code);

/* new PhaseVocoderProcessor(options) — phase-vocoder.js:24-43.  State (input
 * history, overlap-add accumulator, timeCursor) starts at zero like the JS. */
PVB_API int32_t pvb_create(const pvb_config *cfg, pvb_processor **out);
/* garbage collection of the processor */
PVB_API void pvb_destroy(pvb_processor *p);
/* text of the last error seen on this handle (never NULL) */
PVB_API const char *pvb_last_error(const pvb_processor *p);

/* process(inputs, outputs, parameters) — ola-processor.js:159-171 — on HOST
 * buffers.  in / out: [num_channels][hop_size] float32, packed.  `in` is not
 * modified (ola-processor.js:105 copies first), `out` is fully overwritten
 * (ola-processor.js:115).  in == NULL is the paused case: every channel gets a
 * block of zeros and timeCursor still advances (ola-processor.js:93-100).
 * pitch_factor is the LAST element of parameters.pitchFactor
 * (phase-vocoder.js:47), a float32 like every AudioParam value.
 * Synchronous: `out` is valid on return, as in the JS. */
PVB_API int32_t pvb_process(pvb_processor *p, const float *in, float *out, float pitch_factor);

/* Same call on DEVICE buffers, asynchronous on `stream` (a cudaStream_t passed
 * as void*; NULL -> the handle's own stream).  Used by hosts that already keep
 * audio in HBM (the multi-GPU sharded host, the benchmark).  in == NULL: paused. */
PVB_API int32_t pvb_process_device(pvb_processor *p, const float *in_dev, float *out_dev,
                           float pitch_factor, void *stream);

/* K consecutive process() calls of one pitch factor in a single submission
 * (in/out: [K][num_channels][hop_size]); bit-identical to K pvb_process_device
 * calls.  With PVB_OPT_MANY_MODE = 1 the calls share kernel launches where the
 * ring-order kernel applies (one DRAM round trip of the state per launch instead
 * of per call).  The host variant overlaps the copies of consecutive groups of
 * calls with the kernels. */
PVB_API int32_t pvb_process_many_device(pvb_processor *p, const float *in_dev, float *out_dev,
                                int32_t num_calls, float pitch_factor, void *stream);
PVB_API int32_t pvb_process_many(pvb_processor *p, const float *in, float *out, int32_t num_calls,
                         float pitch_factor);

/* process() with ONE PITCH FACTOR PER CHANNEL: pitch_factors is a HOST array of num_channels float32
 * (both variants; it is small, and the library needs its range to pick a kernel).  The reference takes
 * one scalar per processor and call (phase-vocoder.js:47); a host that runs thousands of independent
 * streams through one handle gives each its own.  Channel c gets exactly what a processor with the
 * scalar pitch_factors[c] would compute.  The array is cached on the device and re-uploaded only when
 * it changes. */
PVB_API int32_t pvb_process_pf(pvb_processor *p, const float *in, float *out, const float *pitch_factors);
PVB_API int32_t pvb_process_pf_device(pvb_processor *p, const float *in_dev, float *out_dev,
                                      const float *pitch_factors, void *stream);

/* wait for everything submitted on the handle's own stream */
PVB_API int32_t pvb_sync(pvb_processor *p);

/* reallocateChannelsIfNeeded — ola-processor.js:38-52: a changed channel count
 * re-allocates the buffers (state -> 0) and keeps timeCursor. */
PVB_API int32_t pvb_resize(pvb_processor *p, int32_t num_channels);
/* back to the freshly constructed state (all zero, timeCursor 0) */
PVB_API int32_t pvb_reset(pvb_processor *p);

/* Per-handle options (none of them changes results beyond what is stated; 0 == default).
 * Returns PVB_ERR_BAD_ARG for an unknown option or value. */
enum {
    /* Which kernel family serves process() calls: 0 auto (the ring-order kernel wherever it applies),
       1 ring-order, 2 one warp per pair in frame order (frame 1024), 3 CTA-cooperative, 4 generic.
       A family that does not cover the handle's frame / hop / pitch factor falls through to the next
       one, exactly like auto.  Exists for tests and A/B measurements.
       Ranges: ring-order -- hop a multiple of 128 samples (frame 256: 64), hop <= frame / 2, pitch factors in
       [0.33, 64] (below 0.75: its DEEP instances; scalar or per channel; below 0.5 colliding regions are added with
       shared-memory atomics, so the last bit of a result can differ from run to run); families
       2 and 3 -- pitch factors in [0.75, 64], any hop with at most 32 overlaps; generic -- everything. */
    PVB_OPT_KERNEL = 1,
    /* How consecutive launches are chained: 0 auto (programmatic dependent launch + per-pair completion
       flags), 1 programmatic dependent launch, whole-grid wait, 2 plain stream-ordered launches. */
    PVB_OPT_LAUNCH_MODE = 2,
    /* Device entry points only.  0 (default): strict stream order -- the first launch of every
       submission waits for everything enqueued on the stream before it (the caller's input may be
       produced by a kernel the library knows nothing about).  1: the caller guarantees that the input
       buffers it passes are complete and visible when the call is SUBMITTED (resident data, or
       produced by work the host has synchronised with); launches then overlap with whatever precedes
       them on the stream as far as the handle's own state allows. */
    PVB_OPT_INPUTS_READY = 3,
    /* Peak picking (phase-vocoder.js:82-116) compares float32 roundings of float64 |X|^2; the kernels
       compute |X|^2 from a float32 FFT and mark every comparison that falls inside the float32 error
       bound of the frame as uncertain.  A channel frame can be RE-DECIDED with a float64 transform that
       follows fft.js operation by operation (bit-identical peak set on any input).
       0 (default): re-decide frames with five or more uncertain comparisons (noise-free tones, band-limited
       material, silence followed by a tone: bins at the round-off floor come in clusters); one or two
       natural near-ties in a broadband frame keep their float32 decision.
       1: never (float32 decisions only, the test itself is skipped).  2: always (tests).
       3 strict: re-decide frames with one or more uncertain comparisons. */
    PVB_OPT_PEAK_GUARD = 4,
    /* pvb_process_many[_device]: 0 (default) one kernel launch per call, chained (see PVB_OPT_LAUNCH_MODE);
       1: consecutive calls share a kernel launch where the ring-order kernel applies (up to 64 calls per
       launch: each channel pair loops over the calls, its state goes through L1 / L2 instead of HBM and
       the rings make one DRAM round trip per launch).  Results are bit-identical.  The chained single
       launches are the faster of the two on B200 (the path is bound by the SMs, not by DRAM, and the loop
       costs the kernel registers), so this is an option, not the default. */
    PVB_OPT_MANY_MODE = 5
};
PVB_API int32_t pvb_set_option(pvb_processor *p, int32_t option, int64_t value);
PVB_API int64_t pvb_get_option(const pvb_processor *p, int32_t option);

/* introspection */
PVB_API int32_t pvb_frame_size(const pvb_processor *p);
PVB_API int32_t pvb_hop_size(const pvb_processor *p);
PVB_API int32_t pvb_num_channels(const pvb_processor *p);
/* this.timeCursor (phase-vocoder.js:31,71): hop_size * (process calls so far) */
PVB_API double pvb_time_cursor(const pvb_processor *p);
PVB_API int32_t pvb_set_time_cursor(pvb_processor *p, double samples);
/* name of the CUDA kernel a process() call with this pitch factor launches on this handle */
PVB_API const char *pvb_kernel_name(const pvb_processor *p, float pitch_factor);
/* number of CUDA kernels this handle has launched since creation */
PVB_API int64_t pvb_kernel_launches(const pvb_processor *p);
/* Diagnostics: consecutive launches of the ring-order kernel synchronise per channel pair through
   completion flags in device memory.  A flag that does not arrive within the spin bound makes the
   pair fall back to waiting for the whole previous grid; if it is still missing after that, the
   pair is NOT processed, the handle's sticky device-error word is set and every later pvb_process* /
   pvb_sync / pvb_get_state on the handle returns PVB_ERR_CUDA (pvb_reset / pvb_resize clear it).
   This returns the number of such pairs so far (0 on a healthy handle); -1 on CUDA error. */
PVB_API int64_t pvb_ring_stuck_count(pvb_processor *p);
/* Diagnostics: number of channel frames whose peak set was re-decided with the float64 transform so
   far (see PVB_OPT_PEAK_GUARD).  Synchronises the handle's streams; -1 on CUDA error. */
PVB_API int64_t pvb_peak_guard_count(pvb_processor *p);

/* checkpoint / resume of the per-channel state.  Blob layout (float32):
 * [num_channels][frame_size] input history in time order (oldest first),
 * then [num_channels][frame_size] overlap-add accumulator in time order, i.e.
 * exactly inputBuffers[..][0..N) and outputBuffers[..][0..N) of the reference
 * after a process() call.  pvb_state_bytes() == 2*C*N*4. */
PVB_API size_t pvb_state_bytes(const pvb_processor *p);
PVB_API int32_t pvb_get_state(pvb_processor *p, float *blob_host);
PVB_API int32_t pvb_set_state(pvb_processor *p, const float *blob_host);

/* ---- one logical processor over several GPUs of a node, in one process --------------------------------
 * (device_ids[], num_devices.)  Channels are independent on this path (phase-vocoder.js:49-53)
 * and timeCursor advances identically everywhere, so shard i owns a contiguous, pair-aligned block of
 * channels on device_ids[i] and no call exchanges anything but the caller's own blocks.  Results are
 * bit-identical to one handle with all the channels.  cfg->device is ignored; a device may be listed
 * more than once (logical shards). */
typedef struct pvb_multi pvb_multi;
PVB_API int32_t pvb_multi_create(const pvb_config *cfg, const int32_t *device_ids, int32_t num_devices,
                                 pvb_multi **out);
PVB_API void pvb_multi_destroy(pvb_multi *m);
PVB_API const char *pvb_multi_last_error(const pvb_multi *m);
PVB_API int32_t pvb_multi_num_devices(const pvb_multi *m);
PVB_API int32_t pvb_multi_num_channels(const pvb_multi *m);
/* the single-device handle of shard `index` and the channels it owns (introspection, options, state) */
PVB_API pvb_processor *pvb_multi_shard(pvb_multi *m, int32_t index, int32_t *first_channel, int32_t *num_channels);
PVB_API int32_t pvb_multi_set_option(pvb_multi *m, int32_t option, int64_t value);
/* process() on HOST buffers [num_channels][hop] (K calls: [K][num_channels][hop]): every device copies its
 * own slab in and out (pinned host memory recommended), all devices run concurrently; synchronous. */
PVB_API int32_t pvb_multi_process(pvb_multi *m, const float *in, float *out, float pitch_factor);
PVB_API int32_t pvb_multi_process_many(pvb_multi *m, const float *in, float *out, int32_t num_calls,
                                       float pitch_factor);
/* The single-root mode of a sharded deployment: in / out are DEVICE buffers on device_ids[0]
 * ([num_calls][num_channels][hop], complete before the call).  Slabs are scattered to the other devices
 * and results gathered back with peer copies on the copy engines over NVLink; synchronous. */
PVB_API int32_t pvb_multi_process_root(pvb_multi *m, const float *in_dev, float *out_dev, int32_t num_calls,
                                       float pitch_factor);

/* pinned host memory for callers that want the host<->device copies of
 * pvb_process() to run at full PCIe speed */
PVB_API void *pvb_alloc_host(size_t bytes);
PVB_API void pvb_free_host(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* PHAZE_B200_H */
