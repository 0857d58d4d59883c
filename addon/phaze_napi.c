/*
 * phaze_napi.c — thin N-API shim over the C ABI of include/phaze_b200.h.
 *
 * NOT COMPILED IN THIS IMAGE: there is no Node toolchain here (no node, no node_api.h).
 * It is the binding a maintainer builds on a machine that has Node >= 12:
 *
 *     cc -shared -fPIC -I"$(node -p 'process.release.headersUrl' ...)/include/node" \
 *        -I../include phaze_napi.c -L../phaze_b200 -lphaze_b200 -o phaze_b200.node
 *
 * Pure node_api.h C API (ABI-stable across Node versions), no C++ wrapper.  It exposes a
 * `NativeProcessor` class; addon/phase-vocoder-processor.js wraps it into the reference's
 * PhaseVocoderProcessor surface (src/phase-vocoder.js:16-174).
 *
 *   new NativeProcessor(frameSize, hopSize, numChannels, device)
 *   .processPacked(Float32Array in | null, Float32Array out, pitchFactor)   -> true
 *   .resize(numChannels)  .timeCursor()  .close()
 */
#include <node_api.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "phaze_b200.h"

#define NAPI_OK(env, call)                                            \
    do {                                                              \
        if ((call) != napi_ok) {                                      \
            napi_throw_error((env), NULL, "phaze_b200: N-API failure"); \
            return NULL;                                              \
        }                                                             \
    } while (0)

static napi_value throw_pvb(napi_env env, pvb_processor *p, int32_t rc) {
    const char *msg = p ? pvb_last_error(p) : pvb_last_error(NULL);
    napi_throw_error(env, NULL, (msg && msg[0]) ? msg : pvb_error_string(rc));
    return NULL;
}

static void finalize(napi_env env, void *data, void *hint) {
    (void)env; (void)hint;
    pvb_destroy((pvb_processor *)data);
}

static napi_value construct(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4], self;
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, &self, NULL));
    int32_t v[4] = {0, 0, 1, -1};      /* defaults: 2048 / 128 (phase-vocoder.js:6, ola-processor.js:3) */
    for (size_t i = 0; i < argc && i < 4; i++) napi_get_value_int32(env, argv[i], &v[i]);
    pvb_config cfg = {v[0], v[1], v[2], v[3]};
    pvb_processor *p = NULL;
    int32_t rc = pvb_create(&cfg, &p);
    if (rc != PVB_OK) return throw_pvb(env, NULL, rc);   /* fft.js throws on a bad size too (bundle:6-7) */
    NAPI_OK(env, napi_wrap(env, self, p, finalize, NULL, NULL));
    return self;
}

static pvb_processor *unwrap(napi_env env, napi_value self) {
    void *p = NULL;
    if (napi_unwrap(env, self, &p) != napi_ok || !p) {
        napi_throw_error(env, NULL, "phaze_b200: processor is closed");
        return NULL;
    }
    return (pvb_processor *)p;
}

/* processPacked(in, out, pitchFactor): in/out are Float32Array [numChannels * hop] */
static napi_value process_packed(napi_env env, napi_callback_info info) {
    size_t argc = 3;
    napi_value argv[3], self;
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, &self, NULL));
    pvb_processor *p = unwrap(env, self);
    if (!p) return NULL;
    const size_t need = (size_t)pvb_num_channels(p) * (size_t)pvb_hop_size(p);
    float *in = NULL, *out = NULL;
    size_t n_in = 0, n_out = 0;
    napi_typedarray_type ty;
    napi_valuetype vt;
    NAPI_OK(env, napi_typeof(env, argv[0], &vt));
    if (vt != napi_null && vt != napi_undefined) {       /* null == paused (ola-processor.js:93-100) */
        NAPI_OK(env, napi_get_typedarray_info(env, argv[0], &ty, &n_in, (void **)&in, NULL, NULL));
        if (ty != napi_float32_array || n_in != need) {
            napi_throw_range_error(env, NULL, "phaze_b200: input must be Float32Array[channels*hop]");
            return NULL;
        }
    }
    NAPI_OK(env, napi_get_typedarray_info(env, argv[1], &ty, &n_out, (void **)&out, NULL, NULL));
    if (ty != napi_float32_array || n_out != need) {
        napi_throw_range_error(env, NULL, "phaze_b200: output must be Float32Array[channels*hop]");
        return NULL;
    }
    double pf = 1.0;
    napi_get_value_double(env, argv[2], &pf);
    int32_t rc = pvb_process(p, in, out, (float)pf);     /* AudioParam values are float32 (pv:47) */
    if (rc != PVB_OK) return throw_pvb(env, p, rc);
    napi_value t;
    napi_get_boolean(env, true, &t);                     /* process() returns true (ola:170) */
    return t;
}

static napi_value resize(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1], self;
    NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, &self, NULL));
    pvb_processor *p = unwrap(env, self);
    if (!p) return NULL;
    int32_t n = 0;
    napi_get_value_int32(env, argv[0], &n);
    int32_t rc = pvb_resize(p, n);
    if (rc != PVB_OK) return throw_pvb(env, p, rc);
    return self;
}

static napi_value time_cursor(napi_env env, napi_callback_info info) {
    napi_value self, v;
    NAPI_OK(env, napi_get_cb_info(env, info, NULL, NULL, &self, NULL));
    pvb_processor *p = unwrap(env, self);
    if (!p) return NULL;
    napi_create_double(env, pvb_time_cursor(p), &v);
    return v;
}

static napi_value num_channels(napi_env env, napi_callback_info info) {
    napi_value self, v;
    NAPI_OK(env, napi_get_cb_info(env, info, NULL, NULL, &self, NULL));
    pvb_processor *p = unwrap(env, self);
    if (!p) return NULL;
    napi_create_int32(env, pvb_num_channels(p), &v);
    return v;
}

static napi_value close_(napi_env env, napi_callback_info info) {
    napi_value self;
    void *p = NULL;
    NAPI_OK(env, napi_get_cb_info(env, info, NULL, NULL, &self, NULL));
    if (napi_remove_wrap(env, self, &p) == napi_ok && p) pvb_destroy((pvb_processor *)p);
    return NULL;
}

NAPI_MODULE_INIT() {
    napi_property_descriptor props[] = {
        {"processPacked", NULL, process_packed, NULL, NULL, NULL, napi_default, NULL},
        {"resize", NULL, resize, NULL, NULL, NULL, napi_default, NULL},
        {"timeCursor", NULL, time_cursor, NULL, NULL, NULL, napi_default, NULL},
        {"numChannels", NULL, num_channels, NULL, NULL, NULL, napi_default, NULL},
        {"close", NULL, close_, NULL, NULL, NULL, napi_default, NULL},
    };
    napi_value cls;
    if (napi_define_class(env, "NativeProcessor", NAPI_AUTO_LENGTH, construct, NULL,
                          sizeof(props) / sizeof(props[0]), props, &cls) != napi_ok)
        return NULL;
    napi_set_named_property(env, exports, "NativeProcessor", cls);
    return exports;
}
