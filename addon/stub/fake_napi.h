/* fake_napi.h — value model and driver-side helpers of the stand-in N-API runtime (test infrastructure) */
#ifndef FAKE_NAPI_H
#define FAKE_NAPI_H
#include "node_api.h"

#define FAKE_MAX_PROPS 16
struct napi_value__ {
    int type;                       /* napi_valuetype */
    double num;
    bool boolean;
    int is_f32;                     /* Float32Array */
    float *data;
    size_t len;
    void *wrapped;                  /* napi_wrap */
    napi_finalize finalize;
    struct napi_value__ *cls;       /* object -> its class */
    /* class */
    char name[64];
    napi_callback constructor;
    size_t nprops;
    napi_property_descriptor props[FAKE_MAX_PROPS];
    /* exports object */
    size_t nexports;
    char export_names[FAKE_MAX_PROPS][64];
    struct napi_value__ *exports[FAKE_MAX_PROPS];
};

napi_value napi_register_module_v1(napi_env env, napi_value exports);   /* NAPI_MODULE_INIT of the addon */

napi_env fake_env(void);
napi_value fake_number(double x);
napi_value fake_null(void);
napi_value fake_float32_array(float *data, size_t length);
const char *fake_pending_exception(int *is_range);       /* returns and clears the pending exception */
napi_value fake_load_module(void);
napi_value fake_get_export(napi_value exports, const char *name);
napi_value fake_new(napi_value cls, size_t argc, napi_value *argv);
napi_value fake_call(napi_value self, const char *method, size_t argc, napi_value *argv);
void fake_collect(napi_value self);
#endif
