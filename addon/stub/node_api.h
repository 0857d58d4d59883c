/*
 * node_api.h — MINIMAL STUB of Node's N-API header, test infrastructure only.
 *
 * The build image has no Node toolchain (no node, no node_api.h), so addon/phaze_napi.c could never be
 * compiled here.  This stub declares exactly the subset of the real N-API that phaze_napi.c uses, with the
 * real names, signatures and enum values (Node >= 12, N-API version 4), so that the glue code is compiled
 * and driven by addon/stub/fake_napi.c + addon/stub/napi_driver.c (tests/test_addon_stub.py).  A real
 * build uses Node's own header; nothing here ships.
 */
#ifndef SRC_NODE_API_H_
#define SRC_NODE_API_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_ref__ *napi_ref;
typedef struct napi_callback_info__ *napi_callback_info;

typedef enum { napi_ok, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected,
               napi_function_expected, napi_number_expected, napi_boolean_expected, napi_array_expected,
               napi_generic_failure, napi_pending_exception } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object,
               napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
               napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array } napi_typedarray_type;
typedef enum { napi_default = 0, napi_writable = 1, napi_enumerable = 2, napi_configurable = 4,
               napi_static = 1 << 10 } napi_property_attributes;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);

typedef struct {
    const char *utf8name;
    napi_value name;
    napi_callback method;
    napi_callback getter;
    napi_callback setter;
    napi_value value;
    napi_property_attributes attributes;
    void *data;
} napi_property_descriptor;

#define NAPI_AUTO_LENGTH SIZE_MAX
#define NAPI_MODULE_INIT() napi_value napi_register_module_v1(napi_env env, napi_value exports)

napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_throw_range_error(napi_env env, const char *code, const char *msg);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv,
                             napi_value *this_arg, void **data);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t *result);
napi_status napi_get_value_double(napi_env env, napi_value value, double *result);
napi_status napi_wrap(napi_env env, napi_value js_object, void *native_object, napi_finalize finalize_cb,
                      void *finalize_hint, napi_ref *result);
napi_status napi_unwrap(napi_env env, napi_value js_object, void **result);
napi_status napi_remove_wrap(napi_env env, napi_value js_object, void **result);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype *result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type,
                                     size_t *length, void **data, napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_get_boolean(napi_env env, bool value, napi_value *result);
napi_status napi_create_double(napi_env env, double value, napi_value *result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result);
napi_status napi_define_class(napi_env env, const char *utf8name, size_t length, napi_callback constructor,
                              void *data, size_t property_count, const napi_property_descriptor *properties,
                              napi_value *result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char *utf8name, napi_value value);

#ifdef __cplusplus
}
#endif
#endif /* SRC_NODE_API_H_ */
