/*
 * fake_napi.c — a few-hundred-line stand-in for the part of the N-API runtime that addon/phaze_napi.c
 * touches (test infrastructure; see addon/stub/node_api.h).  Values are small tagged structs, a class is
 * a constructor plus a method table, exceptions are a pending message.  The driver (napi_driver.c) plays
 * the role of the JavaScript in addon/phase-vocoder-processor.js.
 */
#include "fake_napi.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct napi_env__ { char pending[512]; int has_pending, range; };
struct napi_callback_info__ { size_t argc; napi_value *argv; napi_value self; };

static struct napi_env__ g_env;
napi_env fake_env(void) { return &g_env; }

static napi_value mk(int type) {
    napi_value v = (napi_value)calloc(1, sizeof(*v));
    v->type = type;
    return v;
}
napi_value fake_number(double x) { napi_value v = mk(napi_number); v->num = x; return v; }
napi_value fake_null(void) { return mk(napi_null); }
napi_value fake_float32_array(float *data, size_t length) {
    napi_value v = mk(napi_object);
    v->is_f32 = 1; v->data = data; v->len = length;
    return v;
}
const char *fake_pending_exception(int *is_range) {
    if (!g_env.has_pending) return NULL;
    if (is_range) *is_range = g_env.range;
    g_env.has_pending = 0;
    return g_env.pending;
}

static napi_status throw_(napi_env env, const char *msg, int range) {
    snprintf(env->pending, sizeof(env->pending), "%s", msg ? msg : "");
    env->has_pending = 1;
    env->range = range;
    return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char *code, const char *msg) { (void)code; return throw_(env, msg, 0); }
napi_status napi_throw_range_error(napi_env env, const char *code, const char *msg) { (void)code; return throw_(env, msg, 1); }

napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t *argc, napi_value *argv,
                             napi_value *this_arg, void **data) {
    (void)env;
    if (argc) {
        const size_t room = *argc;
        for (size_t i = 0; i < room; i++)
            if (argv) argv[i] = i < info->argc ? info->argv[i] : mk(napi_undefined);
        *argc = info->argc;
    }
    if (this_arg) *this_arg = info->self;
    if (data) *data = NULL;
    return napi_ok;
}
napi_status napi_get_value_int32(napi_env env, napi_value v, int32_t *r) {
    (void)env;
    if (!v || v->type != napi_number) return napi_number_expected;
    *r = (int32_t)v->num;
    return napi_ok;
}
napi_status napi_get_value_double(napi_env env, napi_value v, double *r) {
    (void)env;
    if (!v || v->type != napi_number) return napi_number_expected;
    *r = v->num;
    return napi_ok;
}
napi_status napi_wrap(napi_env env, napi_value o, void *native, napi_finalize fin, void *hint, napi_ref *result) {
    (void)env; (void)hint; (void)result;
    if (!o || o->type != napi_object || o->wrapped) return napi_invalid_arg;
    o->wrapped = native;
    o->finalize = fin;
    return napi_ok;
}
napi_status napi_unwrap(napi_env env, napi_value o, void **result) {
    (void)env;
    if (!o || o->type != napi_object || !o->wrapped) return napi_invalid_arg;
    *result = o->wrapped;
    return napi_ok;
}
napi_status napi_remove_wrap(napi_env env, napi_value o, void **result) {
    (void)env;
    if (!o || o->type != napi_object || !o->wrapped) return napi_invalid_arg;
    if (result) *result = o->wrapped;
    o->wrapped = NULL;
    o->finalize = NULL;
    return napi_ok;
}
napi_status napi_typeof(napi_env env, napi_value v, napi_valuetype *result) {
    (void)env;
    *result = v ? (napi_valuetype)v->type : napi_undefined;
    return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env env, napi_value v, napi_typedarray_type *type, size_t *length,
                                     void **data, napi_value *arraybuffer, size_t *byte_offset) {
    (void)env; (void)arraybuffer; (void)byte_offset;
    if (!v || v->type != napi_object || !v->is_f32) return napi_invalid_arg;
    if (type) *type = napi_float32_array;
    if (length) *length = v->len;
    if (data) *data = v->data;
    return napi_ok;
}
napi_status napi_get_boolean(napi_env env, bool value, napi_value *result) {
    (void)env;
    *result = mk(napi_boolean);
    (*result)->boolean = value;
    return napi_ok;
}
napi_status napi_create_double(napi_env env, double value, napi_value *result) { (void)env; *result = fake_number(value); return napi_ok; }
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result) { (void)env; *result = fake_number(value); return napi_ok; }

napi_status napi_define_class(napi_env env, const char *name, size_t length, napi_callback constructor, void *data,
                              size_t property_count, const napi_property_descriptor *properties, napi_value *result) {
    (void)env; (void)length; (void)data;
    napi_value c = mk(napi_function);
    snprintf(c->name, sizeof(c->name), "%s", name);
    c->constructor = constructor;
    c->nprops = property_count < FAKE_MAX_PROPS ? property_count : FAKE_MAX_PROPS;
    memcpy(c->props, properties, c->nprops * sizeof(*properties));
    *result = c;
    return napi_ok;
}
napi_status napi_set_named_property(napi_env env, napi_value object, const char *utf8name, napi_value value) {
    (void)env;
    if (!object || object->nexports >= FAKE_MAX_PROPS) return napi_invalid_arg;
    snprintf(object->export_names[object->nexports], 64, "%s", utf8name);
    object->exports[object->nexports++] = value;
    return napi_ok;
}

/* ---- what the JavaScript side would do ------------------------------------------------------------ */
napi_value fake_load_module(void) {
    napi_value exports = mk(napi_object);
    return napi_register_module_v1(&g_env, exports);
}
napi_value fake_get_export(napi_value exports, const char *name) {
    for (size_t i = 0; exports && i < exports->nexports; i++)
        if (!strcmp(exports->export_names[i], name)) return exports->exports[i];
    return NULL;
}
napi_value fake_new(napi_value cls, size_t argc, napi_value *argv) {
    napi_value self = mk(napi_object);
    self->cls = cls;
    struct napi_callback_info__ info = {argc, argv, self};
    napi_value r = cls->constructor(&g_env, &info);
    return g_env.has_pending ? NULL : (r ? r : self);
}
napi_value fake_call(napi_value self, const char *method, size_t argc, napi_value *argv) {
    for (size_t i = 0; self && self->cls && i < self->cls->nprops; i++)
        if (!strcmp(self->cls->props[i].utf8name, method)) {
            struct napi_callback_info__ info = {argc, argv, self};
            return self->cls->props[i].method(&g_env, &info);
        }
    throw_(&g_env, "fake_napi: no such method", 0);
    return NULL;
}
void fake_collect(napi_value self) {          /* garbage collection of a wrapped object */
    if (self && self->wrapped && self->finalize) self->finalize(&g_env, self->wrapped, NULL);
    if (self) self->wrapped = NULL;
}
