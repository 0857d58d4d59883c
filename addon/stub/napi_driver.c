/*
 * napi_driver.c — plays addon/phase-vocoder-processor.js against addon/phaze_napi.c through the stand-in
 * N-API runtime (test infrastructure; tests/test_addon_stub.py builds and runs it).
 *
 *   napi_driver <frame> <hop> <channels> <calls> <pitch> <in.f32> <out.f32>
 *
 * in.f32: [calls][channels][hop] float32; call 2 is made with a null input (paused).  Writes the outputs
 * of all calls to out.f32 and prints one line per checked behaviour; exit code 0 when every N-API-side
 * expectation held (numeric parity is checked by the Python test against the ctypes binding).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fake_napi.h"

static int fails = 0;
#define EXPECT(cond, what) do { printf("%s: %s\n", (cond) ? "ok" : "FAIL", what); if (!(cond)) fails++; } while (0)

int main(int argc, char **argv) {
    if (argc != 8) { fprintf(stderr, "usage: napi_driver frame hop channels calls pitch in.f32 out.f32\n"); return 2; }
    const int frame = atoi(argv[1]), hop = atoi(argv[2]), channels = atoi(argv[3]), calls = atoi(argv[4]);
    const double pitch = atof(argv[5]);
    const size_t block = (size_t)channels * hop;
    napi_value exports = fake_load_module();
    napi_value cls = fake_get_export(exports, "NativeProcessor");
    EXPECT(cls != NULL, "module exports the class NativeProcessor");
    if (!cls) return 1;

    /* a bad frame size must surface as a JavaScript exception from the constructor (fft.js throws too) */
    napi_value bad[4] = {fake_number(1000), fake_number(250), fake_number(1), fake_number(0)};
    napi_value none = fake_new(cls, 4, bad);
    const char *msg = fake_pending_exception(NULL);
    EXPECT(none == NULL && msg != NULL, "constructor throws on a frame size that is not a power of two");
    if (msg) printf("  exception text: %s\n", msg);

    napi_value args[4] = {fake_number(frame), fake_number(hop), fake_number(channels), fake_number(0)};
    napi_value proc = fake_new(cls, 4, args);
    msg = fake_pending_exception(NULL);
    if (!proc) {
        /* no CUDA device (the CPU-only build container): the library's own message must come through */
        printf("constructor exception: %s\n", msg ? msg : "(none)");
        EXPECT(msg && strstr(msg, "CUDA") != NULL, "without a GPU the constructor throws the library's CUDA error");
        return fails ? 1 : 3;
    }
    float *in = (float *)malloc(block * calls * sizeof(float)), *out = (float *)calloc(block * calls, sizeof(float));
    FILE *f = fopen(argv[6], "rb");
    if (!f || fread(in, sizeof(float), block * calls, f) != block * calls) { fprintf(stderr, "cannot read input\n"); return 2; }
    fclose(f);
    for (int k = 0; k < calls; k++) {
        napi_value a[3] = {k == 2 ? fake_null() : fake_float32_array(in + k * block, block),
                           fake_float32_array(out + k * block, block), fake_number(pitch)};
        napi_value r = fake_call(proc, "processPacked", 3, a);
        if (!(r && r->type == napi_boolean && r->boolean)) { EXPECT(0, "processPacked returns true"); break; }
    }
    EXPECT(fake_pending_exception(NULL) == NULL, "no exception during process calls");
    napi_value tc = fake_call(proc, "timeCursor", 0, NULL);
    EXPECT(tc && tc->num == (double)calls * hop, "timeCursor == calls * hop");
    napi_value nc = fake_call(proc, "numChannels", 0, NULL);
    EXPECT(nc && nc->num == channels, "numChannels");
    /* wrong length -> RangeError, handle stays usable */
    napi_value w[3] = {fake_float32_array(in, block - 1), fake_float32_array(out, block), fake_number(pitch)};
    napi_value r = fake_call(proc, "processPacked", 3, w);
    int is_range = 0;
    msg = fake_pending_exception(&is_range);
    EXPECT(r == NULL && msg && is_range, "a Float32Array of the wrong length throws a RangeError");
    /* resize keeps the cursor (ola-processor.js:38-52) */
    napi_value one[1] = {fake_number(channels + 1)};
    fake_call(proc, "resize", 1, one);
    nc = fake_call(proc, "numChannels", 0, NULL);
    tc = fake_call(proc, "timeCursor", 0, NULL);
    EXPECT(nc && nc->num == channels + 1 && tc && tc->num == (double)calls * hop, "resize changes the channel count and keeps timeCursor");
    fake_call(proc, "close", 0, NULL);
    r = fake_call(proc, "processPacked", 3, w);
    msg = fake_pending_exception(NULL);
    EXPECT(r == NULL && msg && strstr(msg, "closed"), "a closed processor throws");
    /* garbage collection path: a second processor finalised by the runtime */
    napi_value proc2 = fake_new(cls, 4, args);
    EXPECT(proc2 != NULL, "second processor");
    fake_collect(proc2);
    f = fopen(argv[7], "wb");
    fwrite(out, sizeof(float), block * calls, f);
    fclose(f);
    printf("%s\n", fails ? "DRIVER FAILED" : "DRIVER OK");
    return fails ? 1 : 0;
}
