/*
 * phaze_oracle.c — CPU oracle for the phaze process() path.  TEST INFRASTRUCTURE
 * ONLY (see phaze_oracle.h): the product never calls into this file.
 *
 * Restates, function by function, what the reference computes.  "bundle" below
 * is /root/reference/www/phase-vocoder.js (the only copy of fft.js 4.0.3 in the
 * tree), "ola" is /root/reference/src/ola-processor.js and "pv" is
 * /root/reference/src/phase-vocoder.js.
 *
 * Precision map: JS Array / Number -> double, Float32Array ->
 * float (one rounding at every store), Int32Array -> int32_t.
 */
#include "phaze_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ===================================================================== */
/* fft.js 4.0.3                                                           */
/* ===================================================================== */

typedef struct {
    int n;          /* this.size   */
    int n2;         /* this._csize */
    int width;      /* this._width */
    double *tw;     /* this.table : tw[2m], tw[2m+1] = cos, -sin of 2*pi*m/n */
    int32_t *rev;   /* this._bitrev, 1 << width entries                      */
} fftjs;

/* JS `a << b` on int32: shift count is taken modulo 32 */
static int32_t js_shl(int32_t a, int b) { return (int32_t)((uint32_t)a << (b & 31)); }

/* function FFT(size) — bundle:4-43 */
static int fftjs_init(fftjs *f, int size)
{
    memset(f, 0, sizeof(*f));
    if (size <= 1 || (size & (size - 1)) != 0) return -1;     /* bundle:6-7 */
    f->n = size;
    f->n2 = size << 1;
    f->tw = (double *)malloc(sizeof(double) * (size_t)f->n2);
    for (int i = 0; i < f->n2; i += 2) {                      /* bundle:13-17 */
        const double angle = M_PI * i / size;
        f->tw[i] = cos(angle);
        f->tw[i + 1] = -sin(angle);
    }
    int power = 0;                                            /* bundle:21-23 */
    for (int t = 1; size > t; t <<= 1) power++;
    f->width = (power % 2 == 0) ? power - 1 : power;          /* bundle:28 */
    const int cnt = 1 << f->width;
    f->rev = (int32_t *)malloc(sizeof(int32_t) * (size_t)cnt);
    for (int j = 0; j < cnt; j++) {                           /* bundle:32-38 */
        int32_t r = 0;
        for (int shift = 0; shift < f->width; shift += 2) {
            const int back = f->width - shift - 2;            /* may be -1: JS masks it */
            r |= js_shl((j >> shift) & 3, back);
        }
        f->rev[j] = r;
    }
    return 0;
}

static void fftjs_free(fftjs *f)
{
    free(f->tw);
    free(f->rev);
    f->tw = NULL;
    f->rev = NULL;
}

/* completeSpectrum — bundle:69-76 */
static void fftjs_complete_spectrum(const fftjs *f, double *s)
{
    const int size = f->n2, half = size >> 1;
    for (int i = 2; i < half; i += 2) {
        s[size - i] = s[i];
        s[size - i + 1] = -s[i + 1];
    }
}

/* _singleTransform2 — bundle:230-249 */
static void cplx_first2(double *out, const double *in, int o, int off, int step)
{
    const double er = in[off], ei = in[off + 1];
    const double qr = in[off + step], qi = in[off + step + 1];
    out[o] = er + qr;
    out[o + 1] = ei + qi;
    out[o + 2] = er - qr;
    out[o + 3] = ei - qi;
}

/* _singleTransform4 — bundle:254-303 */
static void cplx_first4(double *out, const double *in, int o, int off, int step, double inv)
{
    const double ar = in[off], ai = in[off + 1];
    const double br = in[off + step], bi = in[off + step + 1];
    const double cr = in[off + 2 * step], ci = in[off + 2 * step + 1];
    const double dr = in[off + 3 * step], di = in[off + 3 * step + 1];
    const double s0r = ar + cr, s0i = ai + ci;
    const double s1r = ar - cr, s1i = ai - ci;
    const double s2r = br + dr, s2i = bi + di;
    const double s3r = inv * (br - dr), s3i = inv * (bi - di);
    out[o] = s0r + s2r;
    out[o + 1] = s0i + s2i;
    out[o + 2] = s1r + s3i;
    out[o + 3] = s1i - s3r;
    out[o + 4] = s0r - s2r;
    out[o + 5] = s0i - s2i;
    out[o + 6] = s1r - s3i;
    out[o + 7] = s1i + s3r;
}

/* _transform4 — bundle:120-225 (inverse != 0 conjugates the twiddles) */
static void fftjs_transform4(const fftjs *f, double *out, const double *in, int inverse)
{
    const int size = f->n2;
    int step = 1 << f->width;
    int len = (size / step) << 1;
    const double inv = inverse ? -1.0 : 1.0;
    const double *tw = f->tw;

    if (len == 4) {
        for (int o = 0, t = 0; o < size; o += len, t++) cplx_first2(out, in, o, f->rev[t], step);
    } else {
        for (int o = 0, t = 0; o < size; o += len, t++) cplx_first4(out, in, o, f->rev[t], step, inv);
    }

    for (step >>= 2; step >= 2; step >>= 2) {
        len = (size / step) << 1;
        const int q = len >> 2;
        for (int base = 0; base < size; base += len) {
            const int limit = base + q;
            for (int i = base, k = 0; i < limit; i += 2, k += step) {
                const int pa = i, pb = pa + q, pc = pb + q, pd = pc + q;
                const double ar = out[pa], ai = out[pa + 1];
                const double br = out[pb], bi = out[pb + 1];
                const double cr = out[pc], ci = out[pc + 1];
                const double dr = out[pd], di = out[pd + 1];

                const double wbr = tw[k], wbi = inv * tw[k + 1];
                const double mbr = br * wbr - bi * wbi, mbi = br * wbi + bi * wbr;
                const double wcr = tw[2 * k], wci = inv * tw[2 * k + 1];
                const double mcr = cr * wcr - ci * wci, mci = cr * wci + ci * wcr;
                const double wdr = tw[3 * k], wdi = inv * tw[3 * k + 1];
                const double mdr = dr * wdr - di * wdi, mdi = dr * wdi + di * wdr;

                const double s0r = ar + mcr, s0i = ai + mci;
                const double s1r = ar - mcr, s1i = ai - mci;
                const double s2r = mbr + mdr, s2i = mbi + mdi;
                const double s3r = inv * (mbr - mdr), s3i = inv * (mbi - mdi);

                out[pa] = s0r + s2r;
                out[pa + 1] = s0i + s2i;
                out[pb] = s1r + s3i;
                out[pb + 1] = s1i - s3r;
                out[pc] = s0r - s2r;
                out[pc + 1] = s0i - s2i;
                out[pd] = s1r - s3i;
                out[pd + 1] = s1i + s3r;
            }
        }
    }
}

/* _singleRealTransform2 — bundle:447-463 */
static void real_first2(double *out, const float *in, int o, int off, int step)
{
    const double e = in[off], q = in[off + step];
    out[o] = e + q;
    out[o + 1] = 0;
    out[o + 2] = e - q;
    out[o + 3] = 0;
}

/* _singleRealTransform4 — bundle:468-508 (forward only: this._inv == 0) */
static void real_first4(double *out, const float *in, int o, int off, int step)
{
    const double a = in[off], b = in[off + step];
    const double c = in[off + 2 * step], d = in[off + 3 * step];
    const double s0 = a + c, s1 = a - c, s2 = b + d, s3 = 1.0 * (b - d);
    out[o] = s0 + s2;
    out[o + 1] = 0;
    out[o + 2] = s1;
    out[o + 3] = -s3;
    out[o + 4] = s0 - s2;
    out[o + 5] = 0;
    out[o + 6] = s1;
    out[o + 7] = s3;
}

/* _realTransform4 — bundle:306-442.  In place on `out` after the first pass.
 * Each later stage writes only outputs 0..L/2 of every length-L block (A, B, the
 * middle point C at i==0, and the mirrored SA/SB), so slots above N/2 keep
 * sub-transform values of earlier stages. */
static void fftjs_real_transform4(const fftjs *f, double *out, const float *in)
{
    const int size = f->n2;
    int step = 1 << f->width;
    int len = (size / step) << 1;
    const double inv = 1.0;                 /* realTransform sets _inv = 0 (bundle:96) */
    const double *tw = f->tw;

    if (len == 4) {
        for (int o = 0, t = 0; o < size; o += len, t++)
            real_first2(out, in, o, (int)((uint32_t)f->rev[t] >> 1), step >> 1);
    } else {
        for (int o = 0, t = 0; o < size; o += len, t++)
            real_first4(out, in, o, (int)((uint32_t)f->rev[t] >> 1), step >> 1);
    }

    for (step >>= 2; step >= 2; step >>= 2) {
        len = (size / step) << 1;
        const int h = len >> 1, q = h >> 1, hq = q >> 1;
        for (int base = 0; base < size; base += len) {
            for (int i = 0, k = 0; i <= hq; i += 2, k += step) {
                const int pa = base + i, pb = pa + q, pc = pb + q, pd = pc + q;
                const double ar = out[pa], ai = out[pa + 1];
                const double br = out[pb], bi = out[pb + 1];
                const double cr = out[pc], ci = out[pc + 1];
                const double dr = out[pd], di = out[pd + 1];

                const double wbr = tw[k], wbi = inv * tw[k + 1];
                const double mbr = br * wbr - bi * wbi, mbi = br * wbi + bi * wbr;
                const double wcr = tw[2 * k], wci = inv * tw[2 * k + 1];
                const double mcr = cr * wcr - ci * wci, mci = cr * wci + ci * wcr;
                const double wdr = tw[3 * k], wdi = inv * tw[3 * k + 1];
                const double mdr = dr * wdr - di * wdi, mdi = dr * wdi + di * wdr;

                const double s0r = ar + mcr, s0i = ai + mci;
                const double s1r = ar - mcr, s1i = ai - mci;
                const double s2r = mbr + mdr, s2i = mbi + mdi;
                const double s3r = inv * (mbr - mdr), s3i = inv * (mbi - mdi);

                out[pa] = s0r + s2r;
                out[pa + 1] = s0i + s2i;
                out[pb] = s1r + s3i;
                out[pb + 1] = s1i - s3r;

                if (i == 0) {               /* middle point, bundle:400-406 */
                    out[pc] = s0r - s2r;
                    out[pc + 1] = s0i - s2i;
                    continue;
                }
                if (i == hq) continue;      /* bundle:409-410 */

                /* mirrored outputs, bundle:417-438 */
                const double u0r = s1r, u0i = -s1i;
                const double u1r = s0r, u1i = -s0i;
                const double u2r = -inv * s3i, u2i = -inv * s3r;
                const double u3r = -inv * s2i, u3i = -inv * s2r;
                const int sa = base + q - i, sb = base + h - i;
                out[sa] = u0r + u2r;
                out[sa + 1] = u0i + u2i;
                out[sb] = u1r + u3i;
                out[sb + 1] = u1i - u3r;
            }
        }
    }
}

/* inverseTransform — bundle:102-114 */
static void fftjs_inverse_transform(const fftjs *f, double *out, const double *in)
{
    fftjs_transform4(f, out, in, 1);
    for (int i = 0; i < f->n2; i++) out[i] /= f->n;
}

/* ===================================================================== */
/* PhaseVocoderProcessor per-frame body (pv:52-67, 75-173)                */
/* ===================================================================== */

typedef struct {
    int n;              /* fftSize == blockSize */
    int nb;             /* magnitudes.length == n/2+1 */
    fftjs fft;
    float *hann;        /* pv:8-14 */
    double *freq;       /* freqComplexBuffer        f64[2n] */
    double *shifted;    /* freqComplexBufferShifted f64[2n] */
    double *timec;      /* timeComplexBuffer        f64[2n] */
    float *mag;         /* magnitudes               f32[nb] */
    int32_t *peaks;     /* peakIndexes              i32[nb] */
    int nb_peaks;
    int max_src_bin;    /* debug: highest bin read by shift_peaks */
} pv_core;

static int pv_core_init(pv_core *c, int n)
{
    memset(c, 0, sizeof(*c));
    if (fftjs_init(&c->fft, n) != 0) return -1;
    c->n = n;
    c->nb = n / 2 + 1;
    c->hann = (float *)malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; i++)                                   /* pv:10-12 */
        c->hann[i] = (float)(0.5 * (1 - cos(2 * M_PI * i / n)));
    c->freq = (double *)calloc((size_t)2 * n, sizeof(double));
    c->shifted = (double *)calloc((size_t)2 * n, sizeof(double));
    c->timec = (double *)calloc((size_t)2 * n, sizeof(double));
    c->mag = (float *)calloc((size_t)c->nb, sizeof(float));
    c->peaks = (int32_t *)calloc((size_t)c->nb, sizeof(int32_t));
    c->max_src_bin = -1;
    return 0;
}

static void pv_core_free(pv_core *c)
{
    fftjs_free(&c->fft);
    free(c->hann); free(c->freq); free(c->shifted); free(c->timec);
    free(c->mag); free(c->peaks);
}

/* applyHannWindow — pv:75-79 (f32 * f32, stored f32) */
static void pv_window(const pv_core *c, float *x)
{
    for (int i = 0; i < c->n; i++) x[i] = (float)((double)x[i] * (double)c->hann[i]);
}

/* computeMagnitudes — pv:82-92 (f64 arithmetic, f32 store) */
static void pv_magnitudes(pv_core *c)
{
    for (int i = 0, j = 0; i < c->nb; i++, j += 2) {
        const double re = c->freq[j], im = c->freq[j + 1];
        c->mag[i] = (float)(re * re + im * im);
    }
}

/* findPeaks — pv:95-116 */
static void pv_find_peaks(pv_core *c)
{
    const float *m = c->mag;
    c->nb_peaks = 0;
    int i = 2;
    const int end = c->nb - 2;
    while (i < end) {
        const float v = m[i];
        if (m[i - 1] >= v || m[i - 2] >= v) { i++; continue; }
        if (m[i + 1] >= v || m[i + 2] >= v) { i++; continue; }
        c->peaks[c->nb_peaks++] = i;
        i += 2;
    }
}

/* shiftPeaks — pv:119-173 */
static void pv_shift_peaks(pv_core *c, double pitch_factor, double time_cursor)
{
    const int n = c->n, nb = c->nb;
    memset(c->shifted, 0, sizeof(double) * (size_t)2 * n);        /* pv:121 */

    for (int i = 0; i < c->nb_peaks; i++) {
        const int p = c->peaks[i];
        /* Math.round: nearest, ties toward +inf; p*pf is exact in f64 here */
        const double psd = floor((double)p * pitch_factor + 0.5);
        /* NaN (pitchFactor NaN): pv:127 and pv:150 compare false and pv:169-170 write the property
         * "NaN" of the Array, i.e. no element: the region contributes nothing */
        if (psd != psd) continue;
        if (psd > (double)nb) break;                              /* pv:127 */
        /* far below zero every shifted bin is negative: named properties again (see bs < 0 below) */
        if (psd < -(double)(2 * n)) continue;
        const int ps = (int)psd;

        int start = 0, end = n;                                   /* pv:132-133 */
        if (i > 0) {
            const int before = c->peaks[i - 1];
            start = p - (int)floor((p - before) / 2.0);
        }
        if (i < c->nb_peaks - 1) {
            const int after = c->peaks[i + 1];
            end = p + (int)ceil((after - p) / 2.0);
        }

        for (int j = start - p; j < end - p; j++) {               /* pv:146 */
            const int b = p + j;
            const int bs = ps + j;
            if (bs >= nb) break;                                  /* pv:150 */

            const double omega = 2 * M_PI * (bs - b) / n;         /* pv:155 */
            const double rot_r = cos(omega * time_cursor);
            const double rot_i = sin(omega * time_cursor);
            const double vr = c->freq[2 * b], vi = c->freq[2 * b + 1];
            if (b > c->max_src_bin) c->max_src_bin = b;
            const double yr = vr * rot_r - vi * rot_i;
            const double yi = vr * rot_i + vi * rot_r;
            /* a negative index on a JS Array sets a named property and leaves
             * every element untouched (pv:169-170 with bs < 0) */
            if (bs < 0) continue;
            c->shifted[2 * bs] += yr;
            c->shifted[2 * bs + 1] += yi;
        }
    }
}

/* the channel body of processOLA — pv:52-67.  `frame` is modified in place by
 * the analysis window exactly like the JS `input`. */
static void pv_core_frame(pv_core *c, float *frame, float *out, double pitch_factor,
                          double time_cursor)
{
    pv_window(c, frame);                                          /* pv:55 */
    fftjs_real_transform4(&c->fft, c->freq, frame);               /* pv:57 */
    pv_magnitudes(c);                                             /* pv:59 */
    pv_find_peaks(c);                                             /* pv:60 */
    pv_shift_peaks(c, pitch_factor, time_cursor);                 /* pv:61 */
    fftjs_complete_spectrum(&c->fft, c->shifted);                 /* pv:63 */
    fftjs_inverse_transform(&c->fft, c->timec, c->shifted);       /* pv:64 */
    for (int i = 0; i < 2 * c->n; i += 2)                         /* pv:65, bundle:46-51 */
        out[i >> 1] = (float)c->timec[i];
    pv_window(c, out);                                            /* pv:67 */
}

/* ===================================================================== */
/* OLAProcessor (ola:7-171) around it                                     */
/* ===================================================================== */

struct pvo_processor {
    int n, hop, overlaps, channels;
    double time_cursor;                 /* pv:31 */
    pv_core core;
    float *in_hist;     /* inputBuffers          [C][n+hop]  (ola:59) */
    float *in_send;     /* inputBuffersToSend    [C][n]      (ola:69) */
    float *out_acc;     /* outputBuffers         [C][n]      (ola:77) */
    float *out_ret;     /* outputBuffersToRetrieve [C][n]    (ola:85) */
};

static int alloc_channels(pvo_processor *p, int channels)
{
    free(p->in_hist); free(p->in_send); free(p->out_acc); free(p->out_ret);
    p->channels = channels;
    const size_t c = (size_t)(channels > 0 ? channels : 1);
    p->in_hist = (float *)calloc(c * (size_t)(p->n + p->hop), sizeof(float));
    p->in_send = (float *)calloc(c * (size_t)p->n, sizeof(float));
    p->out_acc = (float *)calloc(c * (size_t)p->n, sizeof(float));
    p->out_ret = (float *)calloc(c * (size_t)p->n, sizeof(float));
    return (p->in_hist && p->in_send && p->out_acc && p->out_ret) ? 0 : -1;
}

pvo_processor *pvo_create(int frame_size, int hop_size, int num_channels)
{
    if (hop_size <= 0 || num_channels < 0 || frame_size % hop_size != 0) return NULL;
    pvo_processor *p = (pvo_processor *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    if (pv_core_init(&p->core, frame_size) != 0) { free(p); return NULL; }
    p->n = frame_size;
    p->hop = hop_size;
    p->overlaps = frame_size / hop_size;                          /* ola:17 */
    p->time_cursor = 0;
    if (alloc_channels(p, num_channels) != 0) { pvo_destroy(p); return NULL; }
    return p;
}

void pvo_destroy(pvo_processor *p)
{
    if (!p) return;
    pv_core_free(&p->core);
    free(p->in_hist); free(p->in_send); free(p->out_acc); free(p->out_ret);
    free(p);
}

int pvo_resize(pvo_processor *p, int num_channels)
{
    if (!p || num_channels < 0) return -1;
    return alloc_channels(p, num_channels);                       /* ola:54-88 */
}

double pvo_time_cursor(const pvo_processor *p) { return p->time_cursor; }
void pvo_set_time_cursor(pvo_processor *p, double t) { p->time_cursor = t; }
int pvo_max_source_bin(const pvo_processor *p) { return p->core.max_src_bin; }

int pvo_process(pvo_processor *p, const float *in, float *out, float pitch_factor)
{
    const int n = p->n, hop = p->hop, C = p->channels;
    const double pf = (double)pitch_factor;   /* AudioParam value is a float32 (pv:47) */

    /* readInputs — ola:91-108 */
    for (int c = 0; c < C; c++) {
        float *h = p->in_hist + (size_t)c * (n + hop);
        if (in == NULL) memset(h + n, 0, sizeof(float) * (size_t)hop);      /* ola:96 */
        else memcpy(h + n, in + (size_t)c * hop, sizeof(float) * (size_t)hop); /* ola:105 */
    }
    /* shiftInputBuffers — ola:121-127 */
    for (int c = 0; c < C; c++) {
        float *h = p->in_hist + (size_t)c * (n + hop);
        memmove(h, h + hop, sizeof(float) * (size_t)n);
    }
    /* prepareInputBuffersToSend — ola:140-146 */
    for (int c = 0; c < C; c++)
        memcpy(p->in_send + (size_t)c * n, p->in_hist + (size_t)c * (n + hop),
               sizeof(float) * (size_t)n);

    /* processOLA — pv:45-72 */
    for (int c = 0; c < C; c++)
        pv_core_frame(&p->core, p->in_send + (size_t)c * n, p->out_ret + (size_t)c * n, pf,
                      p->time_cursor);
    p->time_cursor += hop;                                        /* pv:71 */

    /* handleOutputBuffersToRetrieve — ola:149-157 */
    for (int c = 0; c < C; c++) {
        float *acc = p->out_acc + (size_t)c * n;
        const float *ret = p->out_ret + (size_t)c * n;
        for (int k = 0; k < n; k++)
            acc[k] = (float)((double)acc[k] + (double)ret[k] / (double)p->overlaps);
    }
    /* writeOutputs — ola:111-118 */
    for (int c = 0; c < C; c++)
        memcpy(out + (size_t)c * hop, p->out_acc + (size_t)c * n, sizeof(float) * (size_t)hop);
    /* shiftOutputBuffers — ola:130-137 */
    for (int c = 0; c < C; c++) {
        float *acc = p->out_acc + (size_t)c * n;
        memmove(acc, acc + hop, sizeof(float) * (size_t)(n - hop));
        memset(acc + (n - hop), 0, sizeof(float) * (size_t)hop);
    }
    return 1;                                                     /* ola:170 */
}

/* ===================================================================== */
/* white-box hooks                                                        */
/* ===================================================================== */

int pvo_fft_real_transform(int n, const float *data, double *out)
{
    fftjs f;
    if (fftjs_init(&f, n) != 0) return -1;
    fftjs_real_transform4(&f, out, data);
    fftjs_free(&f);
    return 0;
}

int pvo_fft_inverse_transform(int n, const double *data, double *out)
{
    fftjs f;
    if (fftjs_init(&f, n) != 0) return -1;
    fftjs_inverse_transform(&f, out, data);
    fftjs_free(&f);
    return 0;
}

int pvo_fft_complete_spectrum(int n, double *spectrum)
{
    fftjs f;
    if (fftjs_init(&f, n) != 0) return -1;
    fftjs_complete_spectrum(&f, spectrum);
    fftjs_free(&f);
    return 0;
}

int pvo_frame(int n, const float *frame_in, float pitch_factor, double time_cursor,
              float *frame_out, double *spectrum, float *magnitudes,
              int32_t *peaks, int32_t *nb_peaks, double *shifted)
{
    pv_core c;
    if (pv_core_init(&c, n) != 0) return -1;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(tmp, frame_in, sizeof(float) * (size_t)n);
    pv_core_frame(&c, tmp, frame_out, (double)pitch_factor, time_cursor);
    if (spectrum) memcpy(spectrum, c.freq, sizeof(double) * (size_t)2 * n);
    if (magnitudes) memcpy(magnitudes, c.mag, sizeof(float) * (size_t)c.nb);
    if (peaks) memcpy(peaks, c.peaks, sizeof(int32_t) * (size_t)c.nb_peaks);
    if (nb_peaks) *nb_peaks = c.nb_peaks;
    if (shifted) memcpy(shifted, c.shifted, sizeof(double) * (size_t)2 * n);
    free(tmp);
    pv_core_free(&c);
    return 0;
}
