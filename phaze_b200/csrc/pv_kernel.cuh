// pv_kernel.cuh — the fused per-call kernel of the phase-vocoder hot path (sm_100a).
//
// One launch == one process() call of the reference for every channel of the
// handle (ola-processor.js:159-171 + phase-vocoder.js:45-72).  Each group of
// T = N/16 threads owns a PAIR of channels; the two channels ride in the two
// halves of packed f32x2 registers (Blackwell FADD2 / FMUL2 / FFMA2), so every
// butterfly instruction works on both channels at once.
//
// Per pair, per call (N = frame, M = N/2, nb = M+1, R = N/hop):
//   P0  gather frame from the history ring + the new block, Hann window      (ola:91-146, pv:55)
//   P1  forward real FFT as an M-point complex FFT of z[m] = x[2m] + j x[2m+1]
//       (radix-8 DIF passes in shared memory)                                (bundle:306-442)
//   P2  real-split -> X[0..M] per channel, |X|^2 in float32                   (pv:82-92)
//   P3  5-point strict-maximum peak bitmap (warp ballots)                     (pv:95-116)
//   P4  region-of-influence shift + rotation                                  (pv:119-173)
//       (source bins above N/2 reproduce what _realTransform4 leaves there)
//   P5  Hermitian C2R pre-pass                                                (bundle:69-76)
//   P6  inverse M-point complex FFT                                           (bundle:102-225)
//   P7  1/N, synthesis window, /R, overlap-add ring, emit hop samples         (pv:65-67, ola:149-157,111-137)
//
// State in HBM: hist[C][N] (input history ring) and acc[C][N] (overlap-add
// ring); nothing is shifted between calls, only `ring_base` moves.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pvb {

struct FrameParams {
    const float *in;       // [C][hop] new input block per channel; nullptr == paused (zeros)
    float *out;            // [C][hop] output block per channel
    float *hist;           // [C][N] input history ring
    float *acc;            // [C][N] overlap-add accumulator ring
    const float *window;   // [N]  Hann window, float32 (pv:8-14)
    const float2 *tw;      // [N]  W_N^j = (cos 2*pi*j/N, -sin 2*pi*j/N)
    int num_channels;
    int hop;
    int overlaps;          // R = N / hop
    int ring_base;         // (calls mod R) * hop : slot that receives the new block / is emitted
    int step_mod_r;        // calls mod R == (timeCursor / hop) mod R
    int src_limit;         // source bins [0, src_limit) can land inside [0, nb)
    float pitch_factor;
    int pf_mant;           // pitch_factor == pf_mant * 2^-pf_shift exactly (float32 mantissa)
    int pf_shift;          // in [1, 62] when the integer form is usable, else 0 (use float64)
    const float *pf_ch;    // per-channel pitch factors [C] (pvb_process_pf) or nullptr: pitch_factor for all
};

// pitch factor of channel c as the exact fraction mant * 2^-shift (shift 0: use float64)
struct PitchFactor {
    float value;
    int mant, shift;
};
__device__ __forceinline__ PitchFactor channel_pitch_factor(const FrameParams &fp, int c) {
    PitchFactor r{fp.pitch_factor, fp.pf_mant, fp.pf_shift};
    if (fp.pf_ch) {
        r.value = __ldg(fp.pf_ch + c);
        const unsigned b = __float_as_uint(r.value);
        const int e = int((b >> 23) & 0xFFu);
        // normal, finite, non-zero: value == (-1)^sign (2^23 + fraction) 2^(e - 150)
        const int m = int((b & 0x7FFFFFu) | 0x800000u);
        r.mant = (b >> 31) ? -m : m;
        r.shift = 150 - e;
        if (e == 0 || e == 255 || r.shift < 1 || r.shift > 62) r.shift = 0;
    }
    return r;
}

// Math.round(p * pitchFactor) (pv:125): round half up of an exact product
__device__ __forceinline__ int round_shifted_peak(int p, int pf_mant, int pf_shift, float pitch_factor) {
    if (pf_shift > 0) {
        const long long v = (long long)pf_mant * p + (1ll << (pf_shift - 1));
        const long long r = v >> pf_shift;
        return r > 0x3fffffff ? 0x3fffffff : (r < -0x3fffffff ? -0x3fffffff : int(r));
    }
    // float64 fallback (pitch factors whose exponent does not fit the integer form, infinities, NaN).
    // Math.round(NaN) is NaN: every comparison of pv:127 / pv:150 is false and the writes go to the
    // property "NaN" of the Array, i.e. nowhere -- the shifted spectrum stays zero.  A value beyond nb
    // has the same effect here (pv:127 `break`).
    const double v = fma(double(p), double(pitch_factor), 0.5);
    if (v != v) return 0x3fffffff;
    const double r = floor(v);
    return r > 1073741823.0 ? 0x3fffffff : (r < -1073741823.0 ? -0x3fffffff : int(r));
}

// ---------------------------------------------------------------------------
// packed two-channel arithmetic (x = channel 0, y = channel 1)
// ---------------------------------------------------------------------------
struct cpx2 {
    float2 re, im;
};

__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }

__device__ __forceinline__ cpx2 cadd(cpx2 a, cpx2 b) { return {add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ cpx2 csub(cpx2 a, cpx2 b) { return {sub2(a.re, b.re), sub2(a.im, b.im)}; }
// multiply both channels by the same scalar complex (wr + j wi)
__device__ __forceinline__ cpx2 cmul_s(cpx2 a, float wr, float wi) {
    const float2 r2 = bc2(wr), i2 = bc2(wi);
    cpx2 o;
    o.re = fma2(a.re, r2, neg2(mul2(a.im, i2)));
    o.im = fma2(a.re, i2, mul2(a.im, r2));
    return o;
}
// times -j (forward) or +j (inverse)
template <bool INV>
__device__ __forceinline__ cpx2 mul_mj(cpx2 a) {
    if (INV) return {neg2(a.im), a.re};
    return {a.im, neg2(a.re)};
}

#define PVB_SQRT1_2 0.70710678118654752440f

// W8^1 (forward: e^{-j pi/4}; inverse: conjugate)
template <bool INV>
__device__ __forceinline__ cpx2 mul_w8_1(cpx2 a) {
    const float2 h = bc2(PVB_SQRT1_2);
    if (INV) return {mul2(sub2(a.re, a.im), h), mul2(add2(a.re, a.im), h)};
    return {mul2(add2(a.re, a.im), h), mul2(sub2(a.im, a.re), h)};
}
// W8^3 (forward: e^{-j 3pi/4})
template <bool INV>
__device__ __forceinline__ cpx2 mul_w8_3(cpx2 a) {
    const float2 h = bc2(PVB_SQRT1_2), nh = bc2(-PVB_SQRT1_2);
    if (INV) return {mul2(add2(a.re, a.im), nh), mul2(sub2(a.re, a.im), h)};
    return {mul2(sub2(a.im, a.re), h), mul2(add2(a.re, a.im), nh)};
}

template <bool INV>
__device__ __forceinline__ void dft4(cpx2 &y0, cpx2 &y1, cpx2 &y2, cpx2 &y3) {
    const cpx2 e0 = cadd(y0, y2), e1 = csub(y0, y2), e2 = cadd(y1, y3);
    const cpx2 e3 = mul_mj<INV>(csub(y1, y3));
    y0 = cadd(e0, e2);
    y1 = cadd(e1, e3);
    y2 = csub(e0, e2);
    y3 = csub(e1, e3);
}

// in-register 8-point DFT, natural order in and out
template <bool INV>
__device__ __forceinline__ void dft8(cpx2 (&x)[8]) {
    cpx2 b0 = cadd(x[0], x[4]), c0 = csub(x[0], x[4]);
    cpx2 b1 = cadd(x[1], x[5]), c1 = mul_w8_1<INV>(csub(x[1], x[5]));
    cpx2 b2 = cadd(x[2], x[6]), c2 = mul_mj<INV>(csub(x[2], x[6]));
    cpx2 b3 = cadd(x[3], x[7]), c3 = mul_w8_3<INV>(csub(x[3], x[7]));
    dft4<INV>(b0, b1, b2, b3);
    dft4<INV>(c0, c1, c2, c3);
    x[0] = b0; x[2] = b1; x[4] = b2; x[6] = b3;
    x[1] = c0; x[3] = c1; x[5] = c2; x[7] = c3;
}

// ---------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------
template <int N>
struct Geo {
    static constexpr int M = N / 2;            // complex FFT length
    static constexpr int NB = M + 1;           // magnitudes.length (pv:40)
    static constexpr int T = M / 8;            // threads per channel pair
    static constexpr int ZSLOTS = M + M / 8 + M / 128 + 1;   // padded float4 slots of the FFT buffer
    static constexpr int NWORDS = (NB + 31) / 32;
    static constexpr int NBP = (NB + 1) & ~1;  // nb rounded up to keep 16-byte alignment
    static constexpr size_t Z_BYTES = size_t(ZSLOTS) * 16;
    static constexpr size_t X_BYTES = size_t(2) * NBP * 8;   // two channels, float2 per bin
    static constexpr size_t MAG_BYTES = size_t(2) * NBP * 4;
    static constexpr size_t PK_BYTES = ((size_t(2) * NWORDS * 4) + 15) & ~size_t(15);
    static constexpr size_t PAIR_BYTES = Z_BYTES + 2 * X_BYTES + MAG_BYTES + PK_BYTES;
    // pairs per CTA: keep the CTA at >= 128 threads
    static constexpr int G = (T >= 128) ? 1 : (128 / T);
    static constexpr int THREADS = G * T;
    static constexpr size_t SMEM_BYTES = size_t(G) * PAIR_BYTES;
    // initial-stage block length of fft.js: 4 when log2(N) is even, else 2 (bundle:28,127-143)
    static constexpr int log2n() { int l = 0; for (int v = N; v > 1; v >>= 1) l++; return l; }
    static constexpr int L0 = (log2n() % 2 == 0) ? 4 : 2;
};

// padded slot of logical complex index i: conflict-free 16-byte accesses for the strides of the
// radix-8 passes (1, 8, 64) and, thanks to the second term, for the digit-reversed reads (stride M/8)
__device__ __forceinline__ int zp(int i) { return i + (i >> 3) + (i >> 7); }

// position (before padding) of natural-order output k after the in-place DIF passes
template <int N>
__device__ __forceinline__ int dif_pos(int k) {
    constexpr int M = N / 2;
    int pos = 0;
    int L = M;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        if (L >= 8) {
            pos += (k & 7) * (L / 8);
            k >>= 3;
            L /= 8;
        }
    }
    if (L == 4) pos += (k & 3);
    if (L == 2) pos += (k & 1);
    return pos;
}

__device__ __forceinline__ cpx2 ldz(const float4 *Z, int i) {
    const float4 v = Z[zp(i)];
    return {make_float2(v.x, v.y), make_float2(v.z, v.w)};
}
__device__ __forceinline__ void stz(float4 *Z, int i, cpx2 c) {
    Z[zp(i)] = make_float4(c.re.x, c.re.y, c.im.x, c.im.y);
}

// one in-place radix-8 DIF pass over blocks of length L (stride S = L/8)
template <int N, int L, bool INV>
__device__ __forceinline__ void pass8(float4 *Z, int t, const float2 *__restrict__ tw) {
    constexpr int S = L / 8;
    const int n = t & (S - 1);
    const int base = (t / S) * L + n;
    cpx2 x[8];
#pragma unroll
    for (int q = 0; q < 8; q++) x[q] = ldz(Z, base + q * S);
    dft8<INV>(x);
    if (S > 1) {
#pragma unroll
        for (int k = 1; k < 8; k++) {
            const float2 w = __ldg(&tw[(N / L) * n * k]);
            x[k] = cmul_s(x[k], w.x, INV ? -w.y : w.y);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) stz(Z, base + k * S, x[k]);
}

// M-point complex FFT, in place, natural order in, dif_pos() order out.
// All threads of the CTA must call it (it contains __syncthreads()).
template <int N, bool INV>
__device__ __forceinline__ void fft_inplace(float4 *Z, int t, const float2 *__restrict__ tw) {
    constexpr int M = N / 2;
    constexpr int T = M / 8;
    if constexpr (M >= 8) { pass8<N, M, INV>(Z, t, tw); __syncthreads(); }
    if constexpr (M >= 64) { pass8<N, M / 8, INV>(Z, t, tw); __syncthreads(); }
    if constexpr (M >= 512) { pass8<N, M / 64, INV>(Z, t, tw); __syncthreads(); }
    if constexpr (M >= 4096) { pass8<N, M / 512, INV>(Z, t, tw); __syncthreads(); }
    constexpr int REM = (M >= 4096) ? M / 4096 : (M >= 512) ? M / 512 : (M >= 64) ? M / 64 : M / 8;
    if constexpr (REM == 4) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int base = 4 * (t + i * T);
            cpx2 y0 = ldz(Z, base), y1 = ldz(Z, base + 1), y2 = ldz(Z, base + 2), y3 = ldz(Z, base + 3);
            dft4<INV>(y0, y1, y2, y3);
            stz(Z, base, y0); stz(Z, base + 1, y1); stz(Z, base + 2, y2); stz(Z, base + 3, y3);
        }
        __syncthreads();
    } else if constexpr (REM == 2) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int base = 2 * (t + i * T);
            const cpx2 a = ldz(Z, base), b = ldz(Z, base + 1);
            stz(Z, base, cadd(a, b));
            stz(Z, base + 1, csub(a, b));
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// peak ownership: which peak's region of influence contains source bin b
// (pv:132-141: regions tile [0, N); boundary between peaks p < q is p + ceil((q-p)/2))
// ---------------------------------------------------------------------------
__device__ __forceinline__ int owner_peak(const uint32_t *pk, int b, int nwords) {
    int w = b >> 5;
    const int bit = b & 31;
    int pl = -1, pr = -1;
    {
        int ww = w;
        uint32_t m;
        if (ww >= nwords) { ww = nwords - 1; m = pk[ww]; }
        else m = pk[ww] & (0xFFFFFFFFu >> (31 - bit));
        while (m == 0 && ww > 0) m = pk[--ww];
        if (m) pl = ww * 32 + 31 - __clz(m);
    }
    if (w < nwords) {
        int ww = w;
        uint32_t m = (bit == 31) ? 0u : (pk[ww] & (0xFFFFFFFEu << bit));
        while (m == 0 && ww < nwords - 1) m = pk[++ww];
        if (m) pr = ww * 32 + __ffs(m) - 1;
    }
    if (pl < 0) return pr;
    if (pr < 0) return pl;
    return (b < pl + ((pr - pl + 1) >> 1)) ? pl : pr;
}

// Value that fft.js leaves in slot `pos` (N/2 < pos < N) of the realTransform output
// (bundle:394-438: every radix-4 stage writes only outputs 0..L/2 of each length-L block,
// so the upper half keeps sub-transform values).  The slot holds bin o of
// DFT_L(xw[r*m + s]) for the (L, r, s, o) found by walking the block tree, and
//   DFT_L(xw[r m + s])[o] = (1/r) * sum_{u<r} W_N^{-s (o + u L)} X[o + u L]
// so it is rebuilt from the valid half-spectrum X[0..N/2] (Hermitian extension above).
template <int N>
__device__ __forceinline__ float2 stale_bin(const float2 *X, int pos, const float2 *__restrict__ tw) {
    int L = N, r = 1, s = 0, o = pos;
    while (L > Geo<N>::L0 && o > (L >> 1)) {
        const int q = L >> 2;
        const int sb = o / q;
        o -= sb * q;
        s += r * sb;
        r <<= 2;
        L = q;
    }
    float ar = 0.f, ai = 0.f;
    for (int u = 0; u < r; u++) {
        const int idx = o + u * L;
        float2 xv;
        if (idx <= N / 2) xv = X[idx];
        else { xv = X[N - idx]; xv.y = -xv.y; }
        const float2 w = __ldg(&tw[(s * idx) & (N - 1)]);   // conj(w) = W_N^{-s idx}
        ar += xv.x * w.x + xv.y * w.y;
        ai += xv.y * w.x - xv.x * w.y;
    }
    const float inv_r = 1.0f / float(r);
    return make_float2(ar * inv_r, ai * inv_r);
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(Geo<N>::THREADS)
pv_process_kernel(const FrameParams p) {
    using G_ = Geo<N>;
    constexpr int M = G_::M, NB = G_::NB, T = G_::T, NWORDS = G_::NWORDS, NBP = G_::NBP;
    constexpr int G = G_::G;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int g = tid / T;          // pair slot inside the CTA
    const int t = tid - g * T;      // thread inside the pair group
    unsigned char *mine = smem_raw + size_t(g) * G_::PAIR_BYTES;
    float4 *Z = reinterpret_cast<float4 *>(mine);
    float2 *X = reinterpret_cast<float2 *>(mine + G_::Z_BYTES);                      // [2][NBP]
    float2 *Y = reinterpret_cast<float2 *>(mine + G_::Z_BYTES + G_::X_BYTES);        // [2][NBP]
    float *mag = reinterpret_cast<float *>(mine + G_::Z_BYTES + 2 * G_::X_BYTES);    // [2][NBP]
    uint32_t *pk = reinterpret_cast<uint32_t *>(mine + G_::Z_BYTES + 2 * G_::X_BYTES + G_::MAG_BYTES);

    const int pair = blockIdx.x * G + g;
    const int c0 = 2 * pair, c1 = c0 + 1;
    const bool has0 = c0 < p.num_channels, has1 = c1 < p.num_channels;
    const int hop = p.hop;
    const int rb = p.ring_base;
    const float2 *__restrict__ tw = p.tw;

    // ---- P0: frame gather + analysis window -------------------------------------------
    {
        const int keep = N - hop;   // samples that come from the history ring
        for (int m = t; m < M; m += T) {
            const int n = 2 * m;
            float2 v0 = make_float2(0.f, 0.f), v1 = v0;
            if (n < keep) {
                const int r = (n + rb + hop) & (N - 1);
                if (has0) v0 = *reinterpret_cast<const float2 *>(p.hist + size_t(c0) * N + r);
                if (has1) v1 = *reinterpret_cast<const float2 *>(p.hist + size_t(c1) * N + r);
            } else {
                const int i = n - keep;
                if (p.in) {
                    if (has0) v0 = __ldg(reinterpret_cast<const float2 *>(p.in + size_t(c0) * hop + i));
                    if (has1) v1 = __ldg(reinterpret_cast<const float2 *>(p.in + size_t(c1) * hop + i));
                }
                if (has0) *reinterpret_cast<float2 *>(p.hist + size_t(c0) * N + rb + i) = v0;
                if (has1) *reinterpret_cast<float2 *>(p.hist + size_t(c1) * N + rb + i) = v1;
            }
            const float2 w = __ldg(reinterpret_cast<const float2 *>(p.window + n));
            Z[zp(m)] = make_float4(v0.x * w.x, v1.x * w.x, v0.y * w.y, v1.y * w.y);
        }
    }
    __syncthreads();

    // ---- P1: forward FFT ---------------------------------------------------------------
    fft_inplace<N, false>(Z, t, tw);

    // ---- P2: real split, |X|^2, clear the shifted spectrum -------------------------------
    // X holds 2*X_true (the factor is folded into the final scale).
    for (int k = t; k <= M / 2; k += T) {
        const cpx2 a = ldz(Z, dif_pos<N>(k));
        const cpx2 b = ldz(Z, dif_pos<N>((M - k) & (M - 1)));
        const float2 e_r = add2(a.re, b.re), e_i = sub2(a.im, b.im);
        const float2 o_r = add2(a.im, b.im), o_i = sub2(b.re, a.re);
        const float2 w = __ldg(&tw[k]);
        const cpx2 tt = cmul_s(cpx2{o_r, o_i}, w.x, w.y);
        const float2 xr = add2(e_r, tt.re), xi = add2(e_i, tt.im);          // X[k]
        const float2 yr = sub2(e_r, tt.re), yi = sub2(tt.im, e_i);          // X[M-k]
        X[k] = make_float2(xr.x, xi.x);
        X[NBP + k] = make_float2(xr.y, xi.y);
        X[M - k] = make_float2(yr.x, yi.x);
        X[NBP + M - k] = make_float2(yr.y, yi.y);
        mag[k] = fmaf(xr.x, xr.x, xi.x * xi.x);
        mag[NBP + k] = fmaf(xr.y, xr.y, xi.y * xi.y);
        mag[M - k] = fmaf(yr.x, yr.x, yi.x * yi.x);
        mag[NBP + M - k] = fmaf(yr.y, yr.y, yi.y * yi.y);
    }
    for (int i = t; i < 2 * NBP; i += T) Y[i] = make_float2(0.f, 0.f);   // pv:121
    __syncthreads();

    // ---- P3: peaks -> bitmap ------------------------------------------------------------
    {
        constexpr int NWARPS = G_::THREADS / 32;
        const int warp = tid >> 5, lane = tid & 31;
        for (int item = warp; item < G * 2 * NWORDS; item += NWARPS) {
            const int w = item % NWORDS;
            const int gc = item / NWORDS;            // pair slot * 2 + channel
            unsigned char *base = smem_raw + size_t(gc >> 1) * G_::PAIR_BYTES;
            const float *mg = reinterpret_cast<const float *>(base + G_::Z_BYTES + 2 * G_::X_BYTES) +
                              (gc & 1) * NBP;
            uint32_t *pw = reinterpret_cast<uint32_t *>(base + G_::Z_BYTES + 2 * G_::X_BYTES +
                                                        G_::MAG_BYTES) + (gc & 1) * NWORDS;
            const int i = w * 32 + lane;
            bool is_peak = false;
            if (i >= 2 && i < NB - 2) {
                const float v = mg[i];
                is_peak = !(mg[i - 1] >= v) && !(mg[i - 2] >= v) && !(mg[i + 1] >= v) && !(mg[i + 2] >= v);
            }
            const uint32_t bits = __ballot_sync(0xFFFFFFFFu, is_peak);
            if (lane == 0) pw[w] = bits;
        }
    }
    __syncthreads();

    // ---- P4: shift every region of influence to its new place -----------------------------
    {
        // per-channel pitch factors (pvb_process_pf): every source bin may land, and colliding regions
        // are possible whenever some channel contracts (adding to a zeroed bin is exact, so the atomic
        // path is taken for all channels then)
        const int limit = p.pf_ch ? N : p.src_limit;
        const bool contract = p.pf_ch ? true : p.pitch_factor < 1.0f;     // only then two sources can hit one bin
        const PitchFactor pfa = channel_pitch_factor(p, has0 ? c0 : 0), pfb = channel_pitch_factor(p, has1 ? c1 : (has0 ? c0 : 0));
        const int rmask = p.overlaps - 1;
        const int rstride = N / p.overlaps;
        for (int idx = t; idx < 2 * limit; idx += T) {
            const int ch = idx >= limit;
            const int b = idx - ch * limit;
            const uint32_t *pkc = pk + ch * NWORDS;
            const int pi = owner_peak(pkc, b, NWORDS);
            if (pi < 0) continue;                                        // no peaks at all
            const PitchFactor &pfc = ch ? pfb : pfa;
            const int ps = round_shifted_peak(pi, pfc.mant, pfc.shift, pfc.value);     // Math.round, pv:125
            if (ps > NB) continue;                                       // pv:127
            const int delta = ps - pi;
            const int d = b + delta;
            if (d < 0 || d >= NB) continue;                              // pv:150 and the negative-index no-op
            const float2 *Xc = X + ch * NBP;
            const float2 v = (b <= M) ? Xc[b] : stale_bin<N>(Xc, b, tw);
            // exp(j*omega*t) with omega*t = 2*pi*delta*(calls)/R  (pv:155-157, integer reduced)
            const int ri = (delta * p.step_mod_r) & rmask;
            const float2 w = __ldg(&tw[ri * rstride]);                   // (cos, -sin)
            const float yr = v.x * w.x + v.y * w.y;
            const float yi = v.y * w.x - v.x * w.y;
            float2 *dst = Y + ch * NBP + d;
            if (contract) {
                atomicAdd(&dst->x, yr);
                atomicAdd(&dst->y, yi);
            } else {
                *dst = make_float2(yr, yi);
            }
        }
    }
    __syncthreads();

    // ---- P5: Hermitian C2R pre-pass: Y[0..M] -> Z'[0..M) (natural order) -------------------
    for (int k = t; k <= M / 2; k += T) {
        float2 a0 = Y[k], a1 = Y[NBP + k], b0 = Y[M - k], b1 = Y[NBP + M - k];
        if (k == 0) { a0.y = 0.f; a1.y = 0.f; b0.y = 0.f; b1.y = 0.f; }   // only Re of DC / Nyquist reaches the output
        const float2 ar = make_float2(a0.x, a1.x), ai = make_float2(a0.y, a1.y);
        const float2 br = make_float2(b0.x, b1.x), bi = make_float2(b0.y, b1.y);
        const float2 e_r = add2(ar, br), e_i = sub2(ai, bi);
        const float2 d_r = sub2(ar, br), d_i = add2(ai, bi);
        const float2 w = __ldg(&tw[k]);
        const cpx2 pp = cmul_s(cpx2{d_r, d_i}, w.x, -w.y);                // D * conj(W_N^k)
        stz(Z, k, cpx2{sub2(e_r, pp.im), add2(e_i, pp.re)});
        if (k != 0) stz(Z, M - k, cpx2{add2(e_r, pp.im), sub2(pp.re, e_i)});
    }
    __syncthreads();

    // ---- P6: inverse FFT --------------------------------------------------------------------
    fft_inplace<N, true>(Z, t, tw);

    // ---- P7: scale, synthesis window, overlap-add ring, emit ----------------------------------
    {
        const float scale = 1.0f / float(2 * N);     // 1/N of inverseTransform and the two folded 1/2
        const float inv_r = 1.0f / float(p.overlaps);
        for (int q = t; q < N / 4; q += T) {
            const float4 za = Z[zp(dif_pos<N>(2 * q))];
            const float4 zb = Z[zp(dif_pos<N>(2 * q + 1))];
            const int k = 4 * q;
            const float4 w = __ldg(reinterpret_cast<const float4 *>(p.window + k));
            const bool head = k < hop;
            const bool tail = k >= N - hop;
            const int ring = (k + rb) & (N - 1);
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                if (!(ch ? has1 : has0)) continue;
                const int c = ch ? c1 : c0;
                float4 y;
                y.x = (ch ? za.y : za.x) * scale;    // fromComplexArray -> float32 (pv:65)
                y.y = (ch ? za.w : za.z) * scale;
                y.z = (ch ? zb.y : zb.x) * scale;
                y.w = (ch ? zb.w : zb.z) * scale;
                y.x = (y.x * w.x) * inv_r;           // applyHannWindow (pv:67), / nbOverlaps (ola:153)
                y.y = (y.y * w.y) * inv_r;
                y.z = (y.z * w.z) * inv_r;
                y.w = (y.w * w.w) * inv_r;
                float4 *ap = reinterpret_cast<float4 *>(p.acc + size_t(c) * N + ring);
                if (!tail) {
                    const float4 a = *ap;
                    y.x += a.x; y.y += a.y; y.z += a.z; y.w += a.w;
                }
                if (head) *reinterpret_cast<float4 *>(p.out + size_t(c) * hop + k) = y;
                else *ap = y;
            }
        }
    }
}

}  // namespace pvb
