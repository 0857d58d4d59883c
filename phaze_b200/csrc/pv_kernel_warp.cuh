// pv_kernel_warp.cuh — warp-synchronous fused kernel for frame size 1024 (sm_100a).
//
// Same arithmetic as pv_kernel.cuh (one launch == one process() call, reference lines
// cited there), re-laid-out so that ONE WARP owns one channel pair end to end:
//
//   * no __syncthreads(): every exchange is shared memory + __syncwarp(), so warps drift
//     apart and the scheduler overlaps FFT math of one pair with the integer-heavy peak /
//     shift phases of another;
//   * the frame goes global memory -> registers -> (2 exchanges) -> registers for both FFTs:
//     first-pass inputs are read straight from the history ring (coalesced 8-byte lanes),
//     last-pass outputs of the inverse go straight to the overlap-add ring;
//   * the last forward pass gives every lane both Z[k] and Z[M-k], so the real-split (and its
//     mirror, the Hermitian C2R pre-pass) happens in registers;
//   * exchanges use an XOR swizzle (slot = 64a + 8b + (c ^ b)): all three access patterns of
//     the radix-8 passes are bank-conflict free without padding;
//   * peak picking walks 16-bin runs per lane (vector loads), regions get 32-bit descriptors
//     indexed by a prefix-popcount of a region-start bitmap, Math.round(p * pitchFactor) is
//     exact integer arithmetic on the float32 mantissa (no FP64, no XU pipe).
//
// Valid for pitch factors >= 2/3 (right halves of consecutive regions, and left halves, are
// then pairwise disjoint after the shift, so two ordered sub-steps replace atomics); the host
// routes other factors and other frame sizes to the generic kernel in pv_kernel.cuh.
#pragma once

#include "pv_kernel.cuh"

namespace pvb {

struct WarpGeo {
    static constexpr int N = 1024, M = 512, NB = 513;
    static constexpr int BUF_BYTES = 8256;                 // >= 513 * 16
    static constexpr int MAG_FLOATS = 656;                 // per channel, padded layout (see mag_idx)
    static constexpr int MAXPK = 176;                      // > 509 / 3 peaks
    static constexpr int SWORDS = 20;                      // region-start bitmap words (bins 0..639)
    static constexpr int OFF_A = 0;                        // Zbuf (forward) / X
    static constexpr int OFF_B = BUF_BYTES;                // mag + peak list / Y / Zbuf (inverse)
    static constexpr int OFF_DESC = 2 * BUF_BYTES;         // [2][MAXPK] u32
    static constexpr int OFF_S = OFF_DESC + 2 * MAXPK * 4; // [2][SWORDS] u32
    static constexpr int OFF_PS = OFF_S + 2 * SWORDS * 4;  // [2][SWORDS] u32 prefix counts
    static constexpr int WARP_BYTES = OFF_PS + 2 * SWORDS * 4;   // 18240
    static constexpr int OFF_LIST = OFF_B + 2 * MAG_FLOATS * 4;  // [2][MAXPK] u16, dead before Y is cleared
    static constexpr int WARPS = 4;                        // warps (pairs) per CTA
    static constexpr int THREADS = WARPS * 32;
    static constexpr size_t SMEM_BYTES = size_t(WARPS) * WARP_BYTES;
};

struct WarpParams {
    FrameParams f;   // pf_shift must be in [1, 62]
};

// swizzled float4 slot of element [a][b][c] (each 0..7) of the exchange buffer
__device__ __forceinline__ int zslot(int a, int b, int c) { return 64 * a + 8 * b + (c ^ b); }

// padded index into the per-channel |X|^2 array: 4 floats of padding after every 16 bins and
// 8 in front, so a lane's run of 24 floats is 6 aligned, conflict-free 16-byte loads
__device__ __forceinline__ int mag_idx(int i) { return i + 4 * (i >> 4) + 8; }

// swizzled float2 slot of spectrum bin k (keeps pairs (2i, 2i+1) adjacent)
__device__ __forceinline__ int xs(int k) { return k ^ (((k >> 4) & 3) << 1); }

__device__ __forceinline__ cpx2 sel(bool c, cpx2 a, cpx2 b) {
    cpx2 o;
    o.re.x = c ? a.re.x : b.re.x; o.re.y = c ? a.re.y : b.re.y;
    o.im.x = c ? a.im.x : b.im.x; o.im.y = c ? a.im.y : b.im.y;
    return o;
}

// forward real-split of one (k, M-k) pair held in registers: writes 2*X[k], 2*X[M-k] and
// their squared magnitudes for both channels
__device__ __forceinline__ void split_store(cpx2 za, cpx2 zb, int k, const float2 *__restrict__ tw,
                                            float2 *X0, float2 *X1, float *mag0, float *mag1) {
    constexpr int M = WarpGeo::M;
    const float2 e_r = add2(za.re, zb.re), e_i = sub2(za.im, zb.im);
    const float2 o_r = add2(za.im, zb.im), o_i = sub2(zb.re, za.re);
    const float2 w = __ldg(&tw[k]);
    const cpx2 tt = cmul_s(cpx2{o_r, o_i}, w.x, w.y);
    const float2 xr = add2(e_r, tt.re), xi = add2(e_i, tt.im);      // X[k]
    const float2 yr = sub2(e_r, tt.re), yi = sub2(tt.im, e_i);      // X[M-k]
    const int k2 = M - k;
    X0[xs(k)] = make_float2(xr.x, xi.x);
    X1[xs(k)] = make_float2(xr.y, xi.y);
    X0[xs(k2)] = make_float2(yr.x, yi.x);
    X1[xs(k2)] = make_float2(yr.y, yi.y);
    const float2 mk = fma2(xr, xr, mul2(xi, xi));
    const float2 mk2 = fma2(yr, yr, mul2(yi, yi));
    mag0[mag_idx(k)] = mk.x;
    mag1[mag_idx(k)] = mk.y;
    mag0[mag_idx(k2)] = mk2.x;
    mag1[mag_idx(k2)] = mk2.y;
}

// Hermitian C2R pre-pass of one (k, M-k) pair: Y -> Z'[k], Z'[M-k]
__device__ __forceinline__ void unsplit_load(int k, const float2 *__restrict__ tw, const float2 *Y0,
                                             const float2 *Y1, cpx2 &zk, cpx2 &zmk) {
    constexpr int M = WarpGeo::M;
    float2 a0 = Y0[k], a1 = Y1[k], b0 = Y0[M - k], b1 = Y1[M - k];
    if (k == 0) { a0.y = 0.f; a1.y = 0.f; b0.y = 0.f; b1.y = 0.f; }
    const float2 ar = make_float2(a0.x, a1.x), ai = make_float2(a0.y, a1.y);
    const float2 br = make_float2(b0.x, b1.x), bi = make_float2(b0.y, b1.y);
    const float2 e_r = add2(ar, br), e_i = sub2(ai, bi);
    const float2 d_r = sub2(ar, br), d_i = add2(ai, bi);
    const float2 w = __ldg(&tw[k]);
    const cpx2 pp = cmul_s(cpx2{d_r, d_i}, w.x, -w.y);
    zk = cpx2{sub2(e_r, pp.im), add2(e_i, pp.re)};
    zmk = cpx2{add2(e_r, pp.im), sub2(pp.re, e_i)};
}

// value fft.js leaves in slot N/2 + q, 1 <= q <= N/8 (first stale level, see stale_bin())
__device__ __forceinline__ float2 stale_level1(const float2 *X, int q, const float2 *__restrict__ tw) {
    constexpr int N = WarpGeo::N;
    const float2 a = X[xs(q)], b = X[xs(N / 4 + q)], c = X[xs(N / 2 - q)], d = X[xs(N / 4 - q)];
    const float sr = (a.x - b.x) + (c.x - d.x);
    const float si = (a.y - b.y) - (c.y - d.y);
    const float2 w = __ldg(&tw[2 * q]);                  // conj(w) = W_N^{-2q}
    return make_float2(0.25f * (sr * w.x + si * w.y), 0.25f * (si * w.x - sr * w.y));
}

// generic stale slot on the swizzled spectrum (deeper levels; rare: pitch factors below 0.75)
__device__ __noinline__ float2 stale_deep(const float2 *X, int pos, const float2 *__restrict__ tw) {
    constexpr int N = WarpGeo::N;
    int L = N, r = 1, s = 0, o = pos;
    while (L > 4 && o > (L >> 1)) {
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
        if (idx <= N / 2) xv = X[xs(idx)];
        else { xv = X[xs(N - idx)]; xv.y = -xv.y; }
        const float2 w = __ldg(&tw[(s * idx) & (N - 1)]);
        ar += xv.x * w.x + xv.y * w.y;
        ai += xv.y * w.x - xv.x * w.y;
    }
    const float inv_r = 1.0f / float(r);
    return make_float2(ar * inv_r, ai * inv_r);
}

__global__ void __launch_bounds__(WarpGeo::THREADS)
pv_process_warp_kernel(const WarpParams wp) {
    using W = WarpGeo;
    constexpr int N = W::N, M = W::M, NB = W::NB;
    const FrameParams &p = wp.f;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int pair = blockIdx.x * W::WARPS + warp;
    if (2 * pair >= p.num_channels) return;          // whole warp leaves; no CTA-wide barriers below
    unsigned char *mine = smem_raw + size_t(warp) * W::WARP_BYTES;
    float4 *Zf = reinterpret_cast<float4 *>(mine + W::OFF_A);
    float2 *X0 = reinterpret_cast<float2 *>(mine + W::OFF_A);
    float2 *X1 = X0 + NB + 1;                                  // 514 slots each (swizzle stays inside 0..513)
    float *mag0 = reinterpret_cast<float *>(mine + W::OFF_B);
    float *mag1 = mag0 + W::MAG_FLOATS;
    uint16_t *plist = reinterpret_cast<uint16_t *>(mine + W::OFF_LIST);   // [2][MAXPK]
    float2 *Y0 = reinterpret_cast<float2 *>(mine + W::OFF_B);
    float2 *Y1 = Y0 + NB + 1;
    float4 *Zi = reinterpret_cast<float4 *>(mine + W::OFF_B);
    uint32_t *desc = reinterpret_cast<uint32_t *>(mine + W::OFF_DESC);    // [2][MAXPK]
    uint32_t *Sw = reinterpret_cast<uint32_t *>(mine + W::OFF_S);         // [2][SWORDS]
    uint32_t *Ps = reinterpret_cast<uint32_t *>(mine + W::OFF_PS);        // [2][SWORDS]

    const int c0 = 2 * pair, c1 = c0 + 1;
    const bool has1 = c1 < p.num_channels;
    const int hop = p.hop;
    const int rb = p.ring_base;
    const int keep = N - hop;
    const float2 *__restrict__ tw = p.tw;
    const unsigned FULL = 0xFFFFFFFFu;

    // lane's two last-pass butterflies A = (k1a, k2a), B = (k1b, k2b); natural bins
    // kA + 64 j and kB + 64 j with kA + kB == 64 (lane 31 owns the two self-paired ones)
    int k1a, k2a, k1b, k2b;
    if (lane < 24) { k1a = (lane >> 3) + 1; k2a = lane & 7; k1b = 8 - k1a; k2b = 7 - k2a; }
    else if (lane < 28) { k1a = 4; k2a = lane - 24; k1b = 4; k2b = 7 - k2a; }
    else { k1a = 0; k2a = 35 - lane; k1b = 0; k2b = (lane == 31) ? 0 : 8 - k2a; }
    const bool l31 = lane == 31;
    const int kA = k1a + 8 * k2a;

    cpx2 a[8], b[8];

    // ---- forward pass 1: butterflies n = lane, lane + 32 over m1 (stride 64), from global ----
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int n = lane + 32 * h;
        const int m2 = n >> 3, m3 = n & 7;
        cpx2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int s = 2 * (n + 64 * j);
            float2 v0, v1 = make_float2(0.f, 0.f);
            if (s < keep) {
                const int r = (s + rb + hop) & (N - 1);
                v0 = *reinterpret_cast<const float2 *>(p.hist + size_t(c0) * N + r);
                if (has1) v1 = *reinterpret_cast<const float2 *>(p.hist + size_t(c1) * N + r);
            } else {
                const int i = s - keep;
                v0 = make_float2(0.f, 0.f);
                if (p.in) {
                    v0 = __ldg(reinterpret_cast<const float2 *>(p.in + size_t(c0) * hop + i));
                    if (has1) v1 = __ldg(reinterpret_cast<const float2 *>(p.in + size_t(c1) * hop + i));
                }
                *reinterpret_cast<float2 *>(p.hist + size_t(c0) * N + rb + i) = v0;
                if (has1) *reinterpret_cast<float2 *>(p.hist + size_t(c1) * N + rb + i) = v1;
            }
            const float2 w = __ldg(reinterpret_cast<const float2 *>(p.window + s));
            x[j].re = mul2(make_float2(v0.x, v1.x), bc2(w.x));
            x[j].im = mul2(make_float2(v0.y, v1.y), bc2(w.y));
        }
        dft8<false>(x);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) {
            const float2 w = __ldg(&tw[2 * n * k1]);             // W_512^{n k1}
            x[k1] = cmul_s(x[k1], w.x, w.y);
        }
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) {
            const cpx2 v = x[k1];
            Zf[zslot(k1, m2, m3)] = make_float4(v.re.x, v.re.y, v.im.x, v.im.y);
        }
    }
    __syncwarp();

    // ---- forward pass 2: butterflies (k1, m3) over m2 ---------------------------------------
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int bf = lane + 32 * h;
        const int k1 = bf >> 3, m3 = bf & 7;
        cpx2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 v = Zf[zslot(k1, j, m3)];
            x[j] = cpx2{make_float2(v.x, v.y), make_float2(v.z, v.w)};
        }
        dft8<false>(x);
#pragma unroll
        for (int k2 = 1; k2 < 8; k2++) {
            const float2 w = __ldg(&tw[16 * m3 * k2]);           // W_64^{m3 k2}
            x[k2] = cmul_s(x[k2], w.x, w.y);
        }
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            const cpx2 v = x[k2];
            Zf[zslot(k1, k2, m3)] = make_float4(v.re.x, v.re.y, v.im.x, v.im.y);
        }
    }
    __syncwarp();

    // ---- forward pass 3: butterflies A and B over m3; outputs stay in registers --------------
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 va = Zf[zslot(k1a, k2a, j)];
        const float4 vb = Zf[zslot(k1b, k2b, j)];
        a[j] = cpx2{make_float2(va.x, va.y), make_float2(va.z, va.w)};
        b[j] = cpx2{make_float2(vb.x, vb.y), make_float2(vb.z, vb.w)};
    }
    dft8<false>(a);      // a[j] = Z[kA + 64 j]
    dft8<false>(b);      // b[j] = Z[kB + 64 j]
    __syncwarp();        // everyone has read Zf: X may overwrite it

    // ---- real split in registers -> X (2x scaled), |X|^2 ---------------------------------------
    // slots 0..3: (a[j], b[7-j]) at k = kA + 64 j          lane 31: (a[j], a[7-j]), kA == 32
    // slots 4..7: (a[j], b[7-j]) at k = kA + 64 j          lane 31: (b[7-j], b[(j+1)&7]) at k = 64 (7-j)
    // lane 31 also owns the self pair k = 256 (b[4])
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const cpx2 zb = sel(l31, a[7 - j], b[7 - j]);
        split_store(a[j], zb, kA + 64 * j, tw, X0, X1, mag0, mag1);
    }
#pragma unroll
    for (int j = 4; j < 8; j++) {
        const cpx2 za = sel(l31, b[7 - j], a[j]);
        const cpx2 zb = sel(l31, b[(j + 1) & 7], b[7 - j]);
        const int k = l31 ? 64 * (7 - j) : kA + 64 * j;
        split_store(za, zb, k, tw, X0, X1, mag0, mag1);
    }
    if (l31) split_store(b[4], b[4], 256, tw, X0, X1, mag0, mag1);
    __syncwarp();

    // ---- peaks, region descriptors (per channel) --------------------------------------------------
    const long long pf_m = p.pf_mant;
    const int pf_s = p.pf_shift;
    const long long pf_half = 1ll << (pf_s - 1);
    int npk[2];
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
        const float *mg = ch ? mag1 : mag0;
        uint16_t *pl = plist + ch * W::MAXPK;
        uint32_t *dsc = desc + ch * W::MAXPK;
        uint32_t *sw = Sw + ch * W::SWORDS;
        uint32_t *ps = Ps + ch * W::SWORDS;

        // 5-point strict maxima (pv:95-116) on this lane's run of 16 bins
        float m[24];
        {
            const float4 *mv = reinterpret_cast<const float4 *>(mg + 20 * lane);
            const float4 q0 = mv[0], q1 = mv[2], q2 = mv[3], q3 = mv[4], q4 = mv[5], q5 = mv[7];
            m[0] = q0.x; m[1] = q0.y; m[2] = q0.z; m[3] = q0.w;
            m[4] = q1.x; m[5] = q1.y; m[6] = q1.z; m[7] = q1.w;
            m[8] = q2.x; m[9] = q2.y; m[10] = q2.z; m[11] = q2.w;
            m[12] = q3.x; m[13] = q3.y; m[14] = q3.z; m[15] = q3.w;
            m[16] = q4.x; m[17] = q4.y; m[18] = q4.z; m[19] = q4.w;
            m[20] = q5.x; m[21] = q5.y; m[22] = q5.z; m[23] = q5.w;
        }
        uint32_t mask = 0;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const float v = m[e + 4];
            const bool pk = (m[e + 2] < v) && (m[e + 3] < v) && (m[e + 5] < v) && (m[e + 6] < v);
            mask |= pk ? (1u << e) : 0u;
        }
        if (lane == 0) mask &= ~3u;            // i >= 2
        if (lane == 31) mask &= ~(1u << 15);   // i <= nb - 3 == 510
        // ordinals
        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        npk[ch] = total;
        if (lane < W::SWORDS) sw[lane] = 0;
        {
            int ord = incl - cnt;
            uint32_t mm = mask;
            while (mm) {
                const int bit = __ffs(mm) - 1;
                mm &= mm - 1;
                pl[ord++] = uint16_t(16 * lane + bit);
            }
        }
        __syncwarp();
        // one descriptor per region of influence (pv:124-141): p | (delta + 1024) << 10 | valid << 22
        for (int i = lane; i < total; i += 32) {
            const int pk = pl[i];
            int start = 0;
            if (i > 0) { const int before = pl[i - 1]; start = pk - ((pk - before) >> 1); }
            const long long psl = (pf_m * pk + pf_half) >> pf_s;          // Math.round(p * pitchFactor)
            const bool valid = (psl <= NB) && (psl - pk > -1024);
            const int delta = valid ? int(psl) - pk : 0;
            dsc[i] = uint32_t(pk) | (uint32_t(delta + 1024) << 10) | (valid ? (1u << 22) : 0u);
            atomicOr(&sw[start >> 5], 1u << (start & 31));
        }
        __syncwarp();
        {
            const uint32_t wv = (lane < W::SWORDS) ? sw[lane] : 0u;
            const int c = __popc(wv);
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += o;
            }
            if (lane < W::SWORDS) ps[lane] = uint32_t(inc - c);
        }
    }
    __syncwarp();

    // ---- clear Y (mag and the peak list are dead now) -------------------------------------------------
    {
        float4 *yz = reinterpret_cast<float4 *>(mine + W::OFF_B);
        constexpr int NV = (2 * (NB + 1) * 8) / 16;      // 514 float4
        for (int i = lane; i < NV; i += 32) yz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    // ---- shift every region of influence (pv:143-171) --------------------------------------------------
    {
        const int limit = p.src_limit;
        const int nsteps = (limit + 31) >> 5;
        const bool contract = p.pitch_factor < 1.0f;
        const int rmask = p.overlaps - 1;
        const int rstride = N / p.overlaps;
        const int stepm = p.step_mod_r;
        const uint32_t le_mask = (2u << lane) - 1u;
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
            if (npk[ch] == 0) continue;                  // silence: no peaks, spectrum stays zero
            const float2 *Xc = ch ? X1 : X0;
            float2 *Yc = ch ? Y1 : Y0;
            const uint32_t *dsc = desc + ch * W::MAXPK;
            const uint32_t *sw = Sw + ch * W::SWORDS;
            const uint32_t *ps = Ps + ch * W::SWORDS;
            for (int s = 0; s < nsteps; s++) {
                const int bin = 32 * s + lane;
                int ord;
                if (s < W::SWORDS) ord = int(ps[s]) + __popc(sw[s] & le_mask) - 1;
                else ord = npk[ch] - 1;
                const uint32_t dv = dsc[ord];
                const int pk = dv & 1023;
                const int delta = int((dv >> 10) & 4095) - 1024;
                const int d = bin + delta;
                const bool ok = (dv >> 22) && bin < limit && unsigned(d) < unsigned(NB);
                float2 v;
                if (bin <= M) v = Xc[xs(bin)];
                else if (bin <= M + N / 8) v = stale_level1(Xc, bin - M, tw);
                else v = ok ? stale_deep(Xc, bin, tw) : make_float2(0.f, 0.f);
                const int ri = (delta * stepm) & rmask;
                const float2 w = __ldg(&tw[ri * rstride]);          // (cos, -sin)
                const float yr = v.x * w.x + v.y * w.y;
                const float yi = v.y * w.x - v.x * w.y;
                if (!contract) {
                    if (ok) Yc[d] = make_float2(yr, yi);           // expansion: at most one source per bin
                } else {
                    const bool right = bin >= pk;
                    if (ok && right) { float2 y = Yc[d]; y.x += yr; y.y += yi; Yc[d] = y; }
                    __syncwarp();
                    if (ok && !right) { float2 y = Yc[d]; y.x += yr; y.y += yi; Yc[d] = y; }
                    __syncwarp();
                }
            }
        }
    }
    __syncwarp();

    // ---- Hermitian C2R pre-pass in registers (mirror of the split) -------------------------------------
    {
        cpx2 zk[8], zmk[8], z256;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = (l31 && j >= 4) ? 64 * (7 - j) : kA + 64 * j;
            unsplit_load(k, tw, Y0, Y1, zk[j], zmk[j]);
        }
        z256 = zk[0];
        if (l31) { cpx2 dummy; unsplit_load(256, tw, Y0, Y1, z256, dummy); }
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = zk[i];
#pragma unroll
        for (int i = 4; i < 8; i++) a[i] = sel(l31, zmk[7 - i], zk[i]);
        b[0] = sel(l31, zk[7], zmk[7]);
        b[1] = sel(l31, zk[6], zmk[6]);
        b[2] = sel(l31, zk[5], zmk[5]);
        b[3] = sel(l31, zk[4], zmk[4]);
        b[4] = sel(l31, z256, zmk[3]);
        b[5] = sel(l31, zmk[4], zmk[2]);
        b[6] = sel(l31, zmk[5], zmk[1]);
        b[7] = sel(l31, zmk[6], zmk[0]);
    }
    __syncwarp();        // everyone has read Y: the inverse exchange buffer may overwrite it

    // ---- inverse pass 1 (DIT): butterflies A and B over k3, twiddle conj(W_64^{k2 m3}) -------------------
    dft8<true>(a);
    dft8<true>(b);
#pragma unroll
    for (int m3 = 0; m3 < 8; m3++) {
        cpx2 va = a[m3], vb = b[m3];
        if (m3 > 0) {
            const float2 wa = __ldg(&tw[16 * k2a * m3]);
            const float2 wb = __ldg(&tw[16 * k2b * m3]);
            va = cmul_s(va, wa.x, -wa.y);
            vb = cmul_s(vb, wb.x, -wb.y);
        }
        Zi[zslot(k1a, k2a, m3)] = make_float4(va.re.x, va.re.y, va.im.x, va.im.y);
        Zi[zslot(k1b, k2b, m3)] = make_float4(vb.re.x, vb.re.y, vb.im.x, vb.im.y);
    }
    __syncwarp();

    // ---- inverse pass 2: butterflies (k1, m3) over k2, twiddle conj(W_512^{k1 (m3 + 8 m2)}) ---------------
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int bf = lane + 32 * h;
        const int k1 = bf >> 3, m3 = bf & 7;
        cpx2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 v = Zi[zslot(k1, j, m3)];
            x[j] = cpx2{make_float2(v.x, v.y), make_float2(v.z, v.w)};
        }
        dft8<true>(x);
#pragma unroll
        for (int m2 = 0; m2 < 8; m2++) {
            const float2 w = __ldg(&tw[2 * k1 * (m3 + 8 * m2)]);
            const cpx2 v = cmul_s(x[m2], w.x, -w.y);
            Zi[zslot(k1, m2, m3)] = make_float4(v.re.x, v.re.y, v.im.x, v.im.y);
        }
    }
    __syncwarp();

    // ---- inverse pass 3: butterflies n over k1 -> z[n + 64 m1]; window, overlap-add, emit ------------------
    {
        const float scale = 1.0f / float(2 * N);
        const float inv_r = 1.0f / float(p.overlaps);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int n = lane + 32 * h;
            const int m2 = n >> 3, m3 = n & 7;
            cpx2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 v = Zi[zslot(j, m2, m3)];
                x[j] = cpx2{make_float2(v.x, v.y), make_float2(v.z, v.w)};
            }
            dft8<true>(x);
#pragma unroll
            for (int m1 = 0; m1 < 8; m1++) {
                const int s = 2 * (n + 64 * m1);
                const float2 w = __ldg(reinterpret_cast<const float2 *>(p.window + s));
                const bool head = s < hop;
                const bool tail = s >= keep;
                const int ring = (s + rb) & (N - 1);
                // fromComplexArray -> f32, applyHannWindow, / nbOverlaps (pv:65-67, ola:153)
                float2 y0 = make_float2(((x[m1].re.x * scale) * w.x) * inv_r, ((x[m1].im.x * scale) * w.y) * inv_r);
                float2 y1 = make_float2(((x[m1].re.y * scale) * w.x) * inv_r, ((x[m1].im.y * scale) * w.y) * inv_r);
                float2 *ap0 = reinterpret_cast<float2 *>(p.acc + size_t(c0) * N + ring);
                float2 *ap1 = reinterpret_cast<float2 *>(p.acc + size_t(c1) * N + ring);
                if (!tail) {
                    const float2 q0 = *ap0;
                    y0.x += q0.x; y0.y += q0.y;
                    if (has1) { const float2 q1 = *ap1; y1.x += q1.x; y1.y += q1.y; }
                }
                if (head) {
                    *reinterpret_cast<float2 *>(p.out + size_t(c0) * hop + s) = y0;
                    if (has1) *reinterpret_cast<float2 *>(p.out + size_t(c1) * hop + s) = y1;
                } else {
                    *ap0 = y0;
                    if (has1) *ap1 = y1;
                }
            }
        }
    }
}

}  // namespace pvb
