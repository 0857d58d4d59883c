// pv_kernel_cta.cuh — fused kernel for every supported frame size (256 .. 4096), sm_100a.
//
// FFT side: as in pv_kernel.cuh (N/16 threads per channel pair, in-place radix-8 DIF passes in
// padded shared memory, both channels in f32x2 registers).
// Middle: the warp-synchronous design of pv_kernel_warp.cuh generalised over N — one warp per
// channel does peak picking on runs of bins (integer compares, funnel-shift masks), builds one
// 32-bit descriptor per region of influence plus a region-start bitmap, rebuilds the first level
// of stale upper bins into an extension of the spectrum, and shifts IN PLACE (right halves of the
// regions by plain stores, left halves by read-add-store; no atomics, no second spectrum buffer,
// no |X|^2 array).  Compared with pv_kernel.cuh this halves the shared memory per pair (more
// resident CTAs) and removes the per-bin owner search and the shared-memory float atomics.
//
// Valid for pitch factors in [0.75, 64] and R <= 32 (like the warp kernel); the host routes
// everything else to pv_kernel.cuh.
#pragma once

#include "pv_kernel.cuh"

namespace pvb {

template <int N>
struct CtaGeo {
    static constexpr int M = N / 2;
    static constexpr int NB = M + 1;
    static constexpr int T = M / 8;                        // threads per channel pair
    static constexpr int ZSLOTS = M + M / 8 + M / 128 + 1;
    static constexpr size_t Z_BYTES = size_t(ZSLOTS) * 16;
    static constexpr int XSLOTS = ((M + 1 + N / 8 + 1) + 127) & ~127;     // spectrum + stale extension, swizzle-block aligned
    static constexpr size_t X_BYTES = size_t(2) * XSLOTS * 8;
    static constexpr int MAXPK = (M / 3 + 16) & ~7;
    static constexpr int SW = (M / 32 > 0) ? M / 32 : 1;  // region-start bitmap words (starts are < M)
    static constexpr size_t TAB_CH_BYTES = size_t(MAXPK + 2 * SW) * 4;
    static constexpr size_t PAIR_BYTES = Z_BYTES + X_BYTES + 2 * TAB_CH_BYTES;
    static constexpr int G = (T >= 128) ? 1 : (128 / T);   // pairs per CTA
    static constexpr int THREADS = G * T;
    static constexpr size_t SMEM_BYTES = size_t(G) * PAIR_BYTES;
    static constexpr int BPL = M / 32;                     // bins per lane in the peak scan
    static constexpr int SRB = (BPL < 16) ? BPL : 16;      // bins per sub-run
};

// swizzled float2 slot of spectrum bin k: the low 4 bits are XORed with the index of the peak-scan
// run the bin belongs to (bins per lane = M/32), so that both 32 consecutive bins (split, sweep)
// and one bin per run (peak scan: lanes BPL bins apart) hit 16 different 8-byte banks
template <int N>
__device__ __forceinline__ int xsw(int k) {
    constexpr int BPL = N / 64;
    constexpr int SH = (BPL >= 32) ? 5 : (BPL >= 16) ? 4 : (BPL >= 8) ? 3 : 2;
    return k ^ ((k >> SH) & 15);
}

// value fft.js leaves in slot N/2 + q, 1 <= q <= N/8 (first stale level; see stale_bin())
template <int N>
__device__ __forceinline__ float2 stale_first_level(const float2 *X, int q, const float2 *__restrict__ tw) {
    const float2 a = X[xsw<N>(q)], b = X[xsw<N>(N / 4 + q)], c = X[xsw<N>(N / 2 - q)], d = X[xsw<N>(N / 4 - q)];
    const float sr = (a.x - b.x) + (c.x - d.x);
    const float si = (a.y - b.y) - (c.y - d.y);
    const float2 w = __ldg(&tw[2 * q]);                  // conj(w) = W_N^{-2q}
    return make_float2(0.25f * (sr * w.x + si * w.y), 0.25f * (si * w.x - sr * w.y));
}

// Peaks -> region descriptors -> in-place shift of ONE channel, executed by ONE warp.
// Xc: swizzled spectrum (2x scaled) of the channel, XSLOTS entries; on return it holds the
// shifted spectrum in bins 0..M.  dsc / sw / ps: this channel's tables.
template <int N>
__device__ __forceinline__ void shift_channel(float2 *Xc, uint32_t *dsc, uint32_t *sw, uint32_t *ps,
                                              const FrameParams &p, const float2 *__restrict__ tw, int lane) {
    using G = CtaGeo<N>;
    constexpr int M = G::M, NB = G::NB, BPL = G::BPL, SRB = G::SRB, SW = G::SW;
    const unsigned FULL = 0xFFFFFFFFu;
    const uint32_t le_mask = (2u << lane) - 1u;
    const bool contract = p.pitch_factor < 1.0f;

    // ---- 5-point strict maxima (pv:95-116); |X|^2 in float32 (pv:88), compared as integers ----
    unsigned long long mask = 0;
#pragma unroll
    for (int sr = 0; sr < BPL / SRB; sr++) {
        const int base = BPL * lane + SRB * sr;
        int m[SRB + 4];
#pragma unroll
        for (int e = 0; e < SRB + 4; e++) {
            int k = base - 2 + e;
            k = k < 0 ? 0 : (k > M ? M : k);
            const float2 v = Xc[xsw<N>(k)];
            m[e] = __float_as_int(fmaf(v.x, v.x, v.y * v.y));
        }
        int q[SRB + 3];
#pragma unroll
        for (int t = 0; t < SRB + 3; t++) q[t] = max(m[t], m[t + 1]);
        uint32_t bits = 0;
#pragma unroll
        for (int e = SRB - 1; e >= 0; e--) {
            const int nb_max = max(q[e], q[e + 3]);        // bins e-2, e-1, e+1, e+2 around m[e + 2]
            bits = __funnelshift_l(uint32_t(nb_max - m[e + 2]), bits, 1);
        }
        mask |= (unsigned long long)bits << (SRB * sr);
    }
    if (lane == 0) mask &= ~3ull;                          // i >= 2
    if (lane == 31) mask &= ~(1ull << (BPL - 1));          // i <= nb - 3 == M - 2

    const int cnt = __popcll(mask);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += o;
    }
    const int npk = __shfl_sync(FULL, incl, 31);
    const int own_last = mask ? (BPL * lane + 63 - __clzll((long long)mask)) : -1;
    const uint32_t nz_below = __ballot_sync(FULL, mask != 0) & (le_mask >> 1);
    const int src = nz_below ? (31 - __clz(nz_below)) : 0;
    int prev = __shfl_sync(FULL, own_last, src);
    if (!nz_below) prev = -1;

    for (int i = lane; i < SW; i += 32) sw[i] = 0;
    __syncwarp();
    // one descriptor per region of influence (pv:124-141):  delta << 16 | rot index << 11 | peak
    {
        const long long pf_m = p.pf_mant;
        const int pf_s = p.pf_shift;
        const long long pf_half = 1ll << (pf_s - 1);
        const int stepm = p.step_mod_r;
        const int rmask = p.overlaps - 1;
        int ord = incl - cnt;
        unsigned long long mm = mask;
        while (mm) {
            const int bit = __ffsll((long long)mm) - 1;
            mm &= mm - 1;
            const int pk = BPL * lane + bit;
            const int start = (prev < 0) ? 0 : pk - ((pk - prev) >> 1);
            const long long psl = (pf_m * pk + pf_half) >> pf_s;       // Math.round(p * pitchFactor)
            const bool valid = psl <= NB;                               // pv:127
            const int delta = valid ? int(psl) - pk : 0x4000;           // 0x4000: lands outside [0, nb)
            const int ri = (delta * stepm) & rmask;
            dsc[ord] = (uint32_t(delta) << 16) | (uint32_t(ri) << 11) | uint32_t(pk);
            atomicOr(&sw[start >> 5], 1u << (start & 31));
            prev = pk;
            ord++;
        }
    }
    __syncwarp();
    // ps[w] = (#region starts in words < w) - 1
    {
        constexpr int WPL = (SW + 31) / 32;
        int local = 0;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            const int wi = WPL * lane + j;
            if (wi < SW) local += __popc(sw[wi]);
        }
        int inc = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        int run = inc - local - 1;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            const int wi = WPL * lane + j;
            if (wi < SW) { ps[wi] = uint32_t(run); run += __popc(sw[wi]); }
        }
    }
    __syncwarp();

    if (npk == 0) {                                        // silence: the shifted spectrum is zero (pv:121)
        for (int i = lane; i <= M + 1; i += 32) Xc[xsw<N>(i)] = make_float2(0.f, 0.f);
        __syncwarp();
        return;
    }

    // ---- stale slots N/2+1 .. N/2+N/8 into the extension of X (only read when contracting) ----
    if (contract) {
        for (int q = lane + 1; q <= N / 8; q += 32) Xc[xsw<N>(M + q)] = stale_first_level<N>(Xc, q, tw);
    }
    __syncwarp();

    float rot_c, rot_s;
    {
        const float2 t = __ldg(&tw[(lane & (p.overlaps - 1)) * (N / p.overlaps)]);
        rot_c = t.x;
        rot_s = -t.y;
    }
    const bool quarter = p.overlaps == 4;
    const int src_bins = contract ? (M + 1 + N / 8) : (M + 1);
    const int steps = (src_bins + 31) >> 5;
    constexpr int CH = 8;
    const int nchunks = (steps + CH - 1) / CH;

    // Contraction writes at or below the bin it read (chunks ascending), expansion at or above
    // (chunks descending; it never maps two sources to one bin, so there is no second pass).
#pragma unroll 1
    for (int it = 0; it < nchunks; it++) {
        const int c = contract ? it : nchunks - 1 - it;
        uint32_t dvs[CH];
        float2 xv[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int s = CH * c + i;
            const int bin = 32 * s + lane;
            int ord = npk - 1;
            if (s < SW) ord = int(ps[s]) + __popc(sw[s] & le_mask);
            dvs[i] = (bin < src_bins) ? dsc[ord] : 0x40000000u;
            xv[i] = make_float2(0.f, 0.f);
            if (bin < src_bins) {
                float2 *xp = Xc + xsw<N>(bin);
                xv[i] = *xp;
                *xp = make_float2(0.f, 0.f);
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const uint32_t dv = dvs[i];
            const int bin = 32 * (CH * c + i) + lane;
            const int d = bin + (int(dv) >> 16);
            const bool right = bin >= int(dv & 2047);
            const bool okd = unsigned(d) < unsigned(NB);
            const int ri = (dv >> 11) & 31;
            float2 y;
            if (quarter) {
                const float ax = (ri & 1) ? -xv[i].y : xv[i].x, ay = (ri & 1) ? xv[i].x : xv[i].y;
                y = make_float2((ri & 2) ? -ax : ax, (ri & 2) ? -ay : ay);
            } else {
                const float rc = __shfl_sync(FULL, rot_c, ri), rs = __shfl_sync(FULL, rot_s, ri);
                y = make_float2(xv[i].x * rc - xv[i].y * rs, xv[i].x * rs + xv[i].y * rc);
            }
            const int slot = xsw<N>(d);
            xv[i] = y;
            if (okd && (right || !contract)) Xc[slot] = y;                 // first writer of that bin
            dvs[i] = (okd && !right && contract) ? uint32_t(slot) : 0xFFFFFFFFu;
        }
        if (contract) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < CH; i++)
                if (dvs[i] != 0xFFFFFFFFu) { float2 o = Xc[dvs[i]]; o.x += xv[i].x; o.y += xv[i].y; Xc[dvs[i]] = o; }
        }
        __syncwarp();
    }
}

template <int N>
__global__ void __launch_bounds__(CtaGeo<N>::THREADS)
pv_process_cta_kernel(const FrameParams p, const float *__restrict__ window_out) {
    using G_ = CtaGeo<N>;
    constexpr int M = G_::M, T = G_::T, G = G_::G, XS = G_::XSLOTS;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int g = tid / T;
    const int t = tid - g * T;
    unsigned char *mine = smem_raw + size_t(g) * G_::PAIR_BYTES;
    float4 *Z = reinterpret_cast<float4 *>(mine);
    float2 *X = reinterpret_cast<float2 *>(mine + G_::Z_BYTES);                      // [2][XSLOTS]

    const int pair = blockIdx.x * G + g;
    const int c0 = 2 * pair, c1 = c0 + 1;
    const bool has0 = c0 < p.num_channels, has1 = c1 < p.num_channels;
    const int hop = p.hop;
    const int rb = p.ring_base;
    const int keep = N - hop;
    const float2 *__restrict__ tw = p.tw;

    // ---- frame gather + analysis window (ola:91-146, pv:55) --------------------------------------
    // M / T == 8 elements per thread: all loads are issued before anything consumes them
    {
        float2 r0[8], r1[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int n = 2 * (t + T * j);
            r0[j] = make_float2(0.f, 0.f);
            r1[j] = make_float2(0.f, 0.f);
            if (n < keep) {
                const int r = (n + rb + hop) & (N - 1);
                if (has0) r0[j] = *reinterpret_cast<const float2 *>(p.hist + size_t(c0) * N + r);
                if (has1) r1[j] = *reinterpret_cast<const float2 *>(p.hist + size_t(c1) * N + r);
            } else if (p.in) {
                const int i = n - keep;
                if (has0) r0[j] = __ldg(reinterpret_cast<const float2 *>(p.in + size_t(c0) * hop + i));
                if (has1) r1[j] = __ldg(reinterpret_cast<const float2 *>(p.in + size_t(c1) * hop + i));
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int m = t + T * j;
            const int n = 2 * m;
            if (n >= keep) {
                const int i = n - keep;
                if (has0) *reinterpret_cast<float2 *>(p.hist + size_t(c0) * N + rb + i) = r0[j];
                if (has1) *reinterpret_cast<float2 *>(p.hist + size_t(c1) * N + rb + i) = r1[j];
            }
            const float2 w = __ldg(reinterpret_cast<const float2 *>(p.window + n));
            Z[zp(m)] = make_float4(r0[j].x * w.x, r1[j].x * w.x, r0[j].y * w.y, r1[j].y * w.y);
        }
    }
    __syncthreads();

    fft_inplace<N, false>(Z, t, tw);

    // (after the frame has arrived, so that it does not compete with the frame loads)
    // warm L2 with the overlap-add ring lines the tail adds to
    for (int line = 32 * t; line < N; line += 32 * T) {
        if (((line - rb + hop) & (N - 1)) >= hop) {
            if (has0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.acc + size_t(c0) * N + line));
            if (has1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.acc + size_t(c1) * N + line));
        }
    }

    // ---- real split -> swizzled per-channel spectrum (2x scaled) -----------------------------------
    for (int k = t; k <= M / 2; k += T) {
        const cpx2 a = ldz(Z, dif_pos<N>(k));
        const cpx2 b = ldz(Z, dif_pos<N>((M - k) & (M - 1)));
        const float2 e_r = add2(a.re, b.re), e_i = sub2(a.im, b.im);
        const float2 o_r = add2(a.im, b.im), o_i = sub2(b.re, a.re);
        const float2 w = __ldg(&tw[k]);
        const cpx2 tt = cmul_s(cpx2{o_r, o_i}, w.x, w.y);
        const float2 xr = add2(e_r, tt.re), xi = add2(e_i, tt.im);          // X[k]
        const float2 yr = sub2(e_r, tt.re), yi = sub2(tt.im, e_i);          // X[M-k]
        const int s1 = xsw<N>(k), s2 = xsw<N>(M - k);
        X[s1] = make_float2(xr.x, xi.x);
        X[XS + s1] = make_float2(xr.y, xi.y);
        X[s2] = make_float2(yr.x, yi.x);
        X[XS + s2] = make_float2(yr.y, yi.y);
    }
    __syncthreads();

    // ---- middle: one warp per (pair, channel) -------------------------------------------------------
    {
        constexpr int NWARPS = G_::THREADS / 32;
        const int warp = tid >> 5, lane = tid & 31;
        for (int item = warp; item < 2 * G; item += NWARPS) {
            unsigned char *base = smem_raw + size_t(item >> 1) * G_::PAIR_BYTES;
            float2 *Xc = reinterpret_cast<float2 *>(base + G_::Z_BYTES) + (item & 1) * XS;
            uint32_t *dsc = reinterpret_cast<uint32_t *>(base + G_::Z_BYTES + G_::X_BYTES +
                                                         (item & 1) * G_::TAB_CH_BYTES);
            shift_channel<N>(Xc, dsc, dsc + G_::MAXPK, dsc + G_::MAXPK + G_::SW, p, tw, lane);
        }
    }
    __syncthreads();

    // ---- Hermitian C2R pre-pass: Y[0..M] -> Z'[0..M) (natural order) ---------------------------------
    for (int k = t; k <= M / 2; k += T) {
        const int s1 = xsw<N>(k), s2 = xsw<N>(M - k);
        float2 a0 = X[s1], a1 = X[XS + s1], b0 = X[s2], b1 = X[XS + s2];
        if (k == 0) { a0.y = 0.f; a1.y = 0.f; b0.y = 0.f; b1.y = 0.f; }
        const float2 ar = make_float2(a0.x, a1.x), ai = make_float2(a0.y, a1.y);
        const float2 br = make_float2(b0.x, b1.x), bi = make_float2(b0.y, b1.y);
        const float2 e_r = add2(ar, br), e_i = sub2(ai, bi);
        const float2 d_r = sub2(ar, br), d_i = add2(ai, bi);
        const float2 w = __ldg(&tw[k]);
        const cpx2 pp = cmul_s(cpx2{d_r, d_i}, w.x, -w.y);
        stz(Z, k, cpx2{sub2(e_r, pp.im), add2(e_i, pp.re)});
        if (k != 0) stz(Z, M - k, cpx2{add2(e_r, pp.im), sub2(pp.re, e_i)});
    }
    __syncthreads();

    fft_inplace<N, true>(Z, t, tw);

    // ---- window (all scales folded in), overlap-add ring, emit (pv:65-67, ola:149-157,111-137) ------
    // N / 4 / T == 4 float4 groups per thread and channel: accumulator loads first, then the math
    {
        float4 a0[4], a1[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = 4 * (t + T * i);
            const int ring = (k + rb) & (N - 1);
            a0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            a1[i] = a0[i];
            if (k < keep) {                                   // the tail slot starts from zero (ola:134)
                if (has0) a0[i] = *reinterpret_cast<const float4 *>(p.acc + size_t(c0) * N + ring);
                if (has1) a1[i] = *reinterpret_cast<const float4 *>(p.acc + size_t(c1) * N + ring);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q = t + T * i;
            const float4 za = Z[zp(dif_pos<N>(2 * q))];
            const float4 zb = Z[zp(dif_pos<N>(2 * q + 1))];
            const int k = 4 * q;
            const float4 w = __ldg(reinterpret_cast<const float4 *>(window_out + k));
            const int ring = (k + rb) & (N - 1);
            const float4 y0 = make_float4(za.x * w.x + a0[i].x, za.z * w.y + a0[i].y, zb.x * w.z + a0[i].z, zb.z * w.w + a0[i].w);
            const float4 y1 = make_float4(za.y * w.x + a1[i].x, za.w * w.y + a1[i].y, zb.y * w.z + a1[i].z, zb.w * w.w + a1[i].w);
            if (k < hop) {                                    // head: emit (ola:111-118)
                if (has0) *reinterpret_cast<float4 *>(p.out + size_t(c0) * hop + k) = y0;
                if (has1) *reinterpret_cast<float4 *>(p.out + size_t(c1) * hop + k) = y1;
            } else {
                if (has0) *reinterpret_cast<float4 *>(p.acc + size_t(c0) * N + ring) = y0;
                if (has1) *reinterpret_cast<float4 *>(p.acc + size_t(c1) * N + ring) = y1;
            }
        }
    }
}

}  // namespace pvb
