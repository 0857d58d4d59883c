// pv_kernel_pair.cuh — fused kernel for frame size 1024 with TWO warps per channel pair (sm_100a).
//
// Same arithmetic and the same shared-memory design as pv_kernel_warp.cuh (packed f32x2 radix-8
// FFT, swizzled re/im planes, in-place shift, integer peak masks, region descriptors), but a
// channel pair is owned by 64 threads instead of 32:
//
//   * every radix-8 pass is ONE butterfly per thread (8 packed complex values in registers),
//     so the kernel fits in 72 registers and 28 warps (14 pairs) are resident per SM;
//   * the per-channel phases (peak picking, region descriptors, stale slots, shift) run in
//     parallel: warp 0 of the pair does channel 0, warp 1 does channel 1;
//   * the two warps meet at named barriers (bar.sync id, 64) that belong to their pair only;
//     other pairs of the CTA are never stalled.
//
// Valid for the same range as the warp kernel: pitch factors in [0.75, 64], R <= 32.
#pragma once

#include "pv_kernel_warp.cuh"

namespace pvb {

struct PairGeo {
    static constexpr int N = 1024, M = 512, NB = 513;
    static constexpr int PLANE = 512;
    static constexpr int XCH = 516;                        // float2 slots per channel of X / Y
    static constexpr int BUF_BYTES = 2 * XCH * 8;          // 8256 >= two planes (8192)
    static constexpr int MAXPK = 176;
    static constexpr int SWORDS = 20;
    static constexpr int TAB_CH_BYTES = (MAXPK + SWORDS) * 4;        // descriptors + start bitmap, per channel
    static constexpr int OFF_TAB = BUF_BYTES;
    static constexpr int PAIR_BYTES = BUF_BYTES + 2 * TAB_CH_BYTES;  // 9824
    static constexpr int MAX_PAIRS = 7;                    // pairs per CTA (named barriers 1..7)
    static constexpr int MAX_THREADS = MAX_PAIRS * 64;
};

// swizzled float2 slot of spectrum bin k for this kernel: conflict-free for 32 consecutive bins
// and for the (k2, k3 parity) pattern of the split / un-split phases
__device__ __forceinline__ int xs2(int k) { return k ^ ((k >> 3) & 6) ^ ((k >> 6) & 1); }

// slot of natural bin k = k1 + 8 k2 + 64 k3 in the exchange planes after the last forward pass
__device__ __forceinline__ int zslot_of_bin(int k) { return zslot(k & 7, (k >> 3) & 7, (k >> 6) & 7); }

__device__ __forceinline__ void pair_barrier(int id) {
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// first stale level on the xs2-swizzled spectrum (see stale_level1)
__device__ __forceinline__ float2 stale_level1_x2(const float2 *X, int q, const float2 *__restrict__ tw) {
    constexpr int N = PairGeo::N;
    const float2 a = X[xs2(q)], b = X[xs2(N / 4 + q)], c = X[xs2(N / 2 - q)], d = X[xs2(N / 4 - q)];
    const float sr = (a.x - b.x) + (c.x - d.x);
    const float si = (a.y - b.y) - (c.y - d.y);
    const float2 w = __ldg(&tw[2 * q]);
    return make_float2(0.25f * (sr * w.x + si * w.y), 0.25f * (si * w.x - sr * w.y));
}

__global__ void __launch_bounds__(PairGeo::MAX_THREADS, 2)
pv_process_pair_kernel(const WarpParams wp) {
    using G = PairGeo;
    constexpr int N = G::N, M = G::M, NB = G::NB;
    const FrameParams &p = wp.f;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int g = tid >> 6;                      // pair slot in the CTA
    const int L = tid & 63;                      // thread in the pair
    const int w = L >> 5;                        // warp in the pair == channel it owns in the middle phases
    const int lane = L & 31;
    const int pair = blockIdx.x * (blockDim.x >> 6) + g;
    if (2 * pair >= p.num_channels) return;      // both warps of the pair leave together
    if (wp.stagger_ns > 0) {                     // see pv_kernel_warp.cuh
        const int slot = 2 * g + (int(blockIdx.x) >= wp.num_sms ? 1 : 0);
        if (slot > 0) __nanosleep(unsigned(slot * wp.stagger_ns));
    }
    const int bar = 1 + g;
    unsigned char *mine = smem_raw + size_t(g) * G::PAIR_BYTES;
    float2 *zre = reinterpret_cast<float2 *>(mine);
    float2 *zim = zre + G::PLANE;
    float2 *X0 = reinterpret_cast<float2 *>(mine);
    float2 *X1 = X0 + G::XCH;

    const int c0 = 2 * pair, c1 = c0 + 1;
    const bool has1 = c1 < p.num_channels;
    const int hop = p.hop;
    const int rb = p.ring_base;
    const int keep = N - hop;
    const float2 *__restrict__ tw = p.tw;
    const unsigned FULL = 0xFFFFFFFFu;

    // pass 1 / inverse pass 3: butterfly n = L = (m2, m3), element [k1][m2][m3]
    const int hi = L >> 3, lo = L & 7;
    const int base1e = 8 * hi + (lo ^ hi);
    const int base1o = 8 * (hi ^ 1) + (lo ^ hi);
    // pass 2: butterfly (k1, m3) = (hi, lo), element [k1][j][m3] == base2 ^ 9j
    const int base2 = 64 * hi + 8 * (hi & 1) + lo;
    // pass 3 / inverse pass 1: butterfly (k1, k2) = (hi, lo), element [k1][k2][j] == base3 ^ j
    const int base3 = 64 * hi + 8 * (lo ^ (hi & 1)) + lo;

    // prefetch the overlap-add ring into L2 (see pv_kernel_warp.cuh)
    {
        const int line = 16 * L;                                        // floats [16 L, 16 L + 16): 64 bytes
        if (((line - rb + hop) & (N - 1)) >= hop) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.acc + size_t(c0) * N + line));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.acc + size_t(c1) * N + line));
        }
    }

    // ---- forward pass 1: butterfly n = L over m1, inputs straight from the rings ------------------
    {
        float2 r0[8], r1[8];
        char *histb = reinterpret_cast<char *>(p.hist + size_t(c0) * N);
        const char *inb0 = reinterpret_cast<const char *>(p.in ? p.in + size_t(c0) * hop : p.hist);
        const char *inb1 = inb0 + (has1 && p.in ? hop * 4 : 0);
        const unsigned rbase = unsigned((2 * L + rb + hop) & (N - 1)) * 4u;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (2 * L + 128 * j < keep) {
                const unsigned off = (rbase + 512u * j) & (N * 4 - 1);
                r0[j] = *reinterpret_cast<const float2 *>(histb + off);
                r1[j] = *reinterpret_cast<const float2 *>(histb + off + N * 4);
            } else {
                const int ib = (2 * L + 128 * j - keep) * 4;
                r0[j] = *reinterpret_cast<const float2 *>(inb0 + ib);
                r1[j] = *reinterpret_cast<const float2 *>(inb1 + ib);
            }
        }
        cpx2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int s = 2 * L + 128 * j;
            if (s >= keep) {                       // the new block; paused input is zeros (ola:93-100)
                if (!p.in) { r0[j] = make_float2(0.f, 0.f); r1[j] = make_float2(0.f, 0.f); }
                if (!has1) r1[j] = make_float2(0.f, 0.f);
                const unsigned off = unsigned(rb + s - keep) * 4u;
                *reinterpret_cast<float2 *>(histb + off) = r0[j];
                *reinterpret_cast<float2 *>(histb + off + N * 4) = r1[j];
            }
            const float2 wv = __ldg(reinterpret_cast<const float2 *>(p.window + s));
            x[j].re = mul2(make_float2(r0[j].x, r1[j].x), bc2(wv.x));
            x[j].im = mul2(make_float2(r0[j].y, r1[j].y), bc2(wv.y));
        }
        dft8<false>(x);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) {
            const float2 t = __ldg(&tw[2 * L * k1]);             // W_512^{n k1}
            x[k1] = cmul_s(x[k1], t.x, t.y);
        }
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) zst(zre, zim, 64 * k1 + ((k1 & 1) ? base1o : base1e), x[k1]);
    }
    pair_barrier(bar);

    // ---- forward pass 2: butterfly (k1, m3) over m2 ----------------------------------------------------
    {
        cpx2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = zld(zre, zim, base2 ^ (9 * j));
        dft8<false>(x);
#pragma unroll
        for (int k2 = 1; k2 < 8; k2++) {
            const float2 t = __ldg(&tw[16 * lo * k2]);           // W_64^{m3 k2}
            x[k2] = cmul_s(x[k2], t.x, t.y);
        }
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) zst(zre, zim, base2 ^ (9 * k2), x[k2]);
    }
    pair_barrier(bar);

    // ---- forward pass 3: butterfly (k1, k2) over m3, in place -------------------------------------------
    {
        cpx2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = zld(zre, zim, base3 ^ j);
        dft8<false>(x);
#pragma unroll
        for (int k3 = 0; k3 < 8; k3++) zst(zre, zim, base3 ^ k3, x[k3]);     // Z[k1 + 8 k2 + 64 k3]
    }
    pair_barrier(bar);

    // ---- real split: (Z[k], Z[M-k]) -> 2 X[k], 2 X[M-k]; every thread owns four pairs --------------------
    // bins k for thread L: it 0..2 -> k1 = it+1, (k2, k3) = (lo, hi); it 3 -> the k1 = 4 and k1 = 0 planes
    int kq[4];
#pragma unroll
    for (int it = 0; it < 3; it++) kq[it] = (it + 1) + 8 * lo + 64 * hi;
    if (L < 32) kq[3] = 4 + 8 * lo + 64 * hi;                      // k2 = lo, k3 = hi in 0..3 <-> (7-k2, 7-k3)
    else {
        const int q = L - 32;
        if (q < 24) kq[3] = 8 * (1 + (q >> 3)) + 64 * (q & 7);     // k2 in 1..3 <-> 8-k2
        else if (q < 28) kq[3] = 32 + 64 * (q - 24);               // k2 = 4, k3 in 0..3 <-> 7-k3
        else if (q < 31) kq[3] = 64 * (q - 27);                    // k2 = 0, k3 in 1..3 <-> 8-k3
        else kq[3] = 0;                                            // DC / Nyquist; this thread also owns k = 256
    }
    {
        cpx2 za[5], zb[5];
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int k = kq[it];
            za[it] = zld(zre, zim, zslot_of_bin(k));
            zb[it] = zld(zre, zim, zslot_of_bin((M - k) & (M - 1)));
        }
        za[4] = zld(zre, zim, zslot_of_bin(256));
        zb[4] = za[4];
        pair_barrier(bar);                                         // planes fully read: X may overwrite them
#pragma unroll
        for (int it = 0; it < 5; it++) {
            if (it == 4 && L != 63) break;
            const int k = (it == 4) ? 256 : kq[it];
            const float2 e_r = add2(za[it].re, zb[it].re), e_i = sub2(za[it].im, zb[it].im);
            const float2 o_r = add2(za[it].im, zb[it].im), o_i = sub2(zb[it].re, za[it].re);
            const float2 t = __ldg(&tw[k]);
            const cpx2 tt = cmul_s(cpx2{o_r, o_i}, t.x, t.y);
            const float2 xr = add2(e_r, tt.re), xi = add2(e_i, tt.im);      // X[k]
            const float2 yr = sub2(e_r, tt.re), yi = sub2(tt.im, e_i);      // X[M-k]
            const int s1 = xs2(k), s2 = xs2(M - k);
            X0[s1] = make_float2(xr.x, xi.x);
            X1[s1] = make_float2(xr.y, xi.y);
            X0[s2] = make_float2(yr.x, yi.x);
            X1[s2] = make_float2(yr.y, yi.y);
        }
    }
    pair_barrier(bar);

    // ---- middle phases: warp w owns channel w -------------------------------------------------------------
    {
        float2 *Xc = w ? X1 : X0;
        uint32_t *dsc = reinterpret_cast<uint32_t *>(mine + G::OFF_TAB + w * G::TAB_CH_BYTES);
        uint32_t *sw = dsc + G::MAXPK;
        const uint32_t le_mask = (2u << lane) - 1u;
        const bool contract = p.pitch_factor < 1.0f;

        // 5-point strict maxima (pv:95-116) on this lane's run of 16 bins; |X|^2 in float32 (pv:88)
        uint32_t mask = 0;
        {
            int m[24];
#pragma unroll
            for (int e = 0; e < 24; e++) {
                int k = 16 * lane - 4 + e;
                k = k < 0 ? 0 : (k > M ? M : k);
                const float2 v = Xc[xs2(k)];
                m[e] = __float_as_int(fmaf(v.x, v.x, v.y * v.y));
            }
            int q[22];
#pragma unroll
            for (int t = 2; t < 21; t++) q[t] = max(m[t], m[t + 1]);
#pragma unroll
            for (int e = 15; e >= 0; e--) {
                const int nb_max = max(q[e + 2], q[e + 5]);
                mask = __funnelshift_l(uint32_t(nb_max - m[e + 4]), mask, 1);
            }
        }
        if (lane == 0) mask &= ~3u;            // i >= 2
        if (lane == 31) mask &= ~(1u << 15);   // i <= nb - 3 == 510

        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        const int npk = __shfl_sync(FULL, incl, 31);
        const int own_last = mask ? (16 * lane + 31 - __clz(mask)) : -1;
        const uint32_t nz_below = __ballot_sync(FULL, mask != 0) & (le_mask >> 1);
        const int src = nz_below ? (31 - __clz(nz_below)) : 0;
        int prev = __shfl_sync(FULL, own_last, src);
        if (!nz_below) prev = -1;

        if (lane < G::SWORDS) sw[lane] = 0;
        __syncwarp();
        {
            const long long pf_m = p.pf_mant;
            const int pf_s = p.pf_shift;
            const long long pf_half = 1ll << (pf_s - 1);
            const int stepm = p.step_mod_r;
            const int rmask = p.overlaps - 1;
            int ord = incl - cnt;
            uint32_t mm = mask;
            while (mm) {
                const int bit = __ffs(mm) - 1;
                mm &= mm - 1;
                const int pk = 16 * lane + bit;
                const int start = (prev < 0) ? 0 : pk - ((pk - prev) >> 1);
                const long long psl = (pf_m * pk + pf_half) >> pf_s;       // Math.round(p * pitchFactor)
                const bool valid = psl <= NB;                               // pv:127
                const int delta = valid ? int(psl) - pk : 0x4000;
                const int ri = (delta * stepm) & rmask;
                dsc[ord] = (uint32_t(delta) << 16) | (uint32_t(ri) << 10) | uint32_t(pk);
                atomicOr(&sw[start >> 5], 1u << (start & 31));
                prev = pk;
                ord++;
            }
        }
        __syncwarp();
        const uint32_t s_reg = (lane < G::SWORDS) ? sw[lane] : 0u;
        int ps_reg;
        {
            const int c = __popc(s_reg);
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += o;
            }
            ps_reg = inc - c - 1;
        }

        if (npk == 0) {
            for (int i = lane; i < G::XCH; i += 32) Xc[i] = make_float2(0.f, 0.f);      // pv:121
        } else {
            float rot_c, rot_s;
            {
                const float2 t = __ldg(&tw[(lane & (p.overlaps - 1)) * (N / p.overlaps)]);
                rot_c = t.x;
                rot_s = -t.y;
            }
            float2 ext[5];
            if (contract) {
#pragma unroll
                for (int t = 0; t < 5; t++) {
                    const int q = 32 * t + lane;                 // bin 512 + q
                    float2 v = make_float2(0.f, 0.f);
                    if (q == 0) v = Xc[xs2(M)];
                    else if (q <= N / 8) v = stale_level1_x2(Xc, q, tw);
                    ext[t] = v;
                }
            } else {
#pragma unroll
                for (int t = 0; t < 5; t++) ext[t] = make_float2(0.f, 0.f);
                if (lane == 0) ext[0] = Xc[xs2(M)];
            }
            __syncwarp();
            if (lane == 0) Xc[xs2(M)] = make_float2(0.f, 0.f);
            const bool quarter = p.overlaps == 4;

#define PVB_SHIFT_FIRST(DV, BIN, V)                                                               \
            {                                                                                      \
                const int d = (BIN) + (int(DV) >> 16);                                             \
                const bool right = (BIN) >= int((DV) & 1023);                                      \
                const bool okd = unsigned(d) < unsigned(NB);                                       \
                const int ri = ((DV) >> 10) & 31;                                                  \
                float2 y;                                                                          \
                if (quarter) {                                                                     \
                    const float ax = (ri & 1) ? -(V).y : (V).x, ay = (ri & 1) ? (V).x : (V).y;      \
                    y = make_float2((ri & 2) ? -ax : ax, (ri & 2) ? -ay : ay);                      \
                } else {                                                                           \
                    const float rc = __shfl_sync(FULL, rot_c, ri), rs = __shfl_sync(FULL, rot_s, ri); \
                    y = make_float2((V).x * rc - (V).y * rs, (V).x * rs + (V).y * rc);              \
                }                                                                                  \
                const int slot = xs2(d);                                                           \
                (V) = y;                                                                           \
                if (okd && (right || !contract)) Xc[slot] = y;                                     \
                (DV) = (okd && !right && contract) ? uint32_t(slot) : 0xFFFFFFFFu;                 \
            }
#define PVB_SHIFT_SECOND(DV, V)                                                                   \
            if ((DV) != 0xFFFFFFFFu) { float2 o = Xc[(DV)]; o.x += (V).x; o.y += (V).y; Xc[(DV)] = o; }

#pragma unroll 1
            for (int it = 0; it < 3; it++) {
                const int c = contract ? it : 2 - it;
                if (c < 2) {
                    uint32_t dvs[8];
                    float2 xv[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int s = 8 * c + i;
                        const uint32_t sw_w = __shfl_sync(FULL, s_reg, s);
                        dvs[i] = dsc[__shfl_sync(FULL, ps_reg, s) + __popc(sw_w & le_mask)];
                        float2 *xp = Xc + xs2(256 * c + 32 * i + lane);
                        xv[i] = *xp;
                        *xp = make_float2(0.f, 0.f);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; i++) PVB_SHIFT_FIRST(dvs[i], 256 * c + 32 * i + lane, xv[i])
                    if (contract) {
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; i++) PVB_SHIFT_SECOND(dvs[i], xv[i])
                    }
                } else {
                    uint32_t dvs[5];
#pragma unroll
                    for (int t = 0; t < 5; t++) {
                        const int s = 16 + t;
                        const uint32_t sw_w = __shfl_sync(FULL, s_reg, s);
                        dvs[t] = dsc[__shfl_sync(FULL, ps_reg, s) + __popc(sw_w & le_mask)];
                    }
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < 5; t++) PVB_SHIFT_FIRST(dvs[t], 512 + 32 * t + lane, ext[t])
                    if (contract) {
                        __syncwarp();
#pragma unroll
                        for (int t = 0; t < 5; t++) PVB_SHIFT_SECOND(dvs[t], ext[t])
                    }
                }
                __syncwarp();
            }
#undef PVB_SHIFT_FIRST
#undef PVB_SHIFT_SECOND
        }
    }
    pair_barrier(bar);                                             // both channels of Y are complete

    // ---- Hermitian C2R pre-pass: mirror of the split, through registers -------------------------------------
    {
        cpx2 zk[5], zmk[5];
#pragma unroll
        for (int it = 0; it < 5; it++) {
            if (it == 4 && L != 63) break;
            const int k = (it == 4) ? 256 : kq[it];
            const int s1 = xs2(k), s2 = xs2(M - k);
            float2 a0 = X0[s1], a1 = X1[s1], b0 = X0[s2], b1 = X1[s2];
            if (k == 0) { a0.y = 0.f; a1.y = 0.f; b0.y = 0.f; b1.y = 0.f; }
            const float2 ar = make_float2(a0.x, a1.x), ai = make_float2(a0.y, a1.y);
            const float2 br = make_float2(b0.x, b1.x), bi = make_float2(b0.y, b1.y);
            const float2 e_r = add2(ar, br), e_i = sub2(ai, bi);
            const float2 d_r = sub2(ar, br), d_i = add2(ai, bi);
            const float2 t = __ldg(&tw[k]);
            const cpx2 pp = cmul_s(cpx2{d_r, d_i}, t.x, -t.y);
            zk[it] = cpx2{sub2(e_r, pp.im), add2(e_i, pp.re)};
            zmk[it] = cpx2{add2(e_r, pp.im), sub2(pp.re, e_i)};
        }
        pair_barrier(bar);                                         // Y fully read: planes may overwrite it
#pragma unroll
        for (int it = 0; it < 5; it++) {
            if (it == 4 && L != 63) break;
            const int k = (it == 4) ? 256 : kq[it];
            zst(zre, zim, zslot_of_bin(k), zk[it]);
            if (k != 0 && k != 256) zst(zre, zim, zslot_of_bin(M - k), zmk[it]);
        }
    }
    pair_barrier(bar);

    // ---- inverse pass 1 (DIT): butterfly (k1, k2) over k3, twiddle conj(W_64^{k2 m3}) ----------------------
    {
        cpx2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = zld(zre, zim, base3 ^ j);
        dft8<true>(x);
#pragma unroll
        for (int m3 = 1; m3 < 8; m3++) {
            const float2 t = __ldg(&tw[16 * lo * m3]);
            x[m3] = cmul_s(x[m3], t.x, -t.y);
        }
#pragma unroll
        for (int m3 = 0; m3 < 8; m3++) zst(zre, zim, base3 ^ m3, x[m3]);
    }
    pair_barrier(bar);

    // ---- inverse pass 2: butterfly (k1, m3) over k2, twiddle conj(W_512^{k1 (m3 + 8 m2)}) ------------------
    {
        cpx2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = zld(zre, zim, base2 ^ (9 * j));
        dft8<true>(x);
#pragma unroll
        for (int m2 = 0; m2 < 8; m2++) {
            const float2 t = __ldg(&tw[2 * hi * (lo + 8 * m2)]);
            zst(zre, zim, base2 ^ (9 * m2), cmul_s(x[m2], t.x, -t.y));
        }
    }
    pair_barrier(bar);

    // ---- inverse pass 3: butterfly n = L over k1 -> z[n + 64 m1]; window, overlap-add, emit ------------------
    {
        char *accb = reinterpret_cast<char *>(p.acc + size_t(c0) * N);
        char *outb0 = reinterpret_cast<char *>(p.out + size_t(c0) * hop);
        char *outb1 = outb0 + (has1 ? hop * 4 : 0);
        const unsigned abase = unsigned((2 * L + rb) & (N - 1)) * 4u;
        cpx2 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = zld(zre, zim, 64 * j + ((j & 1) ? base1o : base1e));
        dft8<true>(x);
#pragma unroll
        for (int half = 0; half < 2; half++) {
            float2 q0[4], q1[4], wo[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int m1 = 4 * half + i;
                const int s = 2 * L + 128 * m1;
                const unsigned off = (abase + 512u * m1) & (N * 4 - 1);
                wo[i] = __ldg(reinterpret_cast<const float2 *>(wp.window_out + s));
                q0[i] = make_float2(0.f, 0.f);
                q1[i] = make_float2(0.f, 0.f);
                if (s < keep) {                                   // the tail slot starts from zero (ola:134)
                    q0[i] = *reinterpret_cast<const float2 *>(accb + off);
                    q1[i] = *reinterpret_cast<const float2 *>(accb + off + N * 4);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int m1 = 4 * half + i;
                const int s = 2 * L + 128 * m1;
                const float2 yr = mul2(x[m1].re, bc2(wo[i].x));   // sample s   of (ch0, ch1)
                const float2 yi = mul2(x[m1].im, bc2(wo[i].y));   // sample s+1 of (ch0, ch1)
                const float2 y0 = make_float2(yr.x + q0[i].x, yi.x + q0[i].y);
                const float2 y1 = make_float2(yr.y + q1[i].x, yi.y + q1[i].y);
                if (s < hop) {                                    // head: emit (ola:111-118)
                    *reinterpret_cast<float2 *>(outb0 + 4 * s) = y0;
                    if (has1) *reinterpret_cast<float2 *>(outb1 + 4 * s) = y1;
                } else {
                    const unsigned off = (abase + 512u * m1) & (N * 4 - 1);
                    *reinterpret_cast<float2 *>(accb + off) = y0;
                    *reinterpret_cast<float2 *>(accb + off + N * 4) = y1;
                }
            }
        }
    }
}

}  // namespace pvb
