// pv_kernel_warp.cuh — warp-synchronous fused kernel for frame size 1024 (sm_100a).
//
// Same arithmetic as pv_kernel.cuh (one launch == one process() call; the reference lines
// are cited there), laid out so that ONE WARP owns one channel pair end to end:
//
//   * no __syncthreads(): every exchange is shared memory + __syncwarp(), so warps drift
//     apart and the scheduler overlaps FFT math of one pair with the integer-heavy peak /
//     shift phases of another;
//   * both FFTs run global -> registers -> 2 shared-memory exchanges -> registers: first-pass
//     inputs come straight from the history ring, last-pass outputs of the inverse go straight
//     to the overlap-add ring; the two channels ride in the halves of f32x2 registers;
//   * the last forward pass gives every lane both Z[k] and Z[M-k], so the real-split (and its
//     mirror, the Hermitian C2R pre-pass) happens in registers;
//   * exchange buffer = separate re / im planes of float2 (ch0, ch1) with an XOR swizzle
//     slot(a,b,c) = 64a + 8(b ^ (a&1)) + (c ^ b): every 64-bit access pattern of the three
//     radix-8 passes is bank-conflict free and needs no register shuffling;
//   * peak picking: 16-bin runs per lane, |X|^2 compared as integers, mask built with funnel
//     shifts; regions get 32-bit descriptors indexed by a prefix-popcount of a region-start
//     bitmap kept in registers; Math.round(p * pitchFactor) is exact integer arithmetic;
//   * the shift is done IN PLACE: the sweep reads X[b], zeroes it, and writes the rotated value
//     to Y[b + delta] in the same buffer (ascending bins when contracting, descending when
//     expanding), so one 8 KB buffer per warp serves Z (forward), X, Y and Z (inverse).
//     13.5 KB of shared memory per warp -> 14-16 resident warps per SM.
//
// Valid for pitch factors in [0.75, 64] and hop >= 32 (R <= 32): then (a) only the first level
// of stale upper bins is read, (b) right halves of consecutive regions, and left halves, stay
// pairwise disjoint after the shift, so two ordered sub-steps replace atomics.  The host routes
// everything else (and other frame sizes) to the generic kernel in pv_kernel.cuh.
#pragma once

#include "pv_kernel.cuh"

namespace pvb {

struct WarpGeo {
    static constexpr int N = 1024, M = 512, NB = 513;
    static constexpr int PLANE = 512;                      // float2 slots per plane
    static constexpr int XCH = 516;                        // float2 slots per channel of X / Y
    static constexpr int BUF_BYTES = 2 * XCH * 8;          // 8256 >= 2 planes (8192)
    static constexpr int MAG_FLOATS = 656;                 // per channel, padded layout (see mag_idx)
    static constexpr int MAG_BYTES = 2 * MAG_FLOATS * 4;   // 5248
    static constexpr int MAXPK = 176;                      // > 509 / 3 peaks
    static constexpr int SWORDS = 20;                      // region-start bitmap words (bins 0..639)
    static constexpr int OFF_BUF = 0;
    static constexpr int OFF_MAG = BUF_BYTES;
    static constexpr int WARP_BYTES = BUF_BYTES + MAG_BYTES;     // 13504
    // descriptors / bitmap of channel c live in the (dead) mag space of channel c
    static constexpr int DESC_OFF = 0;                     // u32[MAXPK]
    static constexpr int S_OFF = MAXPK * 4;                // u32[SWORDS]
    static constexpr int MAX_WARPS = 7;          // 2 CTAs x 7 warps per SM leave 144 registers per thread
};

struct WarpParams {
    FrameParams f;            // pf_shift in [1, 62], 0.75 <= pitch_factor <= 64, overlaps <= 32
    const float *window_out;  // [N] window * 1 / (2 N R): synthesis window with every scale folded in
    int stagger_ns;           // start-time offset between the warps that share an SM (0: none)
    int num_sms;
};

// swizzled slot of exchange element [a][b][c] (each 0..7)
__device__ __forceinline__ int zslot(int a, int b, int c) { return 64 * a + 8 * (b ^ (a & 1)) + (c ^ b); }

// padded index into the per-channel |X|^2 array: 4 floats of padding after every 16 bins and
// 8 in front, so a lane's run of 24 floats is 6 aligned, conflict-free 16-byte loads
__device__ __forceinline__ int mag_idx(int i) { return i + 4 * (i >> 4) + 8; }

// swizzled float2 slot of spectrum bin k (a bijection inside every aligned block of 64)
__device__ __forceinline__ int xs(int k) { return k ^ ((k >> 3) & 6); }

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
    const int s1 = xs(k), s2 = xs(k2);
    X0[s1] = make_float2(xr.x, xi.x);
    X1[s1] = make_float2(xr.y, xi.y);
    X0[s2] = make_float2(yr.x, yi.x);
    X1[s2] = make_float2(yr.y, yi.y);
    const float2 mk = fma2(xr, xr, mul2(xi, xi));
    const float2 mk2 = fma2(yr, yr, mul2(yi, yi));
    const int m1 = mag_idx(k), m2 = mag_idx(k2);
    mag0[m1] = mk.x;
    mag1[m1] = mk.y;
    mag0[m2] = mk2.x;
    mag1[m2] = mk2.y;
}

// Hermitian C2R pre-pass of one (k, M-k) pair: Y -> Z'[k], Z'[M-k]
__device__ __forceinline__ void unsplit_load(int k, const float2 *__restrict__ tw, const float2 *Y0,
                                             const float2 *Y1, cpx2 &zk, cpx2 &zmk) {
    constexpr int M = WarpGeo::M;
    const int s1 = xs(k), s2 = xs(M - k);
    float2 a0 = Y0[s1], a1 = Y1[s1], b0 = Y0[s2], b1 = Y1[s2];
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

// store / load one packed complex (both channels) of the exchange planes
__device__ __forceinline__ void zst(float2 *zre, float2 *zim, int slot, cpx2 v) {
    zre[slot] = v.re;
    zim[slot] = v.im;
}
__device__ __forceinline__ cpx2 zld(const float2 *zre, const float2 *zim, int slot) {
    return cpx2{zre[slot], zim[slot]};
}

// ALIGNED: hop is a multiple of 128.  Frame element j of a lane then sits in ring block
// (j + rot) & 7 with a launch-uniform rot, so the rings are addressed as lane base + immediate,
// old-vs-new-block and head / tail decisions are uniform branches, and the rotation between frame
// order and ring order is folded into twiddle indices (shift theorem of the 8-point DFT).
template <bool ALIGNED>
__global__ void __launch_bounds__(WarpGeo::MAX_WARPS * 32, 2)
pv_process_warp_kernel(const WarpParams wp) {
    using W = WarpGeo;
    constexpr int N = W::N, M = W::M, NB = W::NB;
    const FrameParams &p = wp.f;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int pair = blockIdx.x * (blockDim.x >> 5) + warp;
    if (2 * pair >= p.num_channels) return;          // whole warp leaves; no CTA-wide barriers below
    // A launch that fits in one wave starts every warp in the same phase; they would then fight
    // for the same pipe (FMA in the FFT passes, ALU in the shift, LSU in the exchanges) all the
    // way through.  Spreading the start times lets the phases of different warps interleave.
    if (wp.stagger_ns > 0) {
        const int slot = 2 * warp + (int(blockIdx.x) >= wp.num_sms ? 1 : 0);
        if (slot > 0) __nanosleep(unsigned(slot * wp.stagger_ns));
    }
    unsigned char *mine = smem_raw + size_t(warp) * W::WARP_BYTES;
    float2 *zre = reinterpret_cast<float2 *>(mine + W::OFF_BUF);
    float2 *zim = zre + W::PLANE;
    float2 *X0 = reinterpret_cast<float2 *>(mine + W::OFF_BUF);      // X and (in place) Y, channel 0
    float2 *X1 = X0 + W::XCH;
    float *mag0 = reinterpret_cast<float *>(mine + W::OFF_MAG);
    float *mag1 = mag0 + W::MAG_FLOATS;

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
    // pass-3 slots: zslot(k1, k2, j) == base ^ j
    const int slotA = 64 * k1a + 8 * (k2a ^ (k1a & 1)) + k2a;
    const int slotB = 64 * k1b + 8 * (k2b ^ (k1b & 1)) + k2b;
    // pass-2 slots: butterflies (k1, m3) = (lane >> 3 [+4], lane & 7): zslot(k1, j, m3) == base2 ^ (9 j)
    const int base2 = 64 * (lane >> 3) + 8 * ((lane >> 3) & 1) + (lane & 7);
    // pass-1 slots: n = lane [+32] = (m2, m3): zslot(k1, m2, m3) == 64 k1 + (k1 odd ? base1o : base1e)
    const int m2l = lane >> 3, m3l = lane & 7;
    const int base1e = 8 * m2l + (m3l ^ m2l);
    const int base1o = 8 * (m2l ^ 1) + (m3l ^ m2l);

    cpx2 a[8], b[8];

    // ---- forward pass 1: butterflies n = lane, lane + 32 over m1 (stride 64), from global ----
    if constexpr (ALIGNED) {
        const int rot = ((rb + hop) >> 7) & 7;         // ring block of frame element j: (j + rot) & 7
        const int jk = keep >> 7;                      // frame elements j >= jk are the new block
        char *hl = reinterpret_cast<char *>(p.hist + size_t(c0) * N) + 8 * lane;
        const char *il0 = reinterpret_cast<const char *>(p.in ? p.in + size_t(c0) * hop : p.hist) + 8 * lane;
        const char *il1 = il0 + (has1 && p.in ? hop * 4 : 0);
        const char *wl = reinterpret_cast<const char *>(p.window) + 8 * lane;
        const bool paused = p.in == nullptr;
        float2 r0[16], r1[16];
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int h = e >> 3, b = e & 7;
            const int j = (b - rot) & 7;               // uniform
            if (j < jk) {
                r0[e] = *reinterpret_cast<const float2 *>(hl + 256 * h + 512 * b);
                r1[e] = *reinterpret_cast<const float2 *>(hl + 256 * h + 512 * b + N * 4);
            } else {
                r0[e] = *reinterpret_cast<const float2 *>(il0 + 256 * h + 512 * (j - jk));
                r1[e] = *reinterpret_cast<const float2 *>(il1 + 256 * h + 512 * (j - jk));
            }
        }
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int h = e >> 3, b = e & 7;
            const int j = (b - rot) & 7;
            if (j >= jk) {                             // the new block: paused input is zeros (ola:93-100)
                if (paused) { r0[e] = make_float2(0.f, 0.f); r1[e] = make_float2(0.f, 0.f); }
                if (!has1) r1[e] = make_float2(0.f, 0.f);
                *reinterpret_cast<float2 *>(hl + 256 * h + 512 * b) = r0[e];
                *reinterpret_cast<float2 *>(hl + 256 * h + 512 * b + N * 4) = r1[e];
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int n = lane + 32 * h;
            cpx2 x[8];
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const int j = (b - rot) & 7;
                const float2 w = __ldg(reinterpret_cast<const float2 *>(wl + 256 * h + 512 * j));
                x[b].re = mul2(make_float2(r0[8 * h + b].x, r1[8 * h + b].x), bc2(w.x));
                x[b].im = mul2(make_float2(r0[8 * h + b].y, r1[8 * h + b].y), bc2(w.y));
            }
            dft8<false>(x);                            // inputs in ring order: outputs carry W8^{rot k1}
            const int tb = 2 * n - 128 * rot;
#pragma unroll
            for (int k1 = 1; k1 < 8; k1++) {
                const float2 w = __ldg(&tw[(tb * k1) & (N - 1)]);    // W_512^{n k1} * W8^{-rot k1}
                x[k1] = cmul_s(x[k1], w.x, w.y);
            }
#pragma unroll
            for (int k1 = 0; k1 < 8; k1++)
                zst(zre, zim, 64 * k1 + 32 * h + (((k1 & 1) ? base1o : base1e) ^ (4 * h)), x[k1]);
        }
    } else {
        // all 32 loads of the frame are issued before anything consumes them.  Rows of hist /
        // acc are padded to an even channel count by the host, so channel 1 is always +N floats.
        float2 r0[16], r1[16];
        char *histb = reinterpret_cast<char *>(p.hist + size_t(c0) * N);
        const char *inb0 = reinterpret_cast<const char *>(p.in ? p.in + size_t(c0) * hop : p.hist);
        const char *inb1 = inb0 + (has1 && p.in ? hop * 4 : 0);
        const unsigned rbase = unsigned((2 * lane + rb + hop) & (N - 1)) * 4u;   // ring byte offset of sample 2*lane
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int sd = 64 * (e >> 3) + 128 * (e & 7);         // sample index minus 2*lane
            if (2 * lane + sd < keep) {
                const unsigned off = (rbase + 4u * sd) & (N * 4 - 1);
                r0[e] = *reinterpret_cast<const float2 *>(histb + off);
                r1[e] = *reinterpret_cast<const float2 *>(histb + off + N * 4);
            } else {
                const int ib = (2 * lane + sd - keep) * 4;
                r0[e] = *reinterpret_cast<const float2 *>(inb0 + ib);
                r1[e] = *reinterpret_cast<const float2 *>(inb1 + ib);
            }
        }
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int sd = 64 * (e >> 3) + 128 * (e & 7);
            if (2 * lane + sd >= keep) {           // the new block: paused input is zeros (ola:93-100)
                if (!p.in) { r0[e] = make_float2(0.f, 0.f); r1[e] = make_float2(0.f, 0.f); }
                if (!has1) r1[e] = make_float2(0.f, 0.f);
                const unsigned off = unsigned(rb + 2 * lane + sd - keep) * 4u;
                *reinterpret_cast<float2 *>(histb + off) = r0[e];
                *reinterpret_cast<float2 *>(histb + off + N * 4) = r1[e];
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int n = lane + 32 * h;
            cpx2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int s = 2 * n + 128 * j;
                const float2 w = __ldg(reinterpret_cast<const float2 *>(p.window + s));
                x[j].re = mul2(make_float2(r0[8 * h + j].x, r1[8 * h + j].x), bc2(w.x));
                x[j].im = mul2(make_float2(r0[8 * h + j].y, r1[8 * h + j].y), bc2(w.y));
            }
            dft8<false>(x);
#pragma unroll
            for (int k1 = 1; k1 < 8; k1++) {
                const float2 w = __ldg(&tw[2 * n * k1]);             // W_512^{n k1}
                x[k1] = cmul_s(x[k1], w.x, w.y);
            }
#pragma unroll
            for (int k1 = 0; k1 < 8; k1++)
                zst(zre, zim, 64 * k1 + 32 * h + (((k1 & 1) ? base1o : base1e) ^ (4 * h)), x[k1]);
        }
    }
    __syncwarp();

    // (issued only now, after the frame has arrived: at kernel start it would compete with the
    // frame loads for DRAM bandwidth)
    // warm L2 with the overlap-add ring lines the tail of this kernel adds to (their loads
    // would otherwise be a serial DRAM round trip at the very end); the slot that is only
    // written ([rb - hop, rb)) is skipped
    {
        const int line = 32 * lane;                                     // floats [32 lane, 32 lane + 32)
        if (((line - rb + hop) & (N - 1)) >= hop) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.acc + size_t(c0) * N + line));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.acc + size_t(c1) * N + line));
        }
    }

    // ---- forward pass 2: butterflies (k1, m3) over m2 ---------------------------------------
    {
        int s2[8];
#pragma unroll
        for (int j = 0; j < 8; j++) s2[j] = base2 ^ (9 * j);
        float2 w2[8];
#pragma unroll
        for (int k2 = 1; k2 < 8; k2++) w2[k2] = __ldg(&tw[16 * m3l * k2]);   // W_64^{m3 k2}
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            cpx2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = zld(zre, zim, s2[j] + 256 * h);
            dft8<false>(x);
#pragma unroll
            for (int k2 = 1; k2 < 8; k2++) x[k2] = cmul_s(x[k2], w2[k2].x, w2[k2].y);
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) zst(zre, zim, s2[k2] + 256 * h, x[k2]);
        }
    }
    __syncwarp();

    // ---- forward pass 3: butterflies A and B over m3; outputs stay in registers --------------
#pragma unroll
    for (int j = 0; j < 8; j++) {
        a[j] = zld(zre, zim, slotA ^ j);
        b[j] = zld(zre, zim, slotB ^ j);
    }
    dft8<false>(a);      // a[j] = Z[kA + 64 j]
    dft8<false>(b);      // b[j] = Z[kB + 64 j]
    __syncwarp();        // everyone has read the planes: X may overwrite them

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

    // ---- per channel: peaks -> region descriptors -> in-place shift --------------------------------
    // rotation exp(+j 2 pi r / R) for r = lane (pv:155-157 with t = hop * calls, integer reduced)
    float rot_c, rot_s;
    {
        const float2 w = __ldg(&tw[(lane & (p.overlaps - 1)) * (N / p.overlaps)]);
        rot_c = w.x;
        rot_s = -w.y;
    }
    const long long pf_m = p.pf_mant;
    const int pf_s = p.pf_shift;
    const long long pf_half = 1ll << (pf_s - 1);
    const int stepm = p.step_mod_r;
    const int rmask = p.overlaps - 1;
    const bool contract = p.pitch_factor < 1.0f;
    const uint32_t le_mask = (2u << lane) - 1u;
    const int sw_e = lane ^ ((lane >> 3) & 6);              // xs(32 s + lane) - 32 s, s even
    const int sw_o = lane ^ (((lane >> 3) & 6) ^ 4);        // s odd

#pragma unroll 1
    for (int ch = 0; ch < 2; ch++) {
        float *mg = ch ? mag1 : mag0;
        float2 *Xc = ch ? X1 : X0;
        uint32_t *dsc = reinterpret_cast<uint32_t *>(mg) + W::DESC_OFF / 4;
        uint32_t *sw = reinterpret_cast<uint32_t *>(mg) + W::S_OFF / 4;

        // 5-point strict maxima (pv:95-116) on this lane's run of 16 bins; squared magnitudes are
        // non-negative floats, so they order like their bit patterns
        uint32_t mask = 0;
        {
            int m[24];
            const int4 *mv = reinterpret_cast<const int4 *>(mg + 20 * lane);
            const int4 q0 = mv[0], q1 = mv[2], q2 = mv[3], q3 = mv[4], q4 = mv[5], q5 = mv[7];
            m[0] = q0.x; m[1] = q0.y; m[2] = q0.z; m[3] = q0.w;
            m[4] = q1.x; m[5] = q1.y; m[6] = q1.z; m[7] = q1.w;
            m[8] = q2.x; m[9] = q2.y; m[10] = q2.z; m[11] = q2.w;
            m[12] = q3.x; m[13] = q3.y; m[14] = q3.z; m[15] = q3.w;
            m[16] = q4.x; m[17] = q4.y; m[18] = q4.z; m[19] = q4.w;
            m[20] = q5.x; m[21] = q5.y; m[22] = q5.z; m[23] = q5.w;
            int q[22];
#pragma unroll
            for (int t = 2; t < 21; t++) q[t] = max(m[t], m[t + 1]);
#pragma unroll
            for (int e = 15; e >= 0; e--) {
                const int nb_max = max(q[e + 2], q[e + 5]);        // bins e-2, e-1, e+1, e+2
                mask = __funnelshift_l(uint32_t(nb_max - m[e + 4]), mask, 1);   // 1 iff m[e] > all four
            }
        }
        if (lane == 0) mask &= ~3u;            // i >= 2
        if (lane == 31) mask &= ~(1u << 15);   // i <= nb - 3 == 510
        __syncwarp();                          // all lanes have read mag: descriptors may overwrite it

        // ordinals of this lane's peaks, position of the last peak before this lane's run
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

        if (lane < W::SWORDS) sw[lane] = 0;
        __syncwarp();
        // one descriptor per region of influence (pv:124-141):  delta << 16 | rot index << 10 | peak
        {
            int ord = incl - cnt;
            uint32_t mm = mask;
            while (mm) {
                const int bit = __ffs(mm) - 1;
                mm &= mm - 1;
                const int pk = 16 * lane + bit;
                const int start = (prev < 0) ? 0 : pk - ((pk - prev) >> 1);
                const long long psl = (pf_m * pk + pf_half) >> pf_s;       // Math.round(p * pitchFactor)
                const bool valid = psl <= NB;                               // pv:127
                const int delta = valid ? int(psl) - pk : 0x4000;           // 0x4000: lands outside [0, nb)
                const int ri = (delta * stepm) & rmask;
                dsc[ord] = (uint32_t(delta) << 16) | (uint32_t(ri) << 10) | uint32_t(pk);
                atomicOr(&sw[start >> 5], 1u << (start & 31));
                prev = pk;
                ord++;
            }
        }
        __syncwarp();
        // lane w keeps word w of the region-start bitmap and (#starts in words < w) - 1
        uint32_t s_reg = (lane < W::SWORDS) ? sw[lane] : 0u;
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
            // no peaks (silence): the shifted spectrum is all zero (pv:121)
            for (int i = lane; i < W::XCH; i += 32) Xc[i] = make_float2(0.f, 0.f);
            continue;
        }

        // ---- in-place shift -------------------------------------------------------------------
        // Source bins come in three chunks: bins 0..255, 256..511 (8 steps of 32 each) and the
        // extension 512..671 (X[512] and the stale slots; 5 steps).  Per chunk:
        //  A  region descriptor of every source bin this lane owns (bins 32 s + lane)
        //  B  read those bins of X into registers, zero them in shared memory
        //  C  right halves of the regions: plain stores (pairwise disjoint destinations)
        //  D  left halves: read-add-store on top (pairwise disjoint among themselves)
        // Contraction writes at or below the bin it read (chunks ascending), expansion at or
        // above (chunks descending, and it never maps two sources to one bin: no pass D).
        float2 ext[5];
        if (contract) {
#pragma unroll
            for (int t = 0; t < 5; t++) {
                const int q = 32 * t + lane;                 // bin 512 + q
                float2 v = make_float2(0.f, 0.f);
                if (q == 0) v = Xc[xs(M)];
                else if (q <= N / 8) v = stale_level1(Xc, q, tw);
                ext[t] = v;
            }
        } else {
#pragma unroll
            for (int t = 0; t < 5; t++) ext[t] = make_float2(0.f, 0.f);
            if (lane == 0) ext[0] = Xc[xs(M)];
        }
        __syncwarp();            // the stale slots were computed from an intact X
        if (lane == 0) Xc[xs(M)] = make_float2(0.f, 0.f);

        const bool quarter = p.overlaps == 4;    // R == 4: rotations are exact quarter turns
        // Pass C: rotate the value V of source bin BIN, store it at BIN + delta if this lane is the
        // first writer of that bin (right half of its region, or any bin when expanding); otherwise
        // keep the rotated value in V and the destination slot in DV for pass D.
#define PVB_SHIFT_FIRST(DV, BIN, V)                                                               \
        {                                                                                          \
            const int d = (BIN) + (int(DV) >> 16);                                                 \
            const bool right = (BIN) >= int((DV) & 1023);                                          \
            const bool okd = unsigned(d) < unsigned(NB);                                           \
            const int ri = ((DV) >> 10) & 31;                                                      \
            float2 y;                                                                              \
            if (quarter) {                                                                         \
                const float ax = (ri & 1) ? -(V).y : (V).x, ay = (ri & 1) ? (V).x : (V).y;          \
                y = make_float2((ri & 2) ? -ax : ax, (ri & 2) ? -ay : ay);                          \
            } else {                                                                               \
                const float rc = __shfl_sync(FULL, rot_c, ri), rs = __shfl_sync(FULL, rot_s, ri);  \
                y = make_float2((V).x * rc - (V).y * rs, (V).x * rs + (V).y * rc);                  \
            }                                                                                      \
            const int slot = xs(d);                                                                \
            (V) = y;                                                                               \
            if (okd && (right || !contract)) Xc[slot] = y;                                         \
            (DV) = (okd && !right && contract) ? uint32_t(slot) : 0xFFFFFFFFu;                     \
        }
        // Pass D: left halves add on top of whatever pass C left in their bin
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
                    float2 *xp = Xc + 256 * c + 32 * i + ((i & 1) ? sw_o : sw_e);
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
    __syncwarp();

    // ---- Hermitian C2R pre-pass in registers (mirror of the split) -------------------------------------
    {
        cpx2 zk[8], zmk[8], z256;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = (l31 && j >= 4) ? 64 * (7 - j) : kA + 64 * j;
            unsplit_load(k, tw, X0, X1, zk[j], zmk[j]);
        }
        z256 = zk[0];
        if (l31) { cpx2 dummy; unsplit_load(256, tw, X0, X1, z256, dummy); }
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
    __syncwarp();        // everyone has read Y: the inverse exchange planes may overwrite it

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
        zst(zre, zim, slotA ^ m3, va);
        zst(zre, zim, slotB ^ m3, vb);
    }
    __syncwarp();

    // ---- inverse pass 2: butterflies (k1, m3) over k2, twiddle conj(W_512^{k1 (m3 + 8 m2)}) ---------------
    {
        int s2[8];
#pragma unroll
        for (int j = 0; j < 8; j++) s2[j] = base2 ^ (9 * j);
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int k1 = (lane >> 3) + 4 * h;
            cpx2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = zld(zre, zim, s2[j] + 256 * h);
            dft8<true>(x);
            // ALIGNED: the last pass must deliver its outputs in ring order, i.e. rotated by rotA
            // blocks: pre-multiply its input k1 by W8^{k1 rotA}
            const int rotk = ALIGNED ? 128 * k1 * ((rb >> 7) & 7) : 0;
#pragma unroll
            for (int m2 = 0; m2 < 8; m2++) {
                const float2 w = __ldg(&tw[(2 * k1 * (m3l + 8 * m2) - rotk) & (N - 1)]);
                zst(zre, zim, s2[m2] + 256 * h, cmul_s(x[m2], w.x, -w.y));
            }
        }
    }
    __syncwarp();

    // ---- inverse pass 3: butterflies n over k1 -> z[n + 64 m1]; window, overlap-add, emit ------------------
    if constexpr (ALIGNED) {
        const int rotA = (rb >> 7) & 7;                // ring block of output m1: (m1 + rotA) & 7
        const int jk = keep >> 7, hq = hop >> 7;
        char *al = reinterpret_cast<char *>(p.acc + size_t(c0) * N) + 8 * lane;
        char *ol0 = reinterpret_cast<char *>(p.out + size_t(c0) * hop) + 8 * lane;
        char *ol1 = ol0 + (has1 ? hop * 4 : 0);
        const char *wol = reinterpret_cast<const char *>(wp.window_out) + 8 * lane;
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            float2 q0[8], q1[8], wo[8];
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const int j = (b - rotA) & 7;          // frame element held by ring block b (uniform)
                wo[b] = __ldg(reinterpret_cast<const float2 *>(wol + 256 * h + 512 * j));
                q0[b] = make_float2(0.f, 0.f);
                q1[b] = make_float2(0.f, 0.f);
                if (j < jk) {                          // the tail slot starts from zero (ola:134)
                    q0[b] = *reinterpret_cast<const float2 *>(al + 256 * h + 512 * b);
                    q1[b] = *reinterpret_cast<const float2 *>(al + 256 * h + 512 * b + N * 4);
                }
            }
            cpx2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++)
                x[j] = zld(zre, zim, 64 * j + 32 * h + (((j & 1) ? base1o : base1e) ^ (4 * h)));
            dft8<true>(x);                             // x[b]: output in ring block b
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const int j = (b - rotA) & 7;
                const float2 yr = mul2(x[b].re, bc2(wo[b].x));
                const float2 yi = mul2(x[b].im, bc2(wo[b].y));
                const float2 y0 = make_float2(yr.x + q0[b].x, yi.x + q0[b].y);
                const float2 y1 = make_float2(yr.y + q1[b].x, yi.y + q1[b].y);
                if (j < hq) {                          // head: emit (ola:111-118)
                    *reinterpret_cast<float2 *>(ol0 + 256 * h + 512 * j) = y0;
                    if (has1) *reinterpret_cast<float2 *>(ol1 + 256 * h + 512 * j) = y1;
                } else {
                    *reinterpret_cast<float2 *>(al + 256 * h + 512 * b) = y0;
                    *reinterpret_cast<float2 *>(al + 256 * h + 512 * b + N * 4) = y1;
                }
            }
        }
    } else {
        char *accb = reinterpret_cast<char *>(p.acc + size_t(c0) * N);
        char *outb0 = reinterpret_cast<char *>(p.out + size_t(c0) * hop);
        char *outb1 = outb0 + (has1 ? hop * 4 : 0);
        const unsigned abase = unsigned((2 * lane + rb) & (N - 1)) * 4u;       // ring byte offset of sample 2*lane
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int n = lane + 32 * h;
            // accumulator values first (L2 hits thanks to the prefetch), then the butterflies
            float2 q0[8], q1[8], wo[8];
#pragma unroll
            for (int m1 = 0; m1 < 8; m1++) {
                const int s = 2 * n + 128 * m1;
                const unsigned off = (abase + 4u * (64 * h + 128 * m1)) & (N * 4 - 1);
                wo[m1] = __ldg(reinterpret_cast<const float2 *>(wp.window_out + s));
                q0[m1] = make_float2(0.f, 0.f);
                q1[m1] = make_float2(0.f, 0.f);
                if (s < keep) {                                   // the tail slot starts from zero (ola:134)
                    q0[m1] = *reinterpret_cast<const float2 *>(accb + off);
                    q1[m1] = *reinterpret_cast<const float2 *>(accb + off + N * 4);
                }
            }
            cpx2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++)
                x[j] = zld(zre, zim, 64 * j + 32 * h + (((j & 1) ? base1o : base1e) ^ (4 * h)));
            dft8<true>(x);
#pragma unroll
            for (int m1 = 0; m1 < 8; m1++) {
                const int s = 2 * n + 128 * m1;
                // window_out = hannWindow / (2 N R): fromComplexArray, applyHannWindow and the
                // division by nbOverlaps (pv:65-67, ola:153) in one multiply (the scales are powers of 2)
                const float2 w = wo[m1];
                const float2 yr = mul2(x[m1].re, bc2(w.x));       // sample s   of (ch0, ch1)
                const float2 yi = mul2(x[m1].im, bc2(w.y));       // sample s+1 of (ch0, ch1)
                const float2 y0 = make_float2(yr.x + q0[m1].x, yi.x + q0[m1].y);
                const float2 y1 = make_float2(yr.y + q1[m1].x, yi.y + q1[m1].y);
                if (s < hop) {                                    // head: emit (ola:111-118)
                    *reinterpret_cast<float2 *>(outb0 + 4 * s) = y0;
                    if (has1) *reinterpret_cast<float2 *>(outb1 + 4 * s) = y1;
                } else {
                    const unsigned off = (abase + 4u * (64 * h + 128 * m1)) & (N * 4 - 1);
                    *reinterpret_cast<float2 *>(accb + off) = y0;
                    *reinterpret_cast<float2 *>(accb + off + N * 4) = y1;
                }
            }
        }
    }
}

}  // namespace pvb
