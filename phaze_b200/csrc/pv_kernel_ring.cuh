// pv_kernel_ring.cuh — ring-order fused kernel for frame size 1024 (sm_100a), one warp per
// channel pair.  Same reference arithmetic as pv_kernel.cuh (one launch == one process() call,
// ola-processor.js:159-171 + phase-vocoder.js:45-72), reorganised around one identity:
//
//   Both state rings are kept ALIGNED TO THE TIME CURSOR t (frame sample n lives at ring index
//   (n + t) mod N).  The FFT of the ring-ordered windowed frame is U[k] = X[k] e^{-j 2 pi k t / N},
//   so shiftPeaks' rotation e^{j 2 pi (bs - b) t / N} (phase-vocoder.js:155-170) turns into the
//   identity  V[b + delta] += U[b],  and the inverse FFT of V is the output frame in ring order
//   again.  No rotation, no cos/sin, no re-ordering between ring and frame order; only the two
//   window tables are read at a rotated offset (doubled tables, base + immediate).
//
// Further differences from pv_kernel_warp.cuh:
//   * state layout "paired": hist2 / acc2 hold float2 (ch0, ch1) per sample, so 16-byte global
//     accesses deliver packed f32x2 operands for both channels (no pack / unpack moves);
//   * exchange buffer = 16-byte slots (re0, re1, im0, im1) at 65 k1 + 8 r + c: every 128-bit
//     access pattern of the three radix-8 passes is conflict free and is lane base + immediate;
//   * the spectrum is stored as one float4 per bin for both channels (bin k at k + (k >> 4));
//     every lane owns a RUN of 16 consecutive bins: |X|^2 is recomputed from the run (packed),
//     peaks are a 16-bit mask per lane, and the region of influence of every bin (nearest peak,
//     ties to the higher one, pv:132-141) comes from one forward and one backward scan in
//     registers; delta = round(p * pitchFactor) - p is a 513-entry table built once per CTA;
//   * no peak list, no descriptors, no prefix sums, no atomics.
//
// Valid for hop % 128 == 0, hop <= 512 and pitch factors in [0.75, 64] (first stale level only,
// right halves and left halves of regions stay pairwise disjoint after the shift).  tests/
// ring_kernel_model.py restates this file lane by lane in numpy (CPU test of the design).
#pragma once

#include "pv_kernel.cuh"
#include "pv_kernel_warp.cuh"

namespace pvb {

struct RingGeo {
    static constexpr int N = 1024, M = 512, NB = 513;
    static constexpr int EX_SLOTS = 65 * 7 + 64;            // 519 exchange slots of 16 bytes
    static constexpr int XQ_SLOTS = 546;                    // bin k at k + (k >> 4); 545 = halo dummy
    static constexpr int WARP_BYTES = XQ_SLOTS * 16;        // 8736 (>= 519 * 16)
    // CTA-shared tables (bytes), in this order at the start of dynamic shared memory
    static constexpr int DTAB_BYTES = 2592;                 // key table: int32, bin p at p + 4 (p >> 4), p <= 513
    static constexpr int TW1_ROW = 72;                      // float2 per row of tw1 (row stride = 16 banks mod 32)
    static constexpr int TW1_BYTES = 8 * TW1_ROW * 8;       // tw1[k1][n] = W_512^{n k1}, n < 64
    static constexpr int W64_BYTES = 64 * 8;                // w64[a][b] = W_64^{a b}
    static constexpr int TWH_BYTES = 520 * 8;               // twh[k] = W_1024^k, k <= 512
    static constexpr int GTAB_BYTES = TW1_BYTES + W64_BYTES + TWH_BYTES;    // copied verbatim from global
    static constexpr int WIN_BYTES = 1024 * 4;              // window / synthesis window, rotated by t
    static constexpr int OFF_TW1 = DTAB_BYTES;
    static constexpr int OFF_W64 = OFF_TW1 + TW1_BYTES;
    static constexpr int OFF_TWH = OFF_W64 + W64_BYTES;
    static constexpr int OFF_WIN = OFF_TWH + TWH_BYTES;
    static constexpr int OFF_WOUT = OFF_WIN + WIN_BYTES;
    static constexpr int TAB_BYTES = OFF_WOUT + WIN_BYTES;  // 18512
    static constexpr int MAX_WARPS = 7;
    static constexpr int INVALID_DELTA = 0x3000;            // lands outside [0, nb) from any bin
};

struct RingParams {
    const float *in;            // [C][hop] or nullptr (paused input: zeros, ola:93-100)
    float *out;                 // [C][hop]
    float4 *hist2;              // [pairs][N/2] : (ch0[i], ch1[i], ch0[i+1], ch1[i+1]), ring aligned to t
    float4 *acc2;               // [pairs][N/2] : overlap-add ring, same alignment
    const float *window2;       // [2N] Hann window, twice
    const float *window_out2;   // [2N] window / (2 N R), twice
    const float4 *gtab;         // [GTAB_BYTES / 16] tw1 | w64 | twh, built by the host (ring_host_tables)
    int num_channels;
    int hop;
    int tmod;                   // timeCursor mod N (multiple of hop)
    int stagger_ns;             // experiment: delay odd warps by this much before the first pass
    int skip;                   // experiment (PVB_SKIP, results become wrong): bit 0 the whole middle, bit 1 forward
                                // and inverse pass 2, bit 2 split + unsplit stores / loads of the spectrum
    int early;                  // which state loads may precede griddepcontrol.wait (0, 1, 2; see the kernel)
    // per-pair completion flags: done[pair] holds the sequence number of the last call of this handle
    // whose state / output for that pair is complete (release store at the end of every launch)
    unsigned *done;
    unsigned wait_seq, my_seq;  // this call may touch a pair once done[pair] >= wait_seq; it stores my_seq
    int flag_mode;              // 1: synchronise on done[] per pair instead of waiting for the whole previous grid
    unsigned *stuck;            // incremented if a flag never arrives (bounded spin; see pvb_ring_stuck_count)
    float pitch_factor;
    int pf_mant, pf_shift;      // pitch_factor == pf_mant * 2^-pf_shift (exact)
};

__device__ __forceinline__ float4 pack4(cpx2 v) { return make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
__device__ __forceinline__ cpx2 unpack4(float4 v) { return cpx2{make_float2(v.x, v.y), make_float2(v.z, v.w)}; }

// forward real-split of one (k, M-k) pair: 2 X[k] -> *dk, 2 X[M-k] -> *dm (both channels)
__device__ __forceinline__ void ring_split(cpx2 za, cpx2 zb, float2 w, float4 *dk, float4 *dm) {
    const float2 e_r = add2(za.re, zb.re), e_i = sub2(za.im, zb.im);
    const float2 o_r = add2(za.im, zb.im), o_i = sub2(zb.re, za.re);
    const cpx2 tt = cmul_s(cpx2{o_r, o_i}, w.x, w.y);
    *dk = pack4(cpx2{add2(e_r, tt.re), add2(e_i, tt.im)});
    *dm = pack4(cpx2{sub2(e_r, tt.re), sub2(tt.im, e_i)});
}

// Hermitian C2R pre-pass of one (k, M-k) pair
__device__ __forceinline__ void ring_unsplit(cpx2 yk, cpx2 ym, float2 w, cpx2 &zk, cpx2 &zmk) {
    const float2 e_r = add2(yk.re, ym.re), e_i = sub2(yk.im, ym.im);
    const float2 d_r = sub2(yk.re, ym.re), d_i = add2(yk.im, ym.im);
    const cpx2 pp = cmul_s(cpx2{d_r, d_i}, w.x, -w.y);
    zk = cpx2{sub2(e_r, pp.im), add2(e_i, pp.re)};
    zmk = cpx2{add2(e_r, pp.im), sub2(pp.re, e_i)};
}

// one bin of the shifted spectrum (both channels) from the four planes of Y
__device__ __forceinline__ cpx2 ring_load_planes(const unsigned char *mine, int slot) {
    constexpr int PLW = RingGeo::XQ_SLOTS;                             // words per plane
    const float *y = reinterpret_cast<const float *>(mine) + slot;
    return cpx2{make_float2(y[0], y[PLW]), make_float2(y[2 * PLW], y[3 * PLW])};
}

// 5-point strict maxima (pv:95-116) of bins b0 .. b0+15 from squared magnitudes m[0..19] of bins
// b0-2 .. b0+17 (non-negative floats order like their bit patterns)
__device__ __forceinline__ uint32_t ring_peak_mask(const int (&m)[20]) {
    int q[19];
#pragma unroll
    for (int t = 0; t < 19; t++) q[t] = max(m[t], m[t + 1]);
    uint32_t mask = 0;
#pragma unroll
    for (int e = 15; e >= 0; e--) {
        const int nb_max = max(q[e], q[e + 3]);                      // bins e-2, e-1, e+1, e+2
        mask = __funnelshift_l(uint32_t(nb_max - m[e + 2]), mask, 1);
    }
    return mask;
}

// Region of influence of every bin of the run, for one channel (pv:124-141): the owner of a bin
// is the nearest peak, ties go to the higher one.  Peaks travel as KEYS (see the key table in
// the kernel): high half = 2 * (peak + 2048), low half = delta + 32768.  Returns per bin the byte
// offset of its destination word inside plane 0 of Y (dump slot when it falls outside [0, nb)),
// with the sign bit set when the bin belongs to the LEFT half of its region while contracting:
// those are added on top in the second sub-step, everything else is stored first (right halves are
// pairwise disjoint after the shift, and so are left halves, for pitch factors >= 0.75).
// The integer pipe runs at half rate, so this loop is written for the fewest ALU operations.
__device__ __forceinline__ void ring_owner_scan(uint32_t mask, int lane, uint32_t nz, const int (&rk)[16],
                                                const int *krun, int second_flag, int (&dst)[16],
                                                int &d_last) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int b0 = 16 * lane;
    const int own_last = krun[(31 - __clz(mask)) & 15];              // keys of this lane's last / first peak
    const int own_first = krun[(__ffs(mask) - 1) & 15];
    const uint32_t below = nz & ((1u << lane) - 1u);
    const uint32_t above = nz & ~((2u << lane) - 1u);
    int pkey = __shfl_sync(FULL, own_last, (31 - __clz(below)) & 31);
    int nkey = __shfl_sync(FULL, own_first, (__ffs(above) - 1) & 31);
    if (!below) pkey = 0;                                            // "peak" at -2048: never the nearest
    if (!above) nkey = (2 * 8190) << 16;                             // "peak" at +6142
    const int lkey = __shfl_sync(FULL, own_last, (31 - __clz(nz)) & 31);
    d_last = (lkey & 0xFFFF) - 32768;

    int nx[16];
#pragma unroll
    for (int e = 15; e >= 0; e--) {
        nx[e] = nkey;                                                // first peak above bin e
        if ((mask >> e) & 1u) nkey = rk[e];
    }
    const int thr0 = (4 * (b0 + 2048) + 2) << 16;
    const int cb = b0 - 32768;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        if ((mask >> e) & 1u) pkey = rk[e];                          // last peak at or below bin e
        // next - b <= b - prev  <=>  2 next' + 2 prev' (+ carry of the low halves) < 4 b' + 2
        const int tt = nx[e] + pkey - thr0;
        const bool take_next = tt < ((4 * e) << 16);
        const int okey = take_next ? nx[e] : pkey;
        const int d = (okey & 0xFFFF) + cb + e;
        const unsigned slot = min(unsigned(d + (d >> 4)), 545u);     // d < 0 or d >= nb: dump slot
        dst[e] = int(4u * slot) | ((tt - ((4 * e) << 16)) & second_flag);
    }
}

// NBLK = hop / 128 and JB = ring 128-block that receives the new input block, as template
// parameters (NBLK > 0), make the role of every ring block (history / new input / emitted head /
// zero tail) a compile-time fact: no predicated duplicates of the global accesses.  NBLK == 0 is
// the same kernel with both read from the parameters (launch-uniform branches).
template <int NBLK, int JB>
__global__ void __launch_bounds__(RingGeo::MAX_WARPS * 32, 2)
pv_process_ring_kernel(const RingParams p) {
    using G = RingGeo;
    constexpr int N = G::N, NB = G::NB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int pair = blockIdx.x * (blockDim.x >> 5) + warp;
    const bool live = 2 * pair < p.num_channels;
    const unsigned FULL = 0xFFFFFFFFu;
    int *ktab = reinterpret_cast<int *>(smem_raw);
    const float2 *tw1 = reinterpret_cast<const float2 *>(smem_raw + G::OFF_TW1);
    const float2 *w64 = reinterpret_cast<const float2 *>(smem_raw + G::OFF_W64);
    const float2 *twh = reinterpret_cast<const float2 *>(smem_raw + G::OFF_TWH);
    const float *swin = reinterpret_cast<const float *>(smem_raw + G::OFF_WIN);
    const float *swout = reinterpret_cast<const float *>(smem_raw + G::OFF_WOUT);
    unsigned char *mine = smem_raw + G::TAB_BYTES + size_t(warp) * G::WARP_BYTES;
    float4 *ex = reinterpret_cast<float4 *>(mine);
    float4 *XQ = reinterpret_cast<float4 *>(mine);

    const int c0 = 2 * pair;
    const bool has1 = c0 + 1 < p.num_channels;
    const int hop = p.hop;
    const int t = p.tmod;
    const int nblk = NBLK ? NBLK : (hop >> 7);
    const int jb = NBLK ? JB : (((t - hop + N) >> 7) & 7);    // ring 128-block that receives the new input block
    const int je = (jb + nblk) & 7;                           // ring 128-block of frame sample 0 (emitted)

    // Programmatic dependent launch: our CTAs may become resident while the previous kernel on the
    // stream drains.  Two ways to respect what earlier launches wrote:
    //  * flag mode (the host has checked that the caller's buffers do not alias those of recent
    //    launches): dependents are released at once, every warp waits for ITS pair's completion
    //    flag only, and the CTA waits for the previous grid at its very end, so that "this grid is
    //    complete" still implies "everything before it is complete".  Calls of different handles, and
    //    different pairs of one handle, then overlap freely: the load phase of one CTA runs under
    //    the compute phase of its SM neighbour.
    //  * grid mode: griddepcontrol.wait before the first dependent access; everything up to it
    //    touches only constant tables (and state that is provably older than the previous kernel).
    if (p.flag_mode) asm volatile("griddepcontrol.launch_dependents;");
    // ---- CTA-shared tables: asynchronous 16-byte copies, fixed trip counts (CTAs have 4..7 warps;
    // no division by blockDim) --------------------------------------------------------------------------
    {
        const int rot = (N - t) & (N - 1);
        const float4 *w1 = reinterpret_cast<const float4 *>(p.window2 + rot);
        const float4 *w2 = reinterpret_cast<const float4 *>(p.window_out2 + rot);
        const unsigned s_tab = unsigned(__cvta_generic_to_shared(smem_raw + G::OFF_TW1));
        const unsigned s_win = unsigned(__cvta_generic_to_shared(smem_raw + G::OFF_WIN));
        const unsigned s_wout = unsigned(__cvta_generic_to_shared(smem_raw + G::OFF_WOUT));
#pragma unroll
        for (int k = 0; k < (G::GTAB_BYTES / 16 + 127) / 128; k++) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < G::GTAB_BYTES / 16)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s_tab + 16 * i), "l"(p.gtab + i));
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < N / 4) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s_win + 16 * i), "l"(w1 + i));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s_wout + 16 * i), "l"(w2 + i));
            }
        }
        // key table: what a peak at bin pk contributes to the region scan: its position and
        // delta = round(pk * pitchFactor) - pk in exact integer arithmetic (pv:125-127)
        const long long pf_m = p.pf_mant;
        const int pf_s = p.pf_shift;
        const long long half = 1ll << (pf_s - 1);
#pragma unroll
        for (int k = 0; k < (NB + 1 + 127) / 128; k++) {
            const int pk = threadIdx.x + k * blockDim.x;
            if (pk <= NB) {
                const long long ps = (pf_m * pk + half) >> pf_s;
                const int delta = (ps <= NB) ? int(ps) - pk : G::INVALID_DELTA;
                ktab[pk + 4 * (pk >> 4)] = ((2 * (pk + 2048)) << 16) | (delta + 32768);
            }
        }
    }
    // ---- frame loads: all issued before anything consumes them ---------------------------------
    // History written by launches OLDER than the kernel in front of us on the stream is already
    // complete when our CTAs start (that kernel passed its own griddepcontrol.wait before it let us
    // launch), so those loads are issued before our wait and overlap the previous launch's tail:
    //   p.early == 2: none of this handle's state was written by the previous kernel -> all of hist
    //   p.early == 1: the previous kernel may be this handle's last call -> all but its newest block
    //   p.early == 0: everything after the wait
    // The input block always waits (it belongs to the caller's stream order).
    float4 r[16];
    float2 un0[NBLK ? 2 * NBLK : 1], un1[NBLK ? 2 * NBLK : 1];
    float4 *hl = p.hist2 + size_t(live ? pair : 0) * (N / 2) + lane;
    const int early = p.flag_mode ? 0 : p.early;
    if (p.flag_mode) {
        if (live) {
            // acquire: the previous call of this handle has finished with this pair (bounded spin: a
            // lost flag must not hang the device; the host can read the stuck counter)
            unsigned v;
            int it = 0;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.done + pair) : "memory");
                if (int(v - p.wait_seq) >= 0) break;
                if (++it > 200000) {
                    if (lane == 0) atomicAdd(p.stuck, 1u);
                    break;
                }
                __nanosleep(200);
            }
            __syncwarp();
        }
    } else if (live && early) {
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int h = e >> 3, j = e & 7;
            const int jj = (j - jb) & 7;                              // launch-uniform
            // jj < nblk: new input; jj >= 8 - nblk: the block the previous call wrote
            if (jj >= nblk && (jj < 8 - nblk || early == 2)) r[e] = hl[32 * h + 64 * j];
        }
        if (early == 2) {
            // warm L2 with the overlap-add ring lines the tail of this kernel adds to
            const int line = 16 * lane;                               // float4 index: 256 bytes per lane
            if ((((line >> 6) - jb) & 7) >= nblk) {
                const float4 *ap = p.acc2 + size_t(pair) * (N / 2) + line;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ap));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + 8));
            }
        }
    }
    if (!p.flag_mode) {
        // our dependents may launch only now: whoever starts behind us can rely on everything older
        // than us being complete
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    if (live) {
        const float *i0 = p.in ? p.in + size_t(c0) * hop + 2 * lane : nullptr;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int h = e >> 3, j = e & 7;
            const int jj = (j - jb) & 7;                              // launch-uniform
            if (jj < nblk) {
                float2 u0 = make_float2(0.f, 0.f), u1 = make_float2(0.f, 0.f);
                if (i0) {
                    u0 = __ldg(reinterpret_cast<const float2 *>(i0 + 64 * h + 128 * jj));
                    if (has1) u1 = __ldg(reinterpret_cast<const float2 *>(i0 + hop + 64 * h + 128 * jj));
                }
                if constexpr (NBLK > 0) {
                    un0[h * NBLK + jj] = u0;                          // packed after the barrier: the moves
                    un1[h * NBLK + jj] = u1;                          // must not sit between the loads
                } else {
                    r[e] = make_float4(u0.x, u1.x, u0.y, u1.y);
                }
            } else if (!(early && (jj < 8 - nblk || early == 2))) {
                r[e] = hl[32 * h + 64 * j];
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (!live) return;          // no CTA-wide barriers below

    if (p.stagger_ns > 0 && (warp & 1)) __nanosleep(unsigned(p.stagger_ns));
    // the new block joins the history ring (ola:105)
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int h = e >> 3, j = e & 7;
        const int jj = (j - jb) & 7;
        if (jj < nblk) {
            if constexpr (NBLK > 0) {
                const float2 u0 = un0[h * NBLK + jj], u1 = un1[h * NBLK + jj];
                r[e] = make_float4(u0.x, u1.x, u0.y, u1.y);
            }
            hl[32 * h + 64 * j] = r[e];
        }
    }

    // ---- Hann window (pv:55) + forward pass 1: butterflies n = lane + 32 h over j (stride 64) ---
    {
        const float *wl = swin + 2 * lane;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int nl = lane + 32 * h;
            cpx2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float2 w = *reinterpret_cast<const float2 *>(wl + 64 * h + 128 * j);
                const float4 v = r[8 * h + j];
                x[j].re = mul2(make_float2(v.x, v.y), bc2(w.x));
                x[j].im = mul2(make_float2(v.z, v.w), bc2(w.y));
            }
            dft8<false>(x);
#pragma unroll
            for (int k1 = 1; k1 < 8; k1++) {
                const float2 w = tw1[G::TW1_ROW * k1 + nl];           // W_512^{n k1}
                x[k1] = cmul_s(x[k1], w.x, w.y);
            }
#pragma unroll
            for (int k1 = 0; k1 < 8; k1++) ex[65 * k1 + nl] = pack4(x[k1]);
        }
    }
    __syncwarp();

    // warm L2 with the overlap-add ring lines the tail of this kernel adds to (the slot that is
    // only written, ring [t - hop, t), is skipped)
    {
        const int line = 16 * lane;                                   // float4 index: 256 bytes per lane
        if (early != 2 && (((line >> 6) - jb) & 7) >= nblk) {
            const float4 *ap = p.acc2 + size_t(pair) * (N / 2) + line;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ap));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + 8));
        }
    }

    // ---- forward pass 2: butterflies (k1, m3) over m2, in place -----------------------------------
    const int m3l = lane & 7;
    if (!(p.skip & 2)) {
        float2 w2[8];
#pragma unroll
        for (int k2 = 1; k2 < 8; k2++) w2[k2] = w64[8 * k2 + m3l];              // W_64^{m3 k2}
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float4 *bp = ex + 65 * ((lane >> 3) + 4 * h) + m3l;
            cpx2 x[8];
#pragma unroll
            for (int m2 = 0; m2 < 8; m2++) x[m2] = unpack4(bp[8 * m2]);
            dft8<false>(x);
#pragma unroll
            for (int k2 = 1; k2 < 8; k2++) x[k2] = cmul_s(x[k2], w2[k2].x, w2[k2].y);
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) bp[8 * k2] = pack4(x[k2]);
        }
    }
    __syncwarp();

    // ---- forward pass 3: butterflies A (bins lane + 64 j) and B (bins 64 - lane + 64 j) ------------
    // lane 0 owns the two self-paired butterflies: A = bins 64 j, B = bins 32 + 64 j
    const bool l0 = lane == 0;
    const int kB = l0 ? 32 : 64 - lane;
    const int exA = 65 * (lane & 7) + 8 * (lane >> 3);
    const int exB = 65 * (kB & 7) + 8 * (kB >> 3);
    cpx2 a[8], b[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        a[c] = unpack4(ex[exA + c]);
        b[c] = unpack4(ex[exB + c]);
    }
    dft8<false>(a);      // a[j] = Z[lane + 64 j]
    dft8<false>(b);      // b[j] = Z[kB + 64 j]
    __syncwarp();        // everyone has read the exchange slots: X may overwrite them

    // ---- real split in registers -> XQ (2x scaled) --------------------------------------------------
    // slot j pairs (a[j], b[7-j]) at k = lane + 64 j.  lane 0: j < 4: (b[j], b[7-j]) at k = 32 + 64 j;
    // j >= 4: (a[j-4], a[(12-j)&7]) at k = 64 (j-4); plus the self pair k = 256 (a[4]).
    const int gA = lane + (lane >> 4);                                // slot of bin lane
    const int gB = 544 - lane - ((lane + 15) >> 4);                   // slot of bin 512 - lane
    const int sAlo = l0 ? 34 : gA, sAhi = l0 ? -272 : gA;
    const int sBlo = l0 ? 510 : gB, sBhi = l0 ? 816 : gB;
    const int tlo = l0 ? 32 : lane, thi = l0 ? -256 : lane;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const cpx2 za = sel(l0, b[j], a[j]);
        ring_split(za, b[7 - j], twh[tlo + 64 * j], XQ + sAlo + 68 * j, XQ + sBlo - 68 * j);
    }
#pragma unroll
    for (int j = 4; j < 8; j++) {
        const cpx2 za = sel(l0, a[j - 4], a[j]);
        const cpx2 zb = sel(l0, a[(12 - j) & 7], b[7 - j]);
        ring_split(za, zb, twh[thi + 64 * j], XQ + sAhi + 68 * j, XQ + sBhi - 68 * j);
    }
    if (l0) ring_split(a[4], a[4], twh[256], XQ + 272, XQ + 272);
    __syncwarp();

    // ---- peaks, regions of influence, shift (pv:95-173) -------------------------------------------------
    if (!(p.skip & 1))
    // X lives in float4 slots (both channels per bin); the shifted spectrum Y is written over it as
    // four planes of floats (re0 | re1 | im0 | im1, bin d at word d + (d >> 4)): the 32-bit scatter of
    // lanes that own runs 16 bins apart then spreads over all banks.
    {
        const bool contract = p.pitch_factor < 1.0f;
        const float4 *runp = XQ + 17 * lane;                          // slot of bin 16 lane
        uint32_t mask0, mask1;
        {
            const float4 *hlo = lane ? runp - 3 : XQ;                 // bins 16 lane - 2, - 1 (lane 0: unused)
            int m0[20], m1[20];
#pragma unroll
            for (int i = 0; i < 20; i++) {
                const float4 v = (i < 2) ? hlo[i] : (i < 18) ? runp[i - 2] : runp[i - 1];   // i >= 18: bins 16 lane + 16, + 17 (slot 16 is padding)
                const float2 re = make_float2(v.x, v.y), im = make_float2(v.z, v.w);
                const float2 mg = fma2(re, re, mul2(im, im));         // pv:82-92, float32
                m0[i] = __float_as_int(mg.x);
                m1[i] = __float_as_int(mg.y);
            }
            mask0 = ring_peak_mask(m0);
            mask1 = ring_peak_mask(m1);
            if (lane == 0) { mask0 &= ~3u; mask1 &= ~3u; }            // i >= 2
            if (lane == 31) { mask0 &= ~(1u << 15); mask1 &= ~(1u << 15); }     // i <= nb - 3
        }
        const uint32_t nz0 = __ballot_sync(FULL, mask0 != 0);
        const uint32_t nz1 = __ballot_sync(FULL, mask1 != 0);

        int dst0[16], dst1[16];
        int dl0 = 0, dl1 = 0;
        {
            const int *krun = ktab + 20 * lane;                       // keys of bins 16 lane .. + 15
            int rk[16];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int4 kv = *reinterpret_cast<const int4 *>(krun + 4 * i);
                rk[4 * i] = kv.x; rk[4 * i + 1] = kv.y; rk[4 * i + 2] = kv.z; rk[4 * i + 3] = kv.w;
            }
            // both scans run unconditionally (a channel without peaks ends up with every bin on the
            // dump slot): two independent instruction streams the scheduler can interleave
            const int second_flag = contract ? int(0x80000000u) : 0;
            ring_owner_scan(mask0, lane, nz0, rk, krun, second_flag, dst0, dl0);
            ring_owner_scan(mask1, lane, nz1, rk, krun, second_flag, dst1, dl1);
        }

        // sources into registers: own run, bin 512 and the first stale level (what _realTransform4
        // leaves in slots N/2 + q, bundle:394-438, rebuilt from the valid half); bins 512 + lane + 32 i
        float4 xv[16];
#pragma unroll
        for (int e = 0; e < 16; e++) xv[e] = runp[e];
        float4 ext[4];
        ext[0] = ext[1] = ext[2] = ext[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l0) ext[0] = XQ[544];
        if (contract) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int q = lane + 32 * i;
                const int qq = q ? q : 1;
                const int sq = qq + (qq >> 4);                        // slot of bin q
                const int sm = 544 - qq - ((qq + 15) >> 4);           // slot of bin 512 - q
                const cpx2 A = unpack4(XQ[sq]), Bv = unpack4(XQ[sq + 272]);        // bins q, 256 + q
                const cpx2 Cv = unpack4(XQ[sm]), D = unpack4(XQ[sm - 272]);        // bins 512 - q, 256 - q
                const float2 sr = add2(sub2(A.re, Bv.re), sub2(Cv.re, D.re));
                const float2 si = sub2(sub2(A.im, Bv.im), sub2(Cv.im, D.im));
                const float2 w = twh[2 * qq];
                const cpx2 sv = cmul_s(cpx2{sr, si}, 0.25f * w.x, -0.25f * w.y);
                if (q) ext[i] = pack4(sv);
            }
        }
        __syncwarp();            // every lane holds its sources: the buffer becomes Y
#pragma unroll
        for (int i = 0; i < 17; i++) XQ[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < 2) XQ[544 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();

        constexpr int PL = 4 * G::XQ_SLOTS;                           // bytes per plane (546 words)
        // first sub-step: plain stores (pairwise disjoint destinations)
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (dst0[e] >= 0) {
                *reinterpret_cast<float *>(mine + dst0[e]) = xv[e].x;
                *reinterpret_cast<float *>(mine + dst0[e] + 2 * PL) = xv[e].z;
            }
            if (dst1[e] >= 0) {
                *reinterpret_cast<float *>(mine + dst1[e] + PL) = xv[e].y;
                *reinterpret_cast<float *>(mine + dst1[e] + 3 * PL) = xv[e].w;
            }
        }
        if (nz0) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int d = 512 + lane + 32 * i + dl0;
                if (unsigned(d) < unsigned(NB)) {
                    *reinterpret_cast<float *>(mine + 4 * (d + (d >> 4))) = ext[i].x;
                    *reinterpret_cast<float *>(mine + 4 * (d + (d >> 4)) + 2 * PL) = ext[i].z;
                }
            }
        }
        if (nz1) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int d = 512 + lane + 32 * i + dl1;
                if (unsigned(d) < unsigned(NB)) {
                    *reinterpret_cast<float *>(mine + 4 * (d + (d >> 4)) + PL) = ext[i].y;
                    *reinterpret_cast<float *>(mine + 4 * (d + (d >> 4)) + 3 * PL) = ext[i].w;
                }
            }
        }
        if (contract) {
            // second sub-step: left halves add on top (pairwise disjoint among themselves, so the
            // loads of a batch can all be issued before the first store)
            __syncwarp();
            unsigned char *mine2 = mine + 0x80000000u;                // cancels the flag bit of dst
#pragma unroll
            for (int g = 0; g < 2; g++) {
                float o0r[8], o0i[8], o1r[8], o1i[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int e = 8 * g + i;
                    o0r[i] = o0i[i] = o1r[i] = o1i[i] = 0.f;
                    if (dst0[e] < 0) { o0r[i] = *reinterpret_cast<float *>(mine2 + dst0[e]); o0i[i] = *reinterpret_cast<float *>(mine2 + dst0[e] + 2 * PL); }
                    if (dst1[e] < 0) { o1r[i] = *reinterpret_cast<float *>(mine2 + dst1[e] + PL); o1i[i] = *reinterpret_cast<float *>(mine2 + dst1[e] + 3 * PL); }
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int e = 8 * g + i;
                    if (dst0[e] < 0) {
                        *reinterpret_cast<float *>(mine2 + dst0[e]) = o0r[i] + xv[e].x;
                        *reinterpret_cast<float *>(mine2 + dst0[e] + 2 * PL) = o0i[i] + xv[e].z;
                    }
                    if (dst1[e] < 0) {
                        *reinterpret_cast<float *>(mine2 + dst1[e] + PL) = o1r[i] + xv[e].y;
                        *reinterpret_cast<float *>(mine2 + dst1[e] + 3 * PL) = o1i[i] + xv[e].w;
                    }
                }
            }
        }
    }
    __syncwarp();

    // ---- Hermitian C2R pre-pass in registers (mirror of the split) -------------------------------------
    {
        cpx2 zk[8], zmk[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int sa = (j < 4 ? sAlo : sAhi) + 68 * j, sb = (j < 4 ? sBlo : sBhi) - 68 * j;
            cpx2 yk = ring_load_planes(mine, sa), ym = ring_load_planes(mine, sb);
            if (j == 4) {                        // lane 0: k == 0, bins 0 and N/2 enter with their real part only
                yk.im = make_float2(l0 ? 0.f : yk.im.x, l0 ? 0.f : yk.im.y);
                ym.im = make_float2(l0 ? 0.f : ym.im.x, l0 ? 0.f : ym.im.y);
            }
            const float2 w = twh[(j < 4 ? tlo : thi) + 64 * j];
            ring_unsplit(yk, ym, w, zk[j], zmk[j]);
        }
        cpx2 z256, dummy;
        {
            const cpx2 y = ring_load_planes(mine, 272);
            ring_unsplit(y, y, twh[256], z256, dummy);
        }
        a[0] = sel(l0, zk[4], zk[0]);
        a[1] = sel(l0, zk[5], zk[1]);
        a[2] = sel(l0, zk[6], zk[2]);
        a[3] = sel(l0, zk[7], zk[3]);
        a[4] = sel(l0, z256, zk[4]);
        a[5] = sel(l0, zmk[7], zk[5]);
        a[6] = sel(l0, zmk[6], zk[6]);
        a[7] = sel(l0, zmk[5], zk[7]);
        b[0] = sel(l0, zk[0], zmk[7]);
        b[1] = sel(l0, zk[1], zmk[6]);
        b[2] = sel(l0, zk[2], zmk[5]);
        b[3] = sel(l0, zk[3], zmk[4]);
        b[4] = zmk[3];
        b[5] = zmk[2];
        b[6] = zmk[1];
        b[7] = zmk[0];
    }
    __syncwarp();        // everyone has read Y: the exchange slots may overwrite it

    // ---- inverse pass 1 (DIT): butterflies A and B over k3, twiddle conj(W_64^{k2 m3}) -------------------
    dft8<true>(a);
    dft8<true>(b);
    {
        const int k2a = lane >> 3, k2b = kB >> 3;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            cpx2 va = a[c], vb = b[c];
            if (c > 0) {
                const float2 wa = w64[8 * c + k2a];
                const float2 wb = w64[8 * c + k2b];
                va = cmul_s(va, wa.x, -wa.y);
                vb = cmul_s(vb, wb.x, -wb.y);
            }
            ex[exA + c] = pack4(va);
            ex[exB + c] = pack4(vb);
        }
    }
    __syncwarp();

    // accumulator values (L2 hits thanks to the prefetch) are requested before the last exchange so
    // that their latency hides behind inverse pass 2; the tail slot starts from zero (ola:134)
    float4 *al = p.acc2 + size_t(pair) * (N / 2) + lane;
    float4 q[16];
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int h = e >> 3, j = e & 7;
        q[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (((j - jb) & 7) >= nblk) q[e] = al[32 * h + 64 * j];
    }

    // ---- inverse pass 2: butterflies (k1, m3) over k2, twiddle conj(W_512^{k1 (m3 + 8 m2)}) ---------------
    if (!(p.skip & 2))
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k1 = (lane >> 3) + 4 * h;
        float4 *bp = ex + 65 * k1 + m3l;
        const float2 *twp = tw1 + G::TW1_ROW * k1 + m3l;
        cpx2 x[8];
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) x[k2] = unpack4(bp[8 * k2]);
        dft8<true>(x);
#pragma unroll
        for (int m2 = 0; m2 < 8; m2++) {
            const float2 w = twp[8 * m2];
            bp[8 * m2] = pack4(cmul_s(x[m2], w.x, -w.y));
        }
    }
    __syncwarp();

    // ---- inverse pass 3: butterflies n over k1 -> ring samples; window, overlap-add, emit ------------------
    {
        float *o0 = p.out + size_t(c0) * hop + 2 * lane;
        const float *wol = swout + 2 * lane;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int nl = lane + 32 * h;
            cpx2 x[8];
#pragma unroll
            for (int k1 = 0; k1 < 8; k1++) x[k1] = unpack4(ex[65 * k1 + nl]);
            dft8<true>(x);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                // window_out = hannWindow / (2 N R): fromComplexArray, applyHannWindow and the division
                // by nbOverlaps (pv:65-67, ola:153) in one multiply (the scales are powers of two)
                const float2 wo = *reinterpret_cast<const float2 *>(wol + 64 * h + 128 * j);
                const float4 qv = q[8 * h + j];
                const float2 y0 = fma2(x[j].re, bc2(wo.x), make_float2(qv.x, qv.y));
                const float2 y1 = fma2(x[j].im, bc2(wo.y), make_float2(qv.z, qv.w));
                const int jj = (j - je) & 7;
                if (jj < nblk) {                                      // head: emit (ola:111-118)
                    *reinterpret_cast<float2 *>(o0 + 64 * h + 128 * jj) = make_float2(y0.x, y1.x);
                    if (has1) *reinterpret_cast<float2 *>(o0 + hop + 64 * h + 128 * jj) = make_float2(y0.y, y1.y);
                } else {
                    al[32 * h + 64 * j] = make_float4(y0.x, y0.y, y1.x, y1.y);
                }
            }
        }
    }
    // release: state and output of this pair are complete for call my_seq
    __threadfence();
    __syncwarp();
    if (lane == 0)
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.done + pair), "r"(p.my_seq) : "memory");
    // flag mode skipped the wait at the top: take it here, where the previous grid is long gone, so
    // that completion stays transitive along the stream
    if (p.flag_mode) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// tables the ring-order kernel copies into shared memory: tw1[k1][n] (rows of TW1_ROW), w64[a][b], twh[k]
inline void ring_host_tables(const float2 *tw /* [1024] W_1024^j */, float2 *out /* GTAB_BYTES / 8 */) {
    using G = RingGeo;
    float2 *tw1 = out, *w64 = out + G::TW1_BYTES / 8, *twh = w64 + G::W64_BYTES / 8;
    for (int i = 0; i < G::GTAB_BYTES / 8; i++) out[i] = make_float2(0.f, 0.f);
    for (int k1 = 0; k1 < 8; k1++)
        for (int n = 0; n < 64; n++) tw1[G::TW1_ROW * k1 + n] = tw[(2 * n * k1) & 1023];
    for (int a = 0; a < 8; a++)
        for (int b = 0; b < 8; b++) w64[8 * a + b] = tw[(16 * a * b) & 1023];
    for (int k = 0; k <= 512; k++) twh[k] = tw[k];
}

// planar [C'][N] rings (pv_kernel.cuh conventions: frame sample n at hist[(n + rb + hop) mod N],
// accumulator sample k at acc[(k + rb) mod N]) <-> paired rings aligned to t
__global__ void pv_ring_convert_kernel(float *hist_planar, float *acc_planar, float2 *hist2, float2 *acc2,
                                       int pairs, int n, int hop, int rb, int tmod, int to_paired) {
    const long long total = (long long)pairs * n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int pr = int(idx / n), s = int(idx % n);               // frame-order sample s of pair pr
        const int ih = (s + rb + hop) & (n - 1), ia = (s + rb) & (n - 1), i2 = (s + tmod) & (n - 1);
        float *h0 = hist_planar + (size_t(2 * pr) * n), *h1 = h0 + n;
        float *a0 = acc_planar + (size_t(2 * pr) * n), *a1 = a0 + n;
        if (to_paired) {
            hist2[size_t(pr) * n + i2] = make_float2(h0[ih], h1[ih]);
            acc2[size_t(pr) * n + i2] = (s < n - hop) ? make_float2(a0[ia], a1[ia]) : make_float2(0.f, 0.f);
        } else {
            const float2 hv = hist2[size_t(pr) * n + i2], av = acc2[size_t(pr) * n + i2];
            h0[ih] = hv.x; h1[ih] = hv.y;
            a0[ia] = (s < n - hop) ? av.x : 0.f;
            a1[ia] = (s < n - hop) ? av.y : 0.f;
        }
    }
}

}  // namespace pvb
