// pv_kernel_ring.cuh — ring-order fused kernel, sm_100a, for every frame size: every thread owns 16
// complex points of the half-size FFT, so a channel pair takes a quarter of a warp (frame 256), half a
// warp (512), one warp (1024), two (2048) or four warps (4096).  Same reference arithmetic as
// pv_kernel.cuh (one launch == one process() call, ola-processor.js:159-171 +
// phase-vocoder.js:45-72), reorganised around one identity:
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
// Valid for hop % 128 == 0 (frame 256: hop 64 or 128; frame 4096: hop % 256 == 0), hop <= N / 2 and
// pitch factors in [0.75, 64] (first stale level only, right halves and left halves of regions stay
// pairwise disjoint after the shift).  tests/
// ring_kernel_model.py restates this file lane by lane in numpy (CPU test of the design).
#pragma once

#include "pv_kernel.cuh"
#include "pv_kernel_warp.cuh"

namespace pvb {

// channel pairs (warps) per CTA at frame 1024; two CTAs per SM at 128 registers.  7 fills one wave
// of 4096 channels on 148 SMs exactly; 8 uses the whole register file (16 warps per SM).
#ifndef PVB_RING_LANE_FENCE
#define PVB_RING_LANE_FENCE 0
#endif
// The shifted spectrum Y lives in four planes of 32-bit words, bin d at word d + (d >> YSHIFT).  The
// pad decides which lanes of the scatter collide (lanes own runs 16 bins apart, shifted by their
// region's delta); 5 is the best single choice over pitch factors 0.75 .. 2 (model: 1.8 - 2.8
// wavefronts per store; 4: 1.9 - 4.7, worst at pitch factor 1.25).
#ifndef PVB_RING_YSHIFT
#define PVB_RING_YSHIFT 5
#endif
// Exact first-writer classification while contracting: of a region's left half only the first
// c = delta_prev - delta_next bins land on the previous region's right half; only those take the
// read-add-store sub-step, everything else is stored once, and since the shifted regions then cover
// [0, nb) exactly the zero fill is dropped (5 more integer operations per bin and channel, ~160
// fewer shared-memory wavefronts per pair).
#ifndef PVB_RING_EXACT
#define PVB_RING_EXACT 0
#endif
#ifndef PVB_RING_PAIRS_1024
#define PVB_RING_PAIRS_1024 7
#endif
// pairs per CTA of the MULTI (several calls per launch) and DEEP instances at frame 1024: 6 -> 168 registers.
// (7 at 144 registers would fill one wave with 4096 channels, but 14 warps of 144 registers do not fit the four
// 16 K-register files of an SM -- one sub-partition gets four warps -- so only one CTA becomes resident: measured
// 2x slower, profiles/r02_ab_m7.txt)
#ifndef PVB_RING_MULTI_PAIRS_1024
#define PVB_RING_MULTI_PAIRS_1024 6
#endif
// Destinations outside [0, nb) are clamped to one write-only "dump" word per plane instead of being
// predicated off (one integer operation less per bin): several lanes may store to it at once, which
// compute-sanitizer's racecheck reports as write-after-write hazards.  -DPVB_RING_NO_DUMP=1 predicates
// those stores off instead; that build must be racecheck-clean (profiles/sanitize.sh), which shows that
// the dump word is the only shared-memory word ever written by two lanes between two barriers.
#ifndef PVB_RING_NO_DUMP
#define PVB_RING_NO_DUMP 0
#endif
// Gather middle (see "gather middle" in ring_one_call): the shifted spectrum is GATHERED in the order the
// Hermitian pre-pass wants it, from region descriptors in destination space; it never exists in shared
// memory (no zero fill, no scatter, no read-add-store sub-step, no owner scan per bin).  Measured on B200 at
// frame 1024 (profiles/r02_gather_ab.txt): parity-green on every GPU test, shared-memory wavefronts per
// pair 1706 -> 1410, but 4652 instead of 4296 instructions per pair (the per-peak descriptor loop runs as
// long as the busiest lane's run has peaks, and a quarter of its instructions copy descriptors to group
// starts), and the kernel retires instructions at the same rate either way: 17.9 us against 16.5 us per
// launch.  The default build therefore keeps the scatter middle; `make gather` builds this variant.
#ifndef PVB_RING_GATHER
#define PVB_RING_GATHER 0
#endif
// CTA-shared tables staged by bulk asynchronous copies (cp.async.bulk, one thread issues four copies that
// complete on an mbarrier) instead of ten 16-byte cp.async per thread: ~70 instructions less per warp.
#ifndef PVB_RING_BULK_TABLES
#define PVB_RING_BULK_TABLES 1
#endif
// Frame 1024: the exchange between forward passes 1 and 2 (and between inverse passes 2 and 3) goes through
// TENSOR MEMORY instead of shared memory.  tcgen05.st.32x32b puts register j of lane L at (lane L, column j);
// tcgen05.ld.16x256b hands thread t the columns 2 (t % 4) + {0, 1} (+ 8 g) of lanes t / 4 and t / 4 + 8
// (measured: profiles/tmem/tmem_probe.cu): together they swap two lane bits with two register bits, which is
// exactly what these two exchanges need once pass 2 runs with the lanes renumbered t = 4 m3 + a.  The data path
// of tensor memory is separate from the load/store pipe this kernel is bound by, and an 8 KB round trip costs
// 46 SM-cycles at 14 warps per SM against 88 through shared memory (profiles/tmem/tmem_bw.cu).
#ifndef PVB_RING_TMEM
#define PVB_RING_TMEM 0
#endif

// PCH: per-channel pitch factors (pvb_process_pf): the key table becomes per pair (two deltas per bin)
template <int N_, bool PCH_ = false>
struct RingGeoT {
    static constexpr bool PCH = PCH_;
    static constexpr int N = N_, M = N / 2, NB = M + 1;
    static constexpr int TP = N / 32;                       // threads per channel pair (16 complex points each):
                                                            // half a warp (512), a warp (1024), two warps (2048)
    static constexpr int WPP = TP / 32;                     // warps per pair (0: two pairs share a warp)
    static constexpr int R1 = M / 64;                       // radix of the first pass: 4, 8, 16 or 32
    static constexpr int LR1 = (R1 == 2) ? 1 : (R1 == 4) ? 2 : (R1 == 8) ? 3 : (R1 == 16) ? 4 : 5;
    // frame 256 also takes hop 64: the rings are then rotated by half a 128-sample block on odd
    // calls.  Whether a register sits in the first or the second half of its block is a compile-time
    // fact there (column tp + 8 h >= 32 <=> h >= 4), so roles stay static in units of 64 samples;
    // NBLK counts those units, and there is one first-pass twiddle table per 64-sample offset.
    static constexpr bool HB = (N == 256);
    static constexpr int UNIT = HB ? 64 : 128;              // samples per role unit (NBLK = hop / UNIT)
    static constexpr int NTAB = HB ? 2 * (N / 128) : N / 128;   // first-pass twiddle tables
    // frame 4096: a thread holds the 16 frame blocks of one parity (f = 2 f' + s) of its column and
    // does a 16-point DFT over f'; the radix-2 step that completes the 32-point DFT over f is done by
    // the reader in pass 2 (it holds rows k' and k' + 16 anyway), so there is no fourth exchange
    static constexpr int RT = (R1 > 16) ? 16 : R1;          // radix a thread does in registers in pass 1
    static constexpr int NSPL = R1 / RT;                    // 2: rows of the exchange are (s, k')
    static constexpr int NB1 = 16 / RT;                     // first-pass butterflies per thread
    static constexpr int KS = M / 8;                        // stride between the outputs of a last-pass butterfly
    static constexpr int SS = KS + KS / 16;                 // the same in spectrum slots
    static constexpr int NJ = N / 128;                      // ring blocks of 128 samples
    static constexpr int SM = M + M / 16;                   // slot of bin M
    // exchange slots of 16 bytes: element 8 r + c of row k1 at RS k1 + G8 r + c.  Rows of 64 at stride
    // 65; frame 512 (radix-4 first pass: a quarter-warp of pass 3 spans two r) pads every group of 8
    // and uses stride 74, which keeps all three passes conflict free
    static constexpr int RS = (R1 >= 8) ? 65 : (R1 == 4) ? 74 : 76;
    static constexpr int G8 = (R1 >= 8) ? 8 : 9;
    // Y planes: bin d at word d + (d >> YS); the pad must divide the last-pass butterfly stride
    static constexpr int YS = (KS % (1 << PVB_RING_YSHIFT) == 0) ? PVB_RING_YSHIFT : 4;
    static constexpr int EX_SLOTS = RS * (R1 - 1) + 8 * G8;
    static constexpr int XQ_SLOTS = SM + 2;                 // bin k at k + (k >> 4); last slot = dump / halo dummy
    // cross-warp exchange of the region scan: [4][TP] keys, [2][WPP] ballots; then the peak guard's
    // [2][WPP] energy sums, [2][WPP] uncertainty ballots and its scratch slot number
    static constexpr int SCR_BYTES = (WPP > 1) ? 4 * TP * 4 + 32 + 96 : 0;
    // ---- gather middle: X as four planes of 32-bit words (re0 | re1 | im0 | im1, bin k at word k, the first
    // stale level behind bin M), the squared magnitudes as (ch0, ch1) pairs padded every 16 bins for the run
    // reads, and later in their place one array of region descriptors per channel
    static constexpr bool GATHER = PVB_RING_GATHER && !PCH_ && (N_ == 1024);
    static constexpr int XPW = M + N / 8;                   // words per plane: bins 0 .. M + N/8 - 1
    static constexpr int X_BYTES = 4 * XPW * 4;
    static constexpr int MAG_UNITS = 17 * TP + 6;           // bin k at unit k + (k >> 4) + 3
    static constexpr int DW = M + 4;                        // descriptor words per channel (destinations 0 .. M)
    static constexpr int B_BYTES = GATHER ? ((MAG_UNITS * 8 > 2 * DW * 4) ? MAG_UNITS * 8 : 2 * DW * 4) : 0;
    static constexpr int A_SLOTS = GATHER ? ((X_BYTES / 16 > EX_SLOTS) ? X_BYTES / 16 : EX_SLOTS)
                                          : ((XQ_SLOTS > EX_SLOTS) ? XQ_SLOTS : EX_SLOTS);
    static constexpr int BUF_SLOTS = A_SLOTS + B_BYTES / 16;
    static constexpr int A_BYTES = A_SLOTS * 16;
    // descriptor fields: T' (low TB bits) | overlap c (CB bits) | -delta (signed, the rest)
    static constexpr int TB = (N_ == 256) ? 8 : (N_ == 512) ? 9 : (N_ == 1024) ? 10 : (N_ == 2048) ? 11 : 12;
    static constexpr int CB = 9;
    static constexpr int LRW = (TP >= 32) ? 5 : (TP == 16) ? 4 : 3;     // log2 of the lanes of a pair in one warp
    // two pairs per warp (frame 512): their buffers sit 16 banks apart, so the 32-bit plane accesses
    // of the two half-warps (16 consecutive words each) do not collide
    // (frame 256: four pairs per warp, 8 banks apart)
    // key table: bin p at p + 4 (p >> 4).  One per CTA (scalar pitch factor) or one per pair (PCH)
    static constexpr int DTAB_BYTES = ((NB + 1 + 4 * ((NB >> 4) + 1)) * 4 + 15) & ~15;
    static constexpr int KT_BYTES = PCH ? DTAB_BYTES : 0;
    static constexpr int PAIR_RAW = BUF_SLOTS * 16 + SCR_BYTES + KT_BYTES;
    static constexpr int PAIR_PAD = (TP < 32) ? ((TP == 16 ? 64 : 32) + 128 - PAIR_RAW % 128) % 128 : 0;
    static constexpr int PAIR_BYTES = PAIR_RAW + PAIR_PAD;
    // DEEP instances, per pair: N/8 windowed samples of both channels, then W_N^{16 j} for j < N/16 (the twiddles
    // of the sub-transform sums are all multiples of 16: read from twh they would all sit in one bank)
    // (frame 4096: only the half table W_N^{16 j}, j < N/32, fits beside two CTAs per SM; the other half is its negative)
    static constexpr bool T16_HALF = (N == 4096);
    static constexpr int DEEP_BYTES = N + (T16_HALF ? N / 4 : N / 2);
    // CTA-shared tables (bytes), in this order at the start of dynamic shared memory
    static constexpr int TW1_ROW = 72;                      // float2 per row of tw1 (row stride = 16 banks mod 32)
    static constexpr int TW1_BYTES = R1 * TW1_ROW * 8;      // tw1[k1][n] = W_M^{n k1}, n < 64
    static constexpr int W64_BYTES = 64 * 8;                // w64[a][b] = W_64^{a b}
    static constexpr int TWH_BYTES = (M + 8) * 8;           // twh[k] = W_N^k, k <= M
    static constexpr int W128_BYTES = (NSPL == 2) ? 64 * 8 : 0;             // w128[n] = W_128^n, n < 64 (frame 4096)
    static constexpr int WIN_BYTES = N * 4;                 // window / synthesis window
    // frame 4096: the split twiddles and the two windows (49 KB) stay in global memory (read through
    // L1, every CTA reads the same lines); with them in shared memory only one CTA fits an SM and
    // nothing overlaps its load phase (measured: 35 % of the roofline against 4x % with two CTAs)
    static constexpr bool GT = (N == 4096);
    // gather middle: 14.3 KB per pair instead of 8.5; two CTAs of seven pairs still fit an SM once the split
    // twiddles are read through L1 (22 loads per thread and call) and the synthesis window is derived from the
    // analysis window (times 1 / (2 N R), a power of two: exact)
    static constexpr bool TWH_GLOBAL = GT || GATHER, WOUT_DERIVED = GT || GATHER;
    static constexpr int TWH_SMEM = TWH_GLOBAL ? 0 : TWH_BYTES, WIN_SMEM = GT ? 0 : WIN_BYTES;
    static constexpr int WOUT_SMEM = WOUT_DERIVED ? 0 : WIN_BYTES;
    // global: NTAB x tw1 | w64 | w128 | twh; shared: ktab | tw1 | w64 | w128 | twh | window | window_out
    static constexpr int OFF_TW1 = PCH ? 0 : DTAB_BYTES;
    static constexpr int OFF_W64 = OFF_TW1 + TW1_BYTES;
    static constexpr int OFF_W128 = OFF_W64 + W64_BYTES;
    static constexpr int OFF_TWH = OFF_W128 + W128_BYTES;
    static constexpr int OFF_WIN = OFF_TWH + TWH_SMEM;
    static constexpr int OFF_WOUT = OFF_WIN + WIN_SMEM;
    static constexpr bool TMEMX = PVB_RING_TMEM && (N_ == 1024);   // exchanges through tensor memory (one warp per pair)
    static constexpr int OFF_MBAR = OFF_WOUT + WOUT_SMEM;    // mbarrier of the bulk table copies (+ 8: tensor-memory base)
    static constexpr int TAB_BYTES = OFF_MBAR + 16;
    static constexpr int MAX_PAIRS = (N == 256) ? 32 : (N == 512) ? 16
                                     : (N == 1024) ? ((PCH && PVB_RING_PAIRS_1024 > 7) ? 7 : PVB_RING_PAIRS_1024)
                                     : (N == 2048) ? (PCH ? 3 : 4) : 2;     // pairs per CTA (two CTAs per SM must fit 227 KB)
    // MULTI kernels (a loop over process() calls around the body) need more than 128 registers per thread
    // to stay out of local memory: three quarters of the pairs per CTA, 168 registers
    static constexpr int MULTI_PAIRS = (N == 4096) ? 2 : (N == 1024) ? PVB_RING_MULTI_PAIRS_1024 : (3 * MAX_PAIRS) / 4;
    // DEEP instances share the launch bounds of MULTI (168 registers where that leaves two CTAs per SM); frame 4096:
    // one pair per CTA, three CTAs per SM at 168 registers (two pairs per CTA would mean 128 registers and spills)
    static constexpr int DEEP_PAIRS = (N == 4096) ? 1 : MULTI_PAIRS;
    // registers per thread of those instances: what two CTAs of MULTI_PAIRS pairs leave (168 at 192 threads, 144 at
    // 224, 128 at 256); the one-call instances stay at 128
    static constexpr int BIG_REGS = ((65536 / (2 * MULTI_PAIRS * TP)) / 8) * 8;
    static constexpr int DEEP_REGS = (N == 4096) ? 168 : BIG_REGS;   // frame 4096: three CTAs of one pair per SM
    static constexpr int CTAS_PER_SM = 2;                   // frame 4096: 30 KB of tables + 2 x 36 KB per CTA
    static constexpr int MAX_WARPS = MAX_PAIRS;             // (frame 1024: one warp per pair)
    static constexpr int MIN_THREADS = 128;                 // trip counts of the staging loops assume this
    static constexpr int MIN_PAIRS = (MIN_THREADS + TP - 1) / TP;
    static constexpr int INVALID_DELTA = 0x3000;            // lands outside [0, nb) from any bin
    // (228 KB per SM, 1 KB of it reserved per resident CTA)
    static_assert(2 * (TAB_BYTES + MAX_PAIRS * PAIR_BYTES + 1024) <= 228 * 1024, "two CTAs per SM must fit the shared memory");
};
using RingGeo = RingGeoT<1024>;

struct RingParams {
    const float *in;            // [C][hop] or nullptr (paused input: zeros, ola:93-100)
    float *out;                 // [C][hop]
    float4 *hist2;              // [pairs][N/2] : (ch0[i], ch1[i], ch0[i+1], ch1[i+1]), ring aligned to t
    float4 *acc2;               // [pairs][N/2] : overlap-add ring, same alignment
    const float *window2;       // [2N] Hann window, twice
    const float *window_out2;   // [2N] window / (2 N R), twice
    const float4 *gtab;         // NJ x tw1 | w64 | twh, built by the host (ring_host_tables)
    int num_channels;
    int hop;
    int num_hops;               // MULTI kernels: consecutive process() calls done by this launch (in / out: [K][C][hop])
    int tmod;                   // timeCursor mod N (multiple of hop) at the first of them
    int stagger_ns;             // -DPVB_EXPERIMENTS only: delay odd warps by this much before the first pass
    int skip;                   // -DPVB_EXPERIMENTS only (PVB_SKIP, results become wrong): bit 0 the whole middle,
                                // bit 1 forward and inverse pass 2
    int early;                  // which state loads may precede griddepcontrol.wait (0, 1, 2; see the kernel)
    // per-pair completion flags: done[pair] holds the sequence number of the last call of this handle
    // whose state / output for that pair is complete (release store at the end of every launch)
    unsigned *done;
    unsigned wait_seq, my_seq;  // this call may touch a pair once done[pair] >= wait_seq; it stores my_seq
    int flag_mode;              // 1: synchronise on done[] per pair instead of waiting for the whole previous grid
    unsigned *err;              // sticky device-error word (mapped host memory): pairs whose flag never arrived
    float pitch_factor;
    int pf_mant, pf_shift;      // pitch_factor == pf_mant * 2^-pf_shift (exact)
    const float *pf_ch;         // PCH kernels: [C] pitch factor per channel, every one in the kernel's range
    // peak guard (see ring_exact_peak_mask): a channel frame is re-decided in float64 when it has at least
    // this many uncertain comparisons (0: always; 0x7fffffff: never, and the test itself is skipped)
    int guard_min;
    const double *xtw;          // [2N] fft.js table: cos(pi i / N), -sin(pi i / N) pairs (bundle:13-17), float64
    const int *xrev;            // [1 << width] fft.js _bitrev (bundle:31-38)
    double *xpool;              // scratch slots of 2N doubles, shared by every handle of this frame size on the device
    unsigned *xlocks;           // one lock word per slot
    int xslots;
    unsigned long long *xcount; // number of channel frames re-decided so far (diagnostics)
    // -DPVB_EXPERIMENTS only: [2] earliest CTA start / latest CTA end of this launch in %globaltimer
    // nanoseconds (profiles/overlap_trace.py: how far consecutive launches overlap), or nullptr
    unsigned long long *stamps;
};

__device__ __forceinline__ float4 pack4(cpx2 v) { return make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
__device__ __forceinline__ cpx2 unpack4(float4 v) { return cpx2{make_float2(v.x, v.y), make_float2(v.z, v.w)}; }

// barrier among the threads of one channel pair (a warp, or two warps on a named barrier)
template <int TP>
__device__ __forceinline__ void pair_sync(int pair_in_cta) {
    if constexpr (TP == 32) {
        __syncwarp();
    } else if constexpr (TP == 16) {
        __syncwarp(0xFFFFu << (threadIdx.x & 16));          // the other half-warp is another pair
    } else if constexpr (TP == 8) {
        __syncwarp(0xFFu << (threadIdx.x & 24));            // four pairs per warp
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(pair_in_cta + 1), "n"(TP) : "memory");
    }
}

#define PVB_COS_PI_8 0.92387953251128673848f
#define PVB_SIN_PI_8 0.38268343236508978178f

// in-register 16-point DFT, natural order in and out: one radix-2 DIF step with W16^n, two dft8
template <bool INV>
__device__ __forceinline__ void dft16(cpx2 (&x)[16]) {
    constexpr float sg = INV ? 1.f : -1.f;                  // sign of the imaginary part of W16^n
    cpx2 a[8], b[8];
#pragma unroll
    for (int n = 0; n < 8; n++) {
        a[n] = cadd(x[n], x[n + 8]);
        b[n] = csub(x[n], x[n + 8]);
    }
    b[1] = cmul_s(b[1], PVB_COS_PI_8, sg * PVB_SIN_PI_8);
    b[2] = mul_w8_1<INV>(b[2]);
    b[3] = cmul_s(b[3], PVB_SIN_PI_8, sg * PVB_COS_PI_8);
    b[4] = mul_mj<INV>(b[4]);
    b[5] = cmul_s(b[5], -PVB_SIN_PI_8, sg * PVB_COS_PI_8);
    b[6] = mul_w8_3<INV>(b[6]);
    b[7] = cmul_s(b[7], -PVB_COS_PI_8, sg * PVB_SIN_PI_8);
    dft8<INV>(a);
    dft8<INV>(b);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        x[2 * k] = a[k];
        x[2 * k + 1] = b[k];
    }
}

// R-point DFT of x[0..R) (R = 2, 4, 8 or 16)
template <int R, bool INV>
__device__ __forceinline__ void dft_r(cpx2 *x) {
    if constexpr (R == 2) {
        const cpx2 a = cadd(x[0], x[1]), b = csub(x[0], x[1]);
        x[0] = a;
        x[1] = b;
    } else if constexpr (R == 4) dft4<INV>(x[0], x[1], x[2], x[3]);
    else if constexpr (R == 8) dft8<INV>(*reinterpret_cast<cpx2(*)[8]>(x));
    else dft16<INV>(*reinterpret_cast<cpx2(*)[16]>(x));
}

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

// one bin of the shifted spectrum (both channels) from the four planes of Y (PLW words per plane)
template <int PLW>
__device__ __forceinline__ cpx2 ring_load_planes(const unsigned char *mine, int slot) {
    const float *y = reinterpret_cast<const float *>(mine) + slot;
    return cpx2{make_float2(y[0], y[PLW]), make_float2(y[2 * PLW], y[3 * PLW])};
}

// Layout of the shifted spectrum Y.  PVB_RING_Y64 = 0 (default): four planes of 32-bit words (re0 | re1 | im0 | im1),
// bin d at word d + (d >> YSHIFT).  1 (`make y64`): two planes (one per channel) of (re, im) pairs at the same slot
// numbers: the scatter and its read-add-store sub-steps move one 64-bit word per bin and channel instead of two
// 32-bit ones (half the load / store instructions there; a model of the lanes' destinations gives 3.5 instead of
// 3.8 wavefronts per (re, im) pair), and the Hermitian pre-pass reads two 64-bit words per bin instead of four
// 32-bit ones, but has to re-pair (re0, im0), (re1, im1) into (re0, re1), (im0, im1).  Measured (parity-green on all
// 475 GPU tests, profiles/r02_ab_y64.txt): 15.92 against 15.94 us per launch at pitch 0.8, 14.44 against 14.21 at
// 1.2 -- the wavefronts, not the instructions, are what the middle costs.  Kept as a measured alternative.
#ifndef PVB_RING_Y64
#define PVB_RING_Y64 0
#endif
template <int XQS>
struct RingY {
    static constexpr int SB = PVB_RING_Y64 ? 8 : 4;      // bytes per slot
    static constexpr int PLB = SB * XQS;                 // bytes per plane
    // `at`: address of the slot in plane 0
    static __device__ __forceinline__ void store(unsigned char *at, int ch, float re, float im) {
        if constexpr (PVB_RING_Y64) {
            *reinterpret_cast<float2 *>(at + ch * PLB) = make_float2(re, im);
        } else {
            *reinterpret_cast<float *>(at + ch * PLB) = re;
            *reinterpret_cast<float *>(at + (2 + ch) * PLB) = im;
        }
    }
    static __device__ __forceinline__ float2 load(const unsigned char *at, int ch) {
        if constexpr (PVB_RING_Y64) {
            return *reinterpret_cast<const float2 *>(at + ch * PLB);
        } else {
            return make_float2(*reinterpret_cast<const float *>(at + ch * PLB), *reinterpret_cast<const float *>(at + (2 + ch) * PLB));
        }
    }
    static __device__ __forceinline__ void atomic_add(unsigned char *at, int ch, float re, float im) {
        atomicAdd(reinterpret_cast<float *>(at + ch * PLB), re);
        atomicAdd(reinterpret_cast<float *>(at + (PVB_RING_Y64 ? ch * PLB + 4 : (2 + ch) * PLB)), im);
    }
    // one bin of both channels for the Hermitian pre-pass
    static __device__ __forceinline__ cpx2 load_bin(const unsigned char *mine, int slot) {
        const float2 a = load(mine + SB * slot, 0), b = load(mine + SB * slot, 1);
        return cpx2{make_float2(a.x, b.x), make_float2(a.y, b.y)};
    }
};

// error model of the float32 forward transform (see "Peak guard" below): |dX_k| <= a |X_k| + c ||X||_2
#define PVB_GUARD_A 3.0e-7f
#define PVB_GUARD_C 5.5e-8f

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

// Both channels at once: the peak masks of ring_peak_mask plus, per channel, the 16-bit mask of
// UNCERTAIN comparisons (see "Peak guard" below): bit e is set when the squared magnitude of bin e and
// the largest of its four neighbours are closer than the float32 transform can tell apart,
//   D^2 <= q (rho q + kappa),  D = m - nbmax,  q = m + nbmax,
// nkappa = -(kappa0, kappa1) per channel (16 c^2 times the summed squared magnitudes of the frame),
// rho = 8 a^2.  Five packed float operations and two funnel shifts per bin.
__device__ __forceinline__ void ring_peak_masks_guarded(const int (&m0)[20], const int (&m1)[20], float2 nkappa,
                                                        uint32_t &mask0, uint32_t &mask1, uint32_t &unc0,
                                                        uint32_t &unc1) {
    constexpr float NRHO = -8.0f * PVB_GUARD_A * PVB_GUARD_A;
    int q0[19], q1[19];
#pragma unroll
    for (int t = 0; t < 19; t++) {
        q0[t] = max(m0[t], m0[t + 1]);
        q1[t] = max(m1[t], m1[t + 1]);
    }
    mask0 = mask1 = unc0 = unc1 = 0;
#pragma unroll
    for (int e = 15; e >= 0; e--) {
        const int nb0 = max(q0[e], q0[e + 3]), nb1 = max(q1[e], q1[e + 3]);   // bins e-2, e-1, e+1, e+2
        mask0 = __funnelshift_l(uint32_t(nb0 - m0[e + 2]), mask0, 1);
        mask1 = __funnelshift_l(uint32_t(nb1 - m1[e + 2]), mask1, 1);
        const float2 c = make_float2(__int_as_float(m0[e + 2]), __int_as_float(m1[e + 2]));
        const float2 nbm = make_float2(__int_as_float(nb0), __int_as_float(nb1));
        const float2 D = sub2(c, nbm), qq = add2(c, nbm);
        const float2 v = mul2(qq, fma2(qq, bc2(NRHO), nkappa));               // -q (rho q + kappa)
        const float2 u = fma2(D, D, v);                                       // < 0 <=> uncertain
        unc0 = __funnelshift_l(uint32_t(__float_as_int(u.x)), unc0, 1);
        unc1 = __funnelshift_l(uint32_t(__float_as_int(u.y)), unc1, 1);
    }
}

// Region of influence of every bin of the run, for one channel (pv:124-141): the owner of a bin
// is the nearest peak, ties go to the higher one.  Peaks travel as KEYS (see the key table in
// the kernel): high half = 2 * (peak + 2048), low half = delta + 32768.  pkey / nkey: the nearest
// peak below / above this thread's run.  Returns per bin the byte offset of its destination word
// inside plane 0 of Y (dump slot when it falls outside [0, nb)), with the sign bit set when the bin
// belongs to the LEFT half of its region while contracting: those are added on top in the second
// sub-step, everything else is stored first (right halves are pairwise disjoint after the shift,
// and so are left halves, for pitch factors >= 0.75).
// The integer pipe runs at half rate, so this loop is written for the fewest ALU operations.
// COLOUR (DEEP instances, pitch factors in [0.5, 0.75)): the low two bits of dst carry the ordinal of the owning
// peak mod 3 instead (col3 = (peaks below this run - 1) mod 3 on entry).  Peaks are at least 3 bins apart, so for
// pitch factors >= 0.5 the images of regions i and i + 3 never overlap: three ordered sub-steps, one per colour,
// have pairwise disjoint destinations each.
template <int DUMP, int YS, bool COLOUR = false>
__device__ __forceinline__ void ring_owner_scan(uint32_t mask, int b0, int pkey, int nkey, const int (&rk)[16],
                                                int second_flag, int (&dst)[16], int col3 = 0) {
    int nx[16];
#pragma unroll
    for (int e = 15; e >= 0; e--) {
        nx[e] = nkey;                                                // first peak above bin e
        if ((mask >> e) & 1u) nkey = rk[e];
    }
    const int thr0 = (4 * (b0 + 2048) + 2) << 16;
    const int cb = b0 - 32768;
    constexpr unsigned YB = PVB_RING_Y64 ? 8u : 4u;                  // bytes per destination slot
#pragma unroll
    for (int e = 0; e < 16; e++) {
        if ((mask >> e) & 1u) {
            pkey = rk[e];                                            // last peak at or below bin e
            if constexpr (COLOUR) col3 = (col3 == 2) ? 0 : col3 + 1;
        }
        // next - b <= b - prev  <=>  2 next' + 2 prev' (+ carry of the low halves) < 4 b' + 2
        const int tt = nx[e] + pkey - thr0;
        const bool take_next = tt < ((4 * e) << 16);
        const int okey = take_next ? nx[e] : pkey;
        const int d = (okey & 0xFFFF) + cb + e;
        const unsigned slot = min(unsigned(d + (d >> YS)), unsigned(DUMP));     // d < 0 or d >= nb: dump slot
        if constexpr (COLOUR) {
            const int cn_ = col3 + (take_next ? 1 : 0);
            dst[e] = int(YB * slot) | (cn_ == 3 ? 0 : cn_);
        } else {
#if PVB_RING_EXACT
        // w = -(T - 1) 2^16 + (delta_next + delta_prev), T = 4 b + 2 - 2 next - 2 prev (even; >= 2 on a
        // left half), deltas <= 0 while contracting  =>  -T = (w >> 16) & ~1; collide <=> T <= 4 c
        const int w = tt - ((4 * e) << 16);
        const int c4 = 4 * int(short(pkey - nx[e]));
        const int q = ((w >> 16) & ~1) + c4;
        dst[e] = int(YB * slot) | (w & ~q & second_flag);
#else
        dst[e] = int(YB * slot) | ((tt - ((4 * e) << 16)) & second_flag);
#endif
        }
    }
}

// Threads of a pair that hold a peak, as WPP ballot words bal[0..WPP): the nearest such thread
// below / above thread tp, and the last one (-1: none)
template <int WPP>
__device__ __forceinline__ int ring_thread_below(const int *bal, int tp) {
    int res = -1;
#pragma unroll
    for (int w = 0; w < WPP; w++) {
        const uint32_t m = uint32_t(bal[w]);
        const int rel = tp - 32 * w;                                  // bits below rel count
        const uint32_t mm = (rel <= 0) ? 0u : (rel >= 32) ? m : (m & ((1u << rel) - 1u));
        if (mm) res = 32 * w + 31 - __clz(mm);
    }
    return res;
}
template <int WPP>
__device__ __forceinline__ int ring_thread_above(const int *bal, int tp) {
    int res = -1;
#pragma unroll
    for (int w = WPP - 1; w >= 0; w--) {
        const uint32_t m = uint32_t(bal[w]);
        const int rel = tp - 32 * w;                                  // bits above rel count
        const uint32_t mm = (rel >= 31) ? 0u : (rel < 0) ? m : (m & ~((2u << rel) - 1u));
        if (mm) res = 32 * w + __ffs(mm) - 1;
    }
    return res;
}
template <int WPP>
__device__ __forceinline__ int ring_thread_last(const int *bal) {
    int res = -1;
#pragma unroll
    for (int w = 0; w < WPP; w++) {
        const uint32_t m = uint32_t(bal[w]);
        if (m) res = 32 * w + 31 - __clz(m);
    }
    return res;
}

// ---------------------------------------------------------------------------------------------------
// Peak guard.  findPeaks (pv:95-116) compares float32 ROUNDINGS OF FLOAT64 squared magnitudes
// (pv:82-92); this kernel computes them from a float32 FFT.  Where a frame has bins near or below the
// float32 round-off floor of its own energy (noise-free tones, digital silence followed by a tone,
// band-limited material) the two disagree about peaks, and one displaced peak changes which stale
// upper bins a contracting shift pulls in: 5.7e-3 RMS on clean tones at pitch factor 0.8.
//
// Detection (every call, every channel, ~3 % of the kernel's instructions): the float32 transform obeys |dX_k| <= a |X_k| + c ||X||_2
// (the classical FFT error bound; a = 3e-7, c = 5.5e-8 hold with a 1.3x margin at every frame size
// for this kernel's decomposition, tests/test_peak_guard_model.py), so the squared magnitudes m carry
// dm <= 2 a m + 2 c sqrt(S m), S = sum of m over the frame, and the comparison of a bin against the
// largest of its four neighbours is UNCERTAIN when  D^2 <= q (8 a^2 q + 16 c^2 S),  D = m - nbmax,
// q = m + nbmax.  A channel with no uncertain comparison has the reference's peak set.
//
// Policy (RingParams::guard_min; PVB_OPT_PEAK_GUARD).  Frames in trouble have CLUSTERS of bins at the
// round-off floor, i.e. tens of uncertain comparisons; a well-conditioned broadband frame has none, or
// (0.6 % of frames at a -20 dB floor) one natural near-tie between two neighbouring bins, which shows
// up as two uncertain comparisons (each bin against the other) and which the float32 decision gets
// right in most cases (measured: tests/test_gpu_peak_guard.py).  The default re-decides every frame
// with FIVE or more uncertain comparisons (clean tones: ~95 % of frames; the benchmark's input: none
// in 4 million frames); "strict" re-decides on one or more -- its outputs are identical to re-deciding
// everything, at the price of ~25 slow channel pairs in every 4096-channel launch, and since every
// launch of the chain has to wait for its slowest pair that costs 30 % of the throughput.
//
// Re-decision (rare; all of a tonal stream): the pair recomputes that channel's squared magnitudes with
// ring_exact_peak_mask below -- fft.js's realTransform restated operation by operation in float64 on
// the windowed frame (same tables, same association order, no fused multiply-add), rounded to float32
// like pv:88 -- so the peak set is the reference's bit for bit.  Everything downstream of the peak set
// is continuous in the spectrum, and stays in float32.
// ---------------------------------------------------------------------------------------------------
// radix-4 stage of fft.js _realTransform4 (bundle:334-441) on `out` (2N doubles, in place): butterfly
// `ii` of the block at `base`.  Literal order of operations; __d*_rn keeps the compiler from fusing.
// Split into the loads and the rest so that a thread can have the operands of several butterflies in
// flight (the scratch slot lives in L2: a dependent round trip per butterfly would dominate).
struct ExactOperands {
    double ar, ai, br, bi, cr, ci, dr, di;
};
__device__ __forceinline__ ExactOperands exact_real_load(const double *out, int base, int ii, int q) {
    const int pa = base + 2 * ii, pb = pa + q, pc = pb + q, pd = pc + q;
    ExactOperands o;
    o.ar = __ldcg(out + pa); o.ai = __ldcg(out + pa + 1);
    o.br = __ldcg(out + pb); o.bi = __ldcg(out + pb + 1);
    o.cr = __ldcg(out + pc); o.ci = __ldcg(out + pc + 1);
    o.dr = __ldcg(out + pd); o.di = __ldcg(out + pd + 1);
    return o;
}
__device__ __forceinline__ void exact_real_finish(double *out, const double *tw, const ExactOperands &o, int base,
                                                  int ii, int step, int q, int h, int hq) {
    const int i = 2 * ii, k = ii * step;
    const int pa = base + i, pb = pa + q, pc = pb + q;
    const double wbr = __ldg(tw + k), wbi = __ldg(tw + k + 1);
    const double wcr = __ldg(tw + 2 * k), wci = __ldg(tw + 2 * k + 1);
    const double wdr = __ldg(tw + 3 * k), wdi = __ldg(tw + 3 * k + 1);
    const double mbr = __dsub_rn(__dmul_rn(o.br, wbr), __dmul_rn(o.bi, wbi)), mbi = __dadd_rn(__dmul_rn(o.br, wbi), __dmul_rn(o.bi, wbr));
    const double mcr = __dsub_rn(__dmul_rn(o.cr, wcr), __dmul_rn(o.ci, wci)), mci = __dadd_rn(__dmul_rn(o.cr, wci), __dmul_rn(o.ci, wcr));
    const double mdr = __dsub_rn(__dmul_rn(o.dr, wdr), __dmul_rn(o.di, wdi)), mdi = __dadd_rn(__dmul_rn(o.dr, wdi), __dmul_rn(o.di, wdr));
    const double s0r = __dadd_rn(o.ar, mcr), s0i = __dadd_rn(o.ai, mci);
    const double s1r = __dsub_rn(o.ar, mcr), s1i = __dsub_rn(o.ai, mci);
    const double s2r = __dadd_rn(mbr, mdr), s2i = __dadd_rn(mbi, mdi);
    const double s3r = __dsub_rn(mbr, mdr), s3i = __dsub_rn(mbi, mdi);        // inv == 1
    __stcg(out + pa, __dadd_rn(s0r, s2r));
    __stcg(out + pa + 1, __dadd_rn(s0i, s2i));
    __stcg(out + pb, __dadd_rn(s1r, s3i));
    __stcg(out + pb + 1, __dsub_rn(s1i, s3r));
    if (i == 0) {                                                              // bundle:400-406
        __stcg(out + pc, __dsub_rn(s0r, s2r));
        __stcg(out + pc + 1, __dsub_rn(s0i, s2i));
        return;
    }
    if (i == hq) return;                                                       // bundle:409-410
    // mirrored outputs, bundle:417-438: ST0 = (s1r, -s1i), ST1 = (s0r, -s0i), ST2 = (-s3i, -s3r), ST3 = (-s2i, -s2r)
    const int sa = base + q - i, sb = base + h - i;
    __stcg(out + sa, __dadd_rn(s1r, -s3i));
    __stcg(out + sa + 1, __dadd_rn(-s1i, -s3r));
    __stcg(out + sb, __dadd_rn(s0r, -s2r));
    __stcg(out + sb + 1, __dsub_rn(-s0i, -s2i));
}

// 16-bit peak mask of this thread's run (bins 16 tp .. 16 tp + 15) of channel `ch`, decided exactly as
// the reference does: float64 realTransform in fft.js's own order, |X|^2 in float64, rounded to float32
// (pv:82-92), 5-point strict maxima (pv:95-116).  Cooperative among the TP threads of the pair; `out`
// is the pair's scratch slot (2N doubles in global memory, L2 resident).
template <int N, int TP>
__device__ __noinline__ uint32_t ring_exact_peak_mask(const float2 *__restrict__ ring /* the pair's history ring [N] */,
                                                      const float *__restrict__ win /* [N] Hann, float32 */,
                                                      const double *__restrict__ tw, const int *__restrict__ rev,
                                                      double *out, int t, int ch, int tp, int pin) {
    constexpr int SIZE = 2 * N;
    constexpr int POWER = (N == 256) ? 8 : (N == 512) ? 9 : (N == 1024) ? 10 : (N == 2048) ? 11 : 12;
    constexpr int WIDTH = (POWER % 2 == 0) ? POWER - 1 : POWER;               // bundle:28
    // windowed sample n of the frame (applyHannWindow, pv:75-79: one correctly rounded float32 product)
    auto xw = [&](int n) -> double {
        const float2 sv = __ldcg(ring + ((n + t) & (N - 1)));
        return double(__fmul_rn(ch ? sv.y : sv.x, __ldg(win + n)));
    };
    if constexpr (POWER % 2 == 0) {
        // _singleRealTransform4 (bundle:468-508): N/4 four-point transforms of the digit-reversed input
#pragma unroll 4
        for (int u = tp; u < N / 4; u += TP) {
            const int off = int(unsigned(__ldg(rev + u)) >> 1);
            const double a = xw(off), b = xw(off + N / 4), c = xw(off + N / 2), d = xw(off + 3 * (N / 4));
            const double s0 = __dadd_rn(a, c), s1 = __dsub_rn(a, c), s2 = __dadd_rn(b, d), s3 = __dsub_rn(b, d);
            double *o = out + 8 * u;
            __stcg(o, __dadd_rn(s0, s2)); __stcg(o + 1, 0.0);
            __stcg(o + 2, s1); __stcg(o + 3, -s3);
            __stcg(o + 4, __dsub_rn(s0, s2)); __stcg(o + 5, 0.0);
            __stcg(o + 6, s1); __stcg(o + 7, s3);
        }
    } else {
        // _singleRealTransform2 (bundle:447-463): N/2 two-point transforms
#pragma unroll 8
        for (int u = tp; u < N / 2; u += TP) {
            const int off = int(unsigned(__ldg(rev + u)) >> 1);
            const double e = xw(off), qv = xw(off + N / 2);
            double *o = out + 4 * u;
            __stcg(o, __dadd_rn(e, qv)); __stcg(o + 1, 0.0);
            __stcg(o + 2, __dsub_rn(e, qv)); __stcg(o + 3, 0.0);
        }
    }
    pair_sync<TP>(pin);
    // (butterflies of one stage touch disjoint slots -- the mirrored outputs land in the half of their
    // sub-block no butterfly of the stage reads -- so they run in any order and in parallel)
#pragma unroll
    for (int step = (1 << WIDTH) >> 2; step >= 2; step >>= 2) {
        const int len = (SIZE / step) << 1, h = len >> 1, q = h >> 1, hq = q >> 1;
        const int cpb = (hq >> 1) + 1;                                        // butterflies per block: i = 0, 2, .. hq
        const int total = (SIZE / len) * cpb;
        constexpr int BATCH = 3;                                              // operand sets in flight per thread
        for (int j0 = tp; j0 < total; j0 += BATCH * TP) {
            ExactOperands ops[BATCH];
            int blk[BATCH], ii[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; u++) {
                const int j = j0 + u * TP;
                blk[u] = j / cpb;
                ii[u] = j - blk[u] * cpb;
                if (j < total) ops[u] = exact_real_load(out, blk[u] * len, ii[u], q);
            }
#pragma unroll
            for (int u = 0; u < BATCH; u++)
                if (j0 + u * TP < total) exact_real_finish(out, tw, ops[u], blk[u] * len, ii[u], step, q, h, hq);
        }
        pair_sync<TP>(pin);
    }
    int m[20];
#pragma unroll
    for (int i = 0; i < 20; i++) {
        int bin = 16 * tp - 2 + i;
        bin = bin < 0 ? 0 : (bin > N / 2 ? N / 2 : bin);
        const double re = __ldcg(out + 2 * bin), im = __ldcg(out + 2 * bin + 1);
        m[i] = __float_as_int(__double2float_rn(__dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im))));   // pv:85-88
    }
    uint32_t mask = ring_peak_mask(m);
    if (tp == 0) mask &= ~3u;                                                 // i >= 2
    if (tp == TP - 1) mask &= ~(1u << 15);                                    // i <= nb - 3
    pair_sync<TP>(pin);                                                       // the slot may be reused (other channel)
    return mask;
}

// ---- tensor memory as an exchange buffer (frame 1024, see PVB_RING_TMEM) ---------------------------------------
__device__ __forceinline__ uint32_t f2u(float x) { return __float_as_uint(x); }
// word w of a packed two-channel complex value: re0, re1, im0, im1
__device__ __forceinline__ uint32_t cpx2_word(const cpx2 &v, int w) {
    return f2u(w == 0 ? v.re.x : w == 1 ? v.re.y : w == 2 ? v.im.x : v.im.y);
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t ta, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(ta), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t ta, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(ta) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t ta, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(ta) : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t ta, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(ta), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Exchange slot (16 bytes) of element e (0 .. 63) of row k1 (0 .. 7) between passes 2 and 3 when pass 2 runs
// with the lanes renumbered: 8 e + a bank term that makes both the stores of pass 2 (a quarter-warp holds
// k1 & 3 = 0 .. 3 and two consecutive e) and the loads of pass 3 (a quarter-warp holds k1 = 0 .. 7 of one e)
// conflict free.  512 slots, no padding.
__device__ __forceinline__ int tmx_slot(int k1, int e) { return 8 * e + ((2 * (k1 & 3) + (k1 >> 2) + (e & 1)) & 7); }

// forward real-split of one (k, M-k) pair into registers: 2 X[k] -> xk, 2 X[M-k] -> xm (both channels)
__device__ __forceinline__ void ring_split_regs(cpx2 za, cpx2 zb, float2 w, cpx2 &xk, cpx2 &xm) {
    const float2 e_r = add2(za.re, zb.re), e_i = sub2(za.im, zb.im);
    const float2 o_r = add2(za.im, zb.im), o_i = sub2(zb.re, za.re);
    const cpx2 tt = cmul_s(cpx2{o_r, o_i}, w.x, w.y);
    xk = cpx2{add2(e_r, tt.re), add2(e_i, tt.im)};
    xm = cpx2{sub2(e_r, tt.re), sub2(tt.im, e_i)};
}

// Gather middle, step C: one descriptor per peak of this thread's run (bits of `mask`, bins b0 .. b0 + 15),
// for one channel.  pkey / nkey: keys (2 (bin + 2048) << 16 | delta + 32768) of the nearest peaks below /
// above the run (pkey 0: none below; nkey: a far-away bin when there is none above); krun[e]: key of bin
// b0 + e.  See ring_one_call for the descriptor semantics.
template <int NB, int TB, int CB, int LRW, int INVALID>
__device__ __forceinline__ void ring_emit_descriptors(uint32_t mask, int b0, int pkey, int nkey, const int *krun,
                                                      bool contract, int *D) {
    constexpr int RW = 1 << LRW, TMAX = (1 << TB) - 2;
    if (!mask) return;
    int e = __ffs(mask) - 1;
    mask &= mask - 1;
    int pos = b0 + e, dl = (krun[e] & 0xFFFF) - 32768;
    int s = 0, dp = contract ? dl : 0;                                 // no peak below: the region starts at bin 0
    if (pkey) {
        const int pp = (pkey >> 17) - 2048;
        dp = (pkey & 0xFFFF) - 32768;
        s = pos - ((pos - pp) >> 1);
    }
    int q = s + min(dl, dp), T = s + max(dl, dp), c = contract ? dp - dl : 0;
#pragma unroll 1
    for (;;) {
        // the next peak: its region starts where this one ends
        const bool more = mask != 0;
        const int en = (__ffs(mask) - 1) & 15;
        mask &= mask - 1;
        const int keyn = more ? krun[en] : nkey;
        const int pn = (keyn >> 17) - 2048, dn = (keyn & 0xFFFF) - 32768;
        const int sn = pn - ((pn - pos) >> 1);
        const int qn = sn + min(dn, dl);
        const int qs = max(q, 0);
        if (qs < NB) {
            const int Tp = min(max(T, 0), TMAX) + 1;
            const int nd = (dl == INVALID) ? 0 : -dl;                  // (an invalid peak only ever zeroes)
            const int word = Tp | (c << TB) | int(unsigned(nd) << (TB + CB));
            D[qs] = word;
            const int lim = min(qn, NB);
            // copies at the bins = 0, 1 (mod RW) inside (qs, lim): most regions have none
            if ((((lim - 1) ^ qs) >> LRW) != 0 || (qs & (RW - 1)) == 0) {
#pragma unroll 1
                for (int x0 = qs & ~(RW - 1); x0 < lim; x0 += RW) {
                    if (x0 > qs) D[x0] = word;
                    if (x0 + 1 > qs && x0 + 1 < lim) D[x0 + 1] = word;
                }
            }
        }
        if (!more) break;
        T = sn + max(dn, dl);
        c = contract ? dl - dn : 0;
        q = qn;
        pos = pn;
        dl = dn;
    }
}

// Gather middle, step D: one bin of the shifted spectrum of channel `ch`.  base = pair buffer + 4 d (d: the
// destination bin of this thread), dp1 = d + 1.  The descriptor in force is the latest one at or below d: the
// lanes of the pair in this warp hold one aligned group of bins, ascending with the lane (MTYPE false) or
// descending (MTYPE true); SELF: the bin is a multiple of RW, where a descriptor always is.
template <int N, bool CONTRACT, bool MTYPE, bool SELF>
__device__ __forceinline__ float2 ring_gather_one(const unsigned char *base, int ch, int dp1, unsigned FULL, unsigned lmask) {
    using G = RingGeoT<N>;
    const int w = *reinterpret_cast<const int *>(base + G::A_BYTES + ch * 4 * G::DW);
    int got = w;
    if (!SELF) {
        const unsigned ball = __ballot_sync(FULL, w != 0) & lmask;
        const int src = MTYPE ? __ffs(ball) - 1 : 31 - __clz(ball);
        got = __shfl_sync(FULL, w, src);
    }
    const float *xs = reinterpret_cast<const float *>(base + ch * 4 * G::XPW) + (got >> (G::TB + G::CB));   // X[d - delta]
    const bool zone = dp1 < (got & ((1 << G::TB) - 1));
    float re, im;
    if (CONTRACT) {
        re = xs[0];
        im = xs[2 * G::XPW];
        if (zone) {                                                    // the previous region reaches this bin too
            const float *x2 = xs - ((got >> G::TB) & ((1 << G::CB) - 1));
            re += x2[0];
            im += x2[2 * G::XPW];
        }
    } else {
        re = im = 0.f;                                                 // gap between two regions
        if (!zone) {
            re = xs[0];
            im = xs[2 * G::XPW];
        }
    }
    return make_float2(re, im);
}

// Gather middle, step D: the shifted spectrum straight into the Hermitian C2R pre-pass (mirror of the split):
// step j takes Y[k] and Y[M - k], k = tp + KS j (thread 0: the bins of its two self-paired butterflies), and
// leaves the inputs of inverse pass 1 in a[] / b[].
template <int N, bool CONTRACT>
__device__ __forceinline__ void ring_gather_unsplit(const unsigned char *mine, const float2 *twh, int tp, int lane,
                                                    unsigned FULL, cpx2 (&a)[8], cpx2 (&b)[8]) {
    using G = RingGeoT<N>;
    constexpr int M = G::M, KS = G::KS;
    const bool l0 = tp == 0;
    const int klo = l0 ? KS / 2 : tp, khi = l0 ? -4 * KS : tp;
    const unsigned lem = (2u << lane) - 1u, gem = ~((1u << lane) - 1u);
    cpx2 zk[8], zmk[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int dk = (j < 4 ? klo : khi) + KS * j, dm = M - dk;
        const unsigned char *bk = mine + 4 * dk, *bm = mine + 4 * dm;
        const float2 k0 = ring_gather_one<N, CONTRACT, false, false>(bk, 0, dk + 1, FULL, lem);
        const float2 k1 = ring_gather_one<N, CONTRACT, false, false>(bk, 1, dk + 1, FULL, lem);
        const float2 q0 = ring_gather_one<N, CONTRACT, true, false>(bm, 0, dm + 1, FULL, gem);
        const float2 q1 = ring_gather_one<N, CONTRACT, true, false>(bm, 1, dm + 1, FULL, gem);
        cpx2 yk{make_float2(k0.x, k1.x), make_float2(k0.y, k1.y)};
        cpx2 ym{make_float2(q0.x, q1.x), make_float2(q0.y, q1.y)};
        if (j == 4) {                            // thread 0: k == 0, bins 0 and N/2 enter with their real part only
            yk.im = make_float2(l0 ? 0.f : yk.im.x, l0 ? 0.f : yk.im.y);
            ym.im = make_float2(l0 ? 0.f : ym.im.x, l0 ? 0.f : ym.im.y);
        }
        const float2 *wp = twh + (j < 4 ? klo : khi) + KS * j;
        const float2 w = G::TWH_GLOBAL ? __ldg(wp) : *wp;
        ring_unsplit(yk, ym, w, zk[j], zmk[j]);
    }
    cpx2 zh, dummy;
    {
        const unsigned char *bh = mine + 4 * (M / 2);
        const float2 h0 = ring_gather_one<N, CONTRACT, false, true>(bh, 0, M / 2 + 1, FULL, 0u);
        const float2 h1 = ring_gather_one<N, CONTRACT, false, true>(bh, 1, M / 2 + 1, FULL, 0u);
        const cpx2 y{make_float2(h0.x, h1.x), make_float2(h0.y, h1.y)};
        const float2 w = G::TWH_GLOBAL ? __ldg(twh + M / 2) : twh[M / 2];
        ring_unsplit(y, y, w, zh, dummy);
    }
    a[0] = sel(l0, zk[4], zk[0]);
    a[1] = sel(l0, zk[5], zk[1]);
    a[2] = sel(l0, zk[6], zk[2]);
    a[3] = sel(l0, zk[7], zk[3]);
    a[4] = sel(l0, zh, zk[4]);
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

// Peak masks of this thread's run for both channels from the squared magnitudes m0 / m1 of bins 16 tp - 2 ..
// 16 tp + 17 (esum: the thread's share of the frame energy per channel), peak guard included (see above).
template <int N, bool PCH>
__device__ __forceinline__ void ring_masks(const RingParams &p, const int (&m0)[20], const int (&m1)[20], float2 esum,
                                           uint32_t &mask0, uint32_t &mask1, unsigned char *mine, int pair, int tp,
                                           int pin, int lane, unsigned FULL, bool has1, int t) {
    using G = RingGeoT<N, PCH>;
    constexpr int TP = G::TP;
    if (p.guard_min == 0x7fffffff) {
        mask0 = ring_peak_mask(m0);
        mask1 = ring_peak_mask(m1);
        if (tp == 0) { mask0 &= ~3u; mask1 &= ~3u; }          // i >= 2
        if (tp == TP - 1) { mask0 &= ~(1u << 15); mask1 &= ~(1u << 15); }       // i <= nb - 3
    } else {
        // ---- peak guard: frame energy per channel, uncertain comparisons, exact re-decision ----
        constexpr int TPW = (TP < 32) ? TP : 32;
#pragma unroll
        for (int off = TPW / 2; off > 0; off >>= 1) {
            esum = add2(esum, make_float2(__shfl_xor_sync(FULL, esum.x, off), __shfl_xor_sync(FULL, esum.y, off)));
        }
        int *gscr = reinterpret_cast<int *>(mine + G::BUF_SLOTS * 16) + 4 * TP + 8;   // multi-warp pairs only
        if constexpr (G::WPP > 1) {
            if (lane == 0) {
                gscr[tp >> 5] = __float_as_int(esum.x);
                gscr[G::WPP + (tp >> 5)] = __float_as_int(esum.y);
            }
            pair_sync<TP>(pin);
            esum = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < G::WPP; w++)
                esum = add2(esum, make_float2(__int_as_float(gscr[w]), __int_as_float(gscr[G::WPP + w])));
        }
        constexpr float KSCALE = -16.0f * PVB_GUARD_C * PVB_GUARD_C;
        uint32_t unc0, unc1;
        ring_peak_masks_guarded(m0, m1, mul2(esum, bc2(KSCALE)), mask0, mask1, unc0, unc1);
        if (tp == 0) { mask0 &= ~3u; mask1 &= ~3u; unc0 &= ~3u; unc1 &= ~3u; }
        if (tp == TP - 1) { mask0 &= ~(1u << 15); mask1 &= ~(1u << 15); unc0 &= ~(1u << 15); unc1 &= ~(1u << 15); }
        // uncertain comparisons of the whole frame, per channel
        int n0 = __reduce_add_sync(FULL, __popc(unc0)), n1 = __reduce_add_sync(FULL, __popc(unc1));
        if constexpr (G::WPP > 1) {
            if (lane == 0) {
                gscr[8 + (tp >> 5)] = n0;
                gscr[8 + G::WPP + (tp >> 5)] = n1;
            }
            pair_sync<TP>(pin);
            n0 = n1 = 0;
#pragma unroll
            for (int w = 0; w < G::WPP; w++) {
                n0 += gscr[8 + w];
                n1 += gscr[8 + G::WPP + w];
            }
        }
        bool redo0 = n0 >= p.guard_min, redo1 = n1 >= p.guard_min;
        redo1 = redo1 && has1;
        if (redo0 | redo1) {
            // a scratch slot of 2N doubles from the pool (at least as many slots as pairs can be
            // resident on the device, so the probe always ends)
            int slot = 0;
            if (tp == 0) {
                slot = int((unsigned(pair) * 2654435761u) % unsigned(p.xslots));
                while (atomicCAS(p.xlocks + slot, 0u, 1u) != 0u) {
                    slot = (slot + 1 == p.xslots) ? 0 : slot + 1;
                    __nanosleep(100);
                }
                __threadfence();
                atomicAdd(p.xcount, (unsigned long long)(int(redo0) + int(redo1)));
            }
            if constexpr (G::WPP > 1) {
                if (tp == 0) gscr[16] = slot;
                pair_sync<TP>(pin);
                slot = gscr[16];
            } else {
                slot = __shfl_sync(FULL, slot, (TP == 32) ? 0 : int(threadIdx.x & (32 - TP)));
            }
            double *xo = p.xpool + size_t(slot) * size_t(2 * N);
            const float2 *ring = reinterpret_cast<const float2 *>(p.hist2 + size_t(pair) * (N / 2));
            if (redo0) mask0 = ring_exact_peak_mask<N, TP>(ring, p.window2, p.xtw, p.xrev, xo, t, 0, tp, pin);
            if (redo1) mask1 = ring_exact_peak_mask<N, TP>(ring, p.window2, p.xtw, p.xrev, xo, t, 1, tp, pin);
            if (tp == 0) {
                __threadfence();
                atomicExch(p.xlocks + slot, 0u);
            }
        }
    }
}

// N = frame size (512: half a warp per pair, 1024: one warp, 2048: two warps, 4096: four warps),
// NBLK = hop / 128.
// Registers of the first / last FFT pass are indexed by FRAME block f (128 samples), so the role
// of every register (history / new input / emitted head / zero tail) is a compile-time fact: no
// predicated duplicates of the global accesses.  Frame block f sits in ring block (f + toff) mod
// NJ, toff = (timeCursor / 128) mod NJ; only the global addresses depend on toff.  The DFT over
// ring blocks of the rotated register array is the DFT over f times W_R1^{toff k1} (shift
// theorem); that factor is folded into the first-pass twiddles, W_M^{(n + 64 toff) k1}: the host
// keeps one such table per toff and the CTA stages the one it needs.
//
// One process() call of one channel pair (the whole body of the kernel).  `hopi`: index of the call
// inside a MULTI launch (0 otherwise); `live`: the pair exists and its state is usable.  Returns `live`
// (false once a completion flag is lost).
// PHASE: 0 the only call of the launch; 1 the first call of a MULTI launch; 2 a later call of a MULTI
// launch (no flags, no waits: the pair itself wrote the state it reads).  In phase 2 the thread indices
// pass through an identity the compiler cannot see through: the body then re-derives its index arithmetic
// at every call instead of hoisting dozens of indices out of the loop and spilling.
// DEEP: pitch factors in [0.5, 0.75).  The shift then reads the stale slots N/2 + q for q up to N/4 - 1 (below
// 0.75 only q < N/8, the first level) and more than two regions can land on one bin.  Slots beyond the first
// level hold outputs of SMALLER sub-transforms of fft.js's radix-4 recursion (bundle:394-438 writes only outputs
// 0 .. L/2 of every length-L block): walking the block tree gives (L, r, s, o) with
//     slot = DFT_L(xw[r m + s])[o] = sum_{m < L} xw[r m + s] W_L^{o m},   r = N / L in {16, 64, 256, 1024},
// a sum of L windowed frame samples that all have n = 10 or 14 (mod 16).  Pass 1 leaves those N/8 samples in a
// per-pair scratch in frame order, the thread that owns slot q sums them (only when the slot lands inside [0, nb)
// after the last peak's shift), turns the result to ring order (times W_N^{(N/2 + q) t}) and adds it; every
// store of the shift becomes a shared-memory atomic add on the zero-filled planes.
template <int N, int NBLK, bool PCH, int PHASE, int DEEP = 0>
__device__ __forceinline__ bool ring_one_call(const RingParams &p, const int hopi, bool live) {
    using G = RingGeoT<N, PCH>;
    static_assert(!DEEP || (PHASE == 0 && !(PVB_RING_GATHER && N == 1024)), "DEEP: one call per launch");
    // DEEP == 1: every pitch factor of the pair >= 0.5 (three coloured sub-steps); 2: down to 0.33 (atomics, and
    // the stale slots of the quarter 3N/4 + o)
    constexpr bool COLOURS = DEEP == 1, WIDE = DEEP == 2;
    constexpr bool MULTI = PHASE != 0;
    constexpr int M = G::M, NB = G::NB, TP = G::TP, R1 = G::R1, KS = G::KS, SS = G::SS, NJ = G::NJ;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int tid = threadIdx.x;
    if constexpr (PHASE == 2) asm volatile("" : "+r"(tid));
    const int lane = tid & 31;
    const int tp = tid % TP;                        // thread within the pair
    const int pin = tid / TP;                       // pair within the CTA
    const int pair = blockIdx.x * (blockDim.x / TP) + pin;
#ifdef PVB_EXPERIMENTS
    const int xskip = p.skip, xstagger = p.stagger_ns;
#else
    constexpr int xskip = 0, xstagger = 0;
#endif
    // lanes of this thread's pair inside its warp (frame 512: half a warp)
    const unsigned FULL = (TP == 16) ? (0xFFFFu << (threadIdx.x & 16))
                          : (TP == 8) ? (0xFFu << (threadIdx.x & 24)) : 0xFFFFFFFFu;
    const float2 *tw1 = reinterpret_cast<const float2 *>(smem_raw + G::OFF_TW1);
    const float2 *w64 = reinterpret_cast<const float2 *>(smem_raw + G::OFF_W64);
    // (frame 4096: twh and the windows are read from global memory, see RingGeoT::GT)
    const float2 *twh = G::TWH_GLOBAL ? reinterpret_cast<const float2 *>(p.gtab + G::NTAB * (G::TW1_BYTES / 16) +
                                                               (G::W64_BYTES + G::W128_BYTES) / 16)
                              : reinterpret_cast<const float2 *>(smem_raw + G::OFF_TWH);
    const float2 *w128 = reinterpret_cast<const float2 *>(smem_raw + G::OFF_W128);   // frame 4096 only
    const float *swin = G::GT ? p.window2 : reinterpret_cast<const float *>(smem_raw + G::OFF_WIN);
    // (frame 4096: the synthesis window is the analysis window times 1 / (2 N R), a power of two, so
    // only one window table competes for L1 there)
    const float *swout = G::WOUT_DERIVED ? swin : reinterpret_cast<const float *>(smem_raw + G::OFF_WOUT);
    constexpr float WOUT_SCALE = G::WOUT_DERIVED ? 1.0f / (2.0f * float(N) * float(N / (NBLK * G::UNIT))) : 1.0f;
#define PVB_TLD2(ptr) (G::GT ? __ldg(ptr) : *(ptr))
#define PVB_TWH(ptr) (G::TWH_GLOBAL ? __ldg(ptr) : *(ptr))
    unsigned char *mine = smem_raw + G::TAB_BYTES + size_t(pin) * G::PAIR_BYTES;
    // key table: what a peak at bin pk contributes to the region scan.  Scalar pitch factor: one table per
    // CTA, entry = position | delta.  PCH: one per pair, entry = delta of channel 0 | delta of channel 1
    // (the position is the index)
    int *ktab = PCH ? reinterpret_cast<int *>(mine + G::BUF_SLOTS * 16 + G::SCR_BYTES) : reinterpret_cast<int *>(smem_raw);
    (void)swout;
    float4 *ex = reinterpret_cast<float4 *>(mine);
    float4 *XQ = reinterpret_cast<float4 *>(mine);
    // DEEP: windowed frame samples n = 10, 14 (mod 16) of both channels, sample n at 2 (n >> 4) + ((n >> 2) & 1);
    // behind the pairs' buffers (the launcher adds DEEP_BYTES per pair to the dynamic shared memory)
    float2 *dsc = reinterpret_cast<float2 *>(smem_raw + G::TAB_BYTES + size_t(blockDim.x / TP) * G::PAIR_BYTES + size_t(pin) * G::DEEP_BYTES);
    (void)dsc;

    const int c0 = 2 * pair;
    const bool has1 = c0 + 1 < p.num_channels;
    const int hop = p.hop;
    constexpr int nblk = NBLK;
    constexpr bool first = PHASE != 2;
    const int t = MULTI ? ((p.tmod + hopi * hop) & (N - 1)) : p.tmod;
    const int toff = (t >> 7) & (NJ - 1);           // ring 128-block of frame block 0
    constexpr bool HB = G::HB;
    const int half = HB ? ((t >> 6) & 1) : 0;       // rings rotated by half a block (frame 256, hop 64)
    const int hh = 32 * half;
    // ring float4 offset of frame block f: 64 ((f + toff) mod NJ) == 64 f + (wraps ? ro_wrap : ro_lin)
    const int ro_lin = 64 * toff, ro_wrap = 64 * toff - 64 * NJ;
#define PVB_RING_OFF(f) (64 * (f) + (((f) + toff >= NJ) ? ro_wrap : ro_lin))

    // Programmatic dependent launch: our CTAs may become resident while the previous kernel on the
    // stream drains.  Two ways to respect what earlier launches wrote:
    //  * flag mode (the host has checked that the caller's buffers do not alias those of recent
    //    launches): dependents are released at once, every pair waits for ITS completion flag only,
    //    and the CTA waits for the previous grid at its very end, so that "this grid is complete"
    //    still implies "everything before it is complete".  Calls of different handles, and
    //    different pairs of one handle, then overlap freely: the load phase of one CTA runs under
    //    the compute phase of its SM neighbour.
    //  * grid mode: griddepcontrol.wait before the first dependent access; everything up to it
    //    touches only constant tables (and state that is provably older than the previous kernel).
    if (first && p.flag_mode) asm volatile("griddepcontrol.launch_dependents;");
    if (MULTI && !first) __syncthreads();           // every pair is done with the previous hop's first-pass twiddles
    // ---- CTA-shared tables: asynchronous 16-byte copies, fixed trip counts (CTAs have at least
    // MIN_THREADS threads; no division by blockDim) --------------------------------------------------
    {
        const float4 *w1 = reinterpret_cast<const float4 *>(p.window2);
        const float4 *w2 = reinterpret_cast<const float4 *>(p.window_out2);
        const unsigned s_tab = unsigned(__cvta_generic_to_shared(smem_raw + G::OFF_TW1));
        const unsigned s_rest = unsigned(__cvta_generic_to_shared(smem_raw + G::OFF_W64));
        const unsigned s_win = unsigned(__cvta_generic_to_shared(smem_raw + G::OFF_WIN));
        const unsigned s_wout = unsigned(__cvta_generic_to_shared(smem_raw + G::OFF_WOUT));
        const float4 *g_tw1 = p.gtab + (HB ? 2 * toff + half : toff) * (G::TW1_BYTES / 16);   // the table of this offset
        const float4 *g_rest = p.gtab + G::NTAB * (G::TW1_BYTES / 16);        // w64 | twh
#if PVB_RING_BULK_TABLES
        // one thread issues the copies; they complete on the mbarrier everyone waits on after the frame loads
        constexpr unsigned REST_BYTES = G::W64_BYTES + G::W128_BYTES + G::TWH_SMEM;
        if (threadIdx.x == 0) {
            const unsigned mb = unsigned(__cvta_generic_to_shared(smem_raw + G::OFF_MBAR));
            if (first) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            // (later hops of a MULTI launch: the previous table was read through the generic proxy)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const unsigned bytes = G::TW1_BYTES + (first ? REST_BYTES + G::WIN_SMEM + G::WOUT_SMEM : 0u);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s_tab), "l"(g_tw1), "n"(G::TW1_BYTES), "r"(mb) : "memory");
            if (first) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s_rest), "l"(g_rest), "n"(REST_BYTES), "r"(mb) : "memory");
                if constexpr (G::WIN_SMEM > 0)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s_win), "l"(w1), "n"(G::WIN_SMEM), "r"(mb) : "memory");
                if constexpr (G::WOUT_SMEM > 0)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s_wout), "l"(w2), "n"(G::WOUT_SMEM), "r"(mb) : "memory");
            }
        }
#else
#pragma unroll
        for (int k = 0; k < (G::TW1_BYTES / 16 + G::MIN_THREADS - 1) / G::MIN_THREADS; k++) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < G::TW1_BYTES / 16)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s_tab + 16 * i), "l"(g_tw1 + i));
        }
        // (everything below is the same for every hop of the launch)
#pragma unroll
        for (int k = 0; k < ((G::W64_BYTES + G::W128_BYTES + G::TWH_SMEM) / 16 + G::MIN_THREADS - 1) / G::MIN_THREADS; k++) {
            const int i = threadIdx.x + k * blockDim.x;
            if (first && i < (G::W64_BYTES + G::W128_BYTES + G::TWH_SMEM) / 16)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s_rest + 16 * i), "l"(g_rest + i));
        }
#pragma unroll
        for (int k = 0; k < (G::WIN_SMEM / 16 + G::MIN_THREADS - 1) / G::MIN_THREADS; k++) {
            const int i = threadIdx.x + k * blockDim.x;
            if (first && i < G::WIN_SMEM / 16) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s_win + 16 * i), "l"(w1 + i));
                if constexpr (G::WOUT_SMEM > 0)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s_wout + 16 * i), "l"(w2 + i));
            }
        }
#endif
        // delta = round(pk * pitchFactor) - pk (pv:125-127): the product of a 10..12-bit integer and a float32 is
        // exact in float64, and Math.round is floor(x + 0.5)
        if (!PCH && first) {
            const double pfd = double(p.pitch_factor);
#pragma unroll
            for (int k = 0; k < (NB + 1 + G::MIN_THREADS - 1) / G::MIN_THREADS; k++) {
                const int pk = threadIdx.x + k * blockDim.x;
                if (pk <= NB) {
                    const int ps = __double2int_rd(fma(double(pk), pfd, 0.5));
                    const int delta = (ps <= NB) ? ps - pk : G::INVALID_DELTA;
                    ktab[pk + 4 * (pk >> 4)] = ((2 * (pk + 2048)) << 16) | (delta + 32768);
                }
            }
        }
    }
    // ---- frame loads: all issued before anything consumes them ---------------------------------
    // Element e of a thread: first-pass butterfly h = e / R1, frame block f = e % R1, ring float4
    // index tp + 32 h + PVB_RING_OFF(f).  The new block (f >= NJ - NBLK) is loaded as (ch0 pair, ch1
    // pair) and interleaved after the wait (moves must not sit between the loads).
    // Grid mode only: history written by launches OLDER than the kernel in front of us on the stream
    // is already complete when our CTAs start (that kernel passed its own griddepcontrol.wait before
    // it let us launch), so those loads are issued before our wait:
    //   p.early == 2: none of this handle's state was written by the previous kernel -> all of hist
    //   p.early == 1: the previous kernel may be this handle's last call -> all but its newest block
    // The input block always waits (it belongs to the caller's stream order).
    constexpr int TPH = (TP < 32) ? TP : 32;        // float4 stride between the first-pass butterflies of a thread
    constexpr int RT = G::RT, NSPL = G::NSPL;
    static_assert(NBLK % NSPL == 0, "frame 4096: the hop must cover an even number of 128-sample blocks");
    // register e of a thread: first-pass butterfly h, role index fr (history / new input / emitted head
    // / zero tail are compile-time facts of fr), frame block f.  Frame 4096: column cn = tp mod 64,
    // parity sp = tp / 64, f = 2 fr + sp.
    const int cn = (NSPL == 2) ? (tp & 63) : tp;
    const int sp = (NSPL == 2) ? (tp >> 6) : 0;
    // roles in units of G::UNIT samples: role(e) = fr, or 2 fr + (second half of the block) for frame 256
    constexpr int NROLE = HB ? 2 * RT : RT;
    constexpr int HSPLIT = HB ? 32 / TPH : 1;       // butterflies h >= HSPLIT sit in the second half of their block
    constexpr int NEWFROM = NROLE - NBLK / NSPL;    // role >= NEWFROM: new input
    constexpr int OLDTO = NROLE - 2 * NBLK / NSPL;  // role < OLDTO: written before the previous call
    constexpr int HEADTO = NBLK / NSPL;             // role < HEADTO: emitted
#define PVB_FR(e) ((NSPL == 2) ? (e) : (e) % RT)
#define PVB_FH(e) ((NSPL == 2) ? 0 : (e) / RT)
#define PVB_FB(fr) ((NSPL == 2) ? 2 * (fr) + sp : (fr))
#define PVB_ROLE(h, fr) (HB ? 2 * (fr) + (((h) >= HSPLIT) ? 1 : 0) : (fr))
    // ring float4 index (relative to hl / al) of butterfly h, frame block f
#define PVB_RING_IDX(h, f)                                                                               \
    (HB ? (((cn + TPH * (h) + hh) & 63) + 64 * (((f) + toff + (((h) >= HSPLIT) ? half : 0)) & (NJ - 1))) \
        : (TPH * (h) + PVB_RING_OFF(f)))
    // ring column of butterfly h (exchange slot, first-pass twiddle)
#define PVB_COL(h) (HB ? ((cn + TPH * (h) + hh) & 63) : (cn + TPH * (h)))
    float4 r[16];
    float4 *hl = p.hist2 + size_t(live ? pair : 0) * (N / 2) + (HB ? 0 : cn);
    const int early = (p.flag_mode || !first) ? 0 : p.early;
    if (!first) {
        // later hops of a MULTI launch: the pair itself wrote the state it reads
    } else if (p.flag_mode) {
        if (live) {
            // acquire: the previous call of this handle has finished with this pair.  The spin is bounded:
            // when the flag does not arrive (time slicing, a debugger, a launch that never ran) the pair
            // falls back to a real ordering -- it waits for the whole previous grid -- and looks again.
            // A flag that is still missing then means the previous call never completed this pair: the
            // pair is not processed (stale state must not be built upon) and the sticky error word makes
            // every later entry point on the handle fail.
            unsigned v;
            int it = 0;
            bool ok = true;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.done + pair) : "memory");
                if (int(v - p.wait_seq) >= 0) break;
                if (++it > 200000) {
                    asm volatile("griddepcontrol.wait;" ::: "memory");
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.done + pair) : "memory");
                    if (int(v - p.wait_seq) < 0) {
                        ok = false;
                        if (tp == 0) atomicAdd(p.err, 1u);
                    }
                    break;
                }
                __nanosleep(200);
            }
            if constexpr (TP < 32) pair_sync<TP>(pin); else __syncwarp();
            live = ok;
        }
    } else if (live && early) {
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int h = PVB_FH(e), fr = PVB_FR(e), f = PVB_FB(fr);
            // fr >= NEWFROM: new input; the blocks below down to OLDTO: what the previous call wrote
            const int role = PVB_ROLE(h, fr);
            if (role < NEWFROM && (role < OLDTO || early == 2)) r[e] = hl[PVB_RING_IDX(h, f)];
        }
        if (early == 2) {
            // warm L2 with the overlap-add ring lines the tail of this kernel adds to
            const int line = 16 * tp;                                 // float4 index: 256 bytes per thread
            if (HB || (((line >> 6) - toff) & (NJ - 1)) < NJ - nblk) {
                const float4 *ap = p.acc2 + size_t(pair) * (N / 2) + line;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ap));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + 8));
            }
        }
    }
    if (first && !p.flag_mode) {
        // our dependents may launch only now: whoever starts behind us can rely on everything older
        // than us being complete
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    if (live) {
        const float *i0 = p.in ? p.in + (size_t(hopi) * p.num_channels + c0) * hop + 2 * cn : nullptr;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int h = PVB_FH(e), fr = PVB_FR(e), f = PVB_FB(fr);
            const int role = PVB_ROLE(h, fr);
            if (role >= NEWFROM) {
                float2 u0 = make_float2(0.f, 0.f), u1 = make_float2(0.f, 0.f);
                if (i0) {
                    // frame sample 2 (cn + TPH h) + 128 f, minus the N - hop samples of history
                    const int so = 2 * TPH * h + 128 * f - (N - NBLK * G::UNIT);
                    u0 = __ldg(reinterpret_cast<const float2 *>(i0 + so));
                    if (has1) u1 = __ldg(reinterpret_cast<const float2 *>(i0 + hop + so));
                }
                r[e] = make_float4(u0.x, u0.y, u1.x, u1.y);           // interleaved below
            } else if (!(early && (role < OLDTO || early == 2))) {
                r[e] = hl[PVB_RING_IDX(h, f)];
            }
        }
    }
    float pfc0 = p.pitch_factor, pfc1 = p.pitch_factor;               // pitch factor of each channel of the pair
    if constexpr (PCH) {
        // (placed here so that the integer work runs under the frame loads issued above)
        if (live && first) {
            pfc0 = __ldg(p.pf_ch + 2 * pair);
            pfc1 = (2 * pair + 1 < p.num_channels) ? __ldg(p.pf_ch + 2 * pair + 1) : 1.0f;
            // float32 == mant * 2^-shift exactly (the host admits only normal values in the kernel's range)
            const unsigned b0 = __float_as_uint(pfc0), b1 = __float_as_uint(pfc1);
            const long long m0 = (long long)((b0 & 0x7FFFFFu) | 0x800000u), m1 = (long long)((b1 & 0x7FFFFFu) | 0x800000u);
            const int s0 = 150 - int((b0 >> 23) & 0xFFu), s1 = 150 - int((b1 >> 23) & 0xFFu);
            const long long h0 = 1ll << (s0 - 1), h1 = 1ll << (s1 - 1);
            for (int pk = tp; pk <= NB; pk += TP) {
                const long long ps0 = (m0 * pk + h0) >> s0, ps1 = (m1 * pk + h1) >> s1;
                const int d0 = (ps0 <= NB) ? int(ps0) - pk : G::INVALID_DELTA;
                const int d1 = (ps1 <= NB) ? int(ps1) - pk : G::INVALID_DELTA;
                ktab[pk + 4 * (pk >> 4)] = (d0 + 32768) | ((d1 + 32768) << 16);
            }
        }
    }
#if PVB_RING_BULK_TABLES
    __syncthreads();            // the key table is complete (and the mbarrier is initialised)
    {
        const unsigned mb = unsigned(__cvta_generic_to_shared(smem_raw + G::OFF_MBAR));
        const unsigned parity = MULTI ? unsigned(hopi & 1) : 0u;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "PVB_TAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra PVB_TAB_DONE;\n"
            "bra PVB_TAB_WAIT;\n"
            "PVB_TAB_DONE:\n"
            "}\n" ::"r"(mb), "r"(parity) : "memory");
    }
#else
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
#endif
    // this warp's window of the CTA's tensor memory: its 32 lanes, 64 columns (allocated in the kernel's prologue)
    uint32_t tmx = 0;
    if constexpr (G::TMEMX) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t w = threadIdx.x >> 5;
        tmx = *reinterpret_cast<const volatile uint32_t *>(smem_raw + G::OFF_MBAR + 8) + ((32u * (w & 3u)) << 16) + 64u * (w >> 2);
    }
    (void)tmx;
    if (!live) return false;    // no CTA-wide barriers below (MULTI: none before the next call's)

    if (xstagger > 0 && (pin & 1)) __nanosleep(unsigned(xstagger));
    // the new block joins the history ring (ola:105)
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int h = PVB_FH(e), fr = PVB_FR(e), f = PVB_FB(fr);
        if (PVB_ROLE(h, fr) >= NEWFROM) {
            r[e] = make_float4(r[e].x, r[e].z, r[e].y, r[e].w);       // (ch0[i], ch1[i], ch0[i+1], ch1[i+1])
            hl[PVB_RING_IDX(h, f)] = r[e];
        }
    }

    // ---- Hann window (pv:55) + forward pass 1: butterflies n = tp (+ 32) over the frame blocks ---
    {
        const float *wl = swin + 2 * cn;
        const int row0 = RT * sp;                                     // frame 4096: rows (s, k') = 16 s + k'
#pragma unroll
        for (int h = 0; h < G::NB1; h++) {
            const int nl = PVB_COL(h);
            cpx2 x[RT];
#pragma unroll
            for (int j = 0; j < RT; j++) {
                const float2 w = PVB_TLD2(reinterpret_cast<const float2 *>(wl + 2 * TPH * h + 128 * PVB_FB(j)));
                const float4 v = r[RT * h + j];
                x[j].re = mul2(make_float2(v.x, v.y), bc2(w.x));
                x[j].im = mul2(make_float2(v.z, v.w), bc2(w.y));
            }
            if constexpr (DEEP) {
                if (h == 0) {
                    float2 *t16 = dsc + N / 8;
                    const float2 wv = PVB_TWH(twh + 16 * tp);
                    t16[tp] = wv;
                    if constexpr (!G::T16_HALF) t16[tp + TP] = make_float2(-wv.x, -wv.y);     // W_N^{16 (tp + TP)} = -W_N^{16 tp}
                }
                if constexpr (HB) {
                    // frame 256: the rings may be rotated by half a block; frame sample of a ring position at run time
#pragma unroll
                    for (int j = 0; j < RT; j++) {
                        const int n = (2 * PVB_RING_IDX(h, PVB_FB(j)) - t) & (N - 1);
                        if (((n >> 1) & 5) == 5) dsc[2 * (n >> 4) + ((n >> 2) & 1)] = x[j].re;
                    }
                } else {
                    // x[j].re is the windowed frame sample 128 f + 2 c of both channels, c = float4 column of the block
                    const int c = cn + TPH * h;
                    if ((c & 5) == 5) {
#pragma unroll
                        for (int j = 0; j < RT; j++) dsc[16 * PVB_FB(j) + 2 * (c >> 3) + ((c >> 1) & 1)] = x[j].re;
                    }
                }
            }
            dft_r<RT, false>(x);
#pragma unroll
            for (int k1 = 1; k1 < RT; k1++) {
                const float2 w = tw1[G::TW1_ROW * (row0 + k1) + nl];  // W_M^{(n + 64 toff) k1} (W_32^{s k1})
                x[k1] = cmul_s(x[k1], w.x, w.y);
            }
            if constexpr (G::TMEMX) {
                // lane = this thread, column = (w & 1) | (k1 & 3) << 1 | (w >> 1) << 3 | h << 4 | (k1 >> 2) << 5
                // for word w of x[k1] of butterfly h: pass 2 then finds k1 & 3 in its lane number
#pragma unroll
                for (int k1hi = 0; k1hi < 2; k1hi++) {
                    uint32_t v[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = cpx2_word(x[((i >> 1) & 3) + 4 * k1hi], (i & 1) | ((i >> 3) << 1));
                    tmem_st_32x32b_x16(tmx + 16 * h + 32 * k1hi, v);
                }
            } else {
#pragma unroll
            for (int k1 = 0; k1 < RT; k1++) ex[G::RS * (row0 + k1) + nl + (G::G8 - 8) * (nl >> 3)] = pack4(x[k1]);
            }
        }
    }
    if constexpr (!G::TMEMX) pair_sync<TP>(pin);

    // warm L2 with the overlap-add ring lines the tail of this kernel adds to (the slot that is
    // only written, ring [t - hop, t), is skipped)
    {
        const int line = 16 * tp;                                     // float4 index: 256 bytes per thread
        if (early != 2 && (HB || (((line >> 6) - toff) & (NJ - 1)) < NJ - nblk)) {
            const float4 *ap = p.acc2 + size_t(pair) * (N / 2) + line;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ap));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + 8));
        }
    }

    // ---- forward pass 2: butterflies (k1, m3) over m2, in place -----------------------------------
    const int m3l = tp & 7;
    if (!(xskip & 2)) {
        float2 w2[8];
        if constexpr (!G::TMEMX) {
#pragma unroll
        for (int k2 = 1; k2 < 8; k2++) w2[k2] = w64[8 * k2 + m3l];              // W_64^{m3 k2}
        }
        if constexpr (NSPL == 2) {
            // rows k' and 16 + k' hold E_0 and E_1 (twiddled); X[k'] = E_0 + E_1 and
            // X[k' + 16] = (E_0 - E_1) W_128^n (-1)^toff, n = m3 + 8 m2; both rows are this thread's
            float4 *b0 = ex + G::RS * (tp >> 3) + m3l, *b1 = b0 + G::RS * 16;
            const float sgn = (toff & 1) ? -1.f : 1.f;
            cpx2 x0[8], x1[8];
#pragma unroll
            for (int m2 = 0; m2 < 8; m2++) {
                const cpx2 e0 = unpack4(b0[G::G8 * m2]), e1 = unpack4(b1[G::G8 * m2]);
                const float2 w = w128[m3l + 8 * m2];
                x0[m2] = cadd(e0, e1);
                x1[m2] = cmul_s(csub(e0, e1), sgn * w.x, sgn * w.y);
            }
            dft8<false>(x0);
            dft8<false>(x1);
#pragma unroll
            for (int k2 = 1; k2 < 8; k2++) {
                x0[k2] = cmul_s(x0[k2], w2[k2].x, w2[k2].y);
                x1[k2] = cmul_s(x1[k2], w2[k2].x, w2[k2].y);
            }
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) {
                b0[G::G8 * k2] = pack4(x0[k2]);
                b1[G::G8 * k2] = pack4(x1[k2]);
            }
        } else if constexpr (G::TMEMX) {
            // lanes renumbered for this pass: t = 4 m3 + a does the butterflies (k1 = a + 4 h, m3) over m2.
            // Load s of the two gives register 4 g + 2 b0 + w0 = column 2 a + w0 + 8 g of lane m3 + 8 (b0 + 2 s):
            // word w0 | (g & 1) << 1 of x[k1 = a + 4 (g >> 2)] of pass-1 thread m3 + 8 (m2 & 3), butterfly m2 >> 2
            // = (g >> 1) & 1, with m2 & 3 = b0 + 2 s.
            const int ta_ = tp & 3, tm3 = tp >> 2;
            uint32_t v0[32], v1[32];
            tmem_wait_st();
            tmem_ld_16x256b_x8(tmx, v0);
            tmem_ld_16x256b_x8(tmx + (16u << 16), v1);
            tmem_wait_ld();
#pragma unroll
            for (int k2 = 1; k2 < 8; k2++) w2[k2] = w64[8 * k2 + tm3];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                cpx2 x[8];
#pragma unroll
                for (int m2 = 0; m2 < 8; m2++) {
                    const uint32_t *vv = ((m2 >> 1) & 1) ? v1 : v0;
                    const int g0 = ((m2 >> 2) << 1) | (h << 2), i0 = 2 * (m2 & 1);
                    x[m2].re = make_float2(__uint_as_float(vv[4 * g0 + i0]), __uint_as_float(vv[4 * g0 + i0 + 1]));
                    x[m2].im = make_float2(__uint_as_float(vv[4 * (g0 | 1) + i0]), __uint_as_float(vv[4 * (g0 | 1) + i0 + 1]));
                }
                dft8<false>(x);
#pragma unroll
                for (int k2 = 1; k2 < 8; k2++) x[k2] = cmul_s(x[k2], w2[k2].x, w2[k2].y);
                // row k1 = a + 4 h, element m3 + 8 k2
                float4 *bp = ex + 8 * tm3 + ((2 * ta_ + h + (tm3 & 1)) & 7);
#pragma unroll
                for (int k2 = 0; k2 < 8; k2++) bp[64 * k2] = pack4(x[k2]);
            }
        } else {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float4 *bp = ex + G::RS * ((tp >> 3) + (R1 / 2) * h) + m3l;
            cpx2 x[8];
#pragma unroll
            for (int m2 = 0; m2 < 8; m2++) x[m2] = unpack4(bp[G::G8 * m2]);
            dft8<false>(x);
#pragma unroll
            for (int k2 = 1; k2 < 8; k2++) x[k2] = cmul_s(x[k2], w2[k2].x, w2[k2].y);
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) bp[G::G8 * k2] = pack4(x[k2]);
        }
        }
    }
    pair_sync<TP>(pin);

    // ---- forward pass 3: butterflies A (bins tp + KS j) and B (bins KS - tp + KS j) ----------------
    // thread 0 owns the two self-paired butterflies: A = bins KS j, B = bins KS/2 + KS j
    const bool l0 = tp == 0;
    const int kB = l0 ? KS / 2 : KS - tp;
    const int exA = G::RS * (tp & (R1 - 1)) + G::G8 * (tp >> G::LR1);
    const int exB = G::RS * (kB & (R1 - 1)) + G::G8 * (kB >> G::LR1);
    cpx2 a[8], b[8];
    if constexpr (G::TMEMX) {
        // row k1 = k & 7, elements 8 (k >> 3) + c of the renumbered pass 2 (tmx_slot)
#pragma unroll
        for (int c = 0; c < 8; c++) {
            a[c] = unpack4(ex[tmx_slot(tp & 7, 8 * (tp >> 3) + c)]);
            b[c] = unpack4(ex[tmx_slot(kB & 7, 8 * (kB >> 3) + c)]);
        }
    } else {
#pragma unroll
    for (int c = 0; c < 8; c++) {
        a[c] = unpack4(ex[exA + c]);
        b[c] = unpack4(ex[exB + c]);
    }
    }
    dft8<false>(a);      // a[j] = Z[tp + KS j]
    dft8<false>(b);      // b[j] = Z[kB + KS j]
    pair_sync<TP>(pin);  // everyone has read the exchange slots: X may overwrite them

    if constexpr (G::GATHER) {
    // =====================================================================================================
    // Gather middle.  Regions of influence in DESTINATION space: region i (peak p_i, delta_i, sources
    // [s_i, s_{i+1}), s_i = p_i - floor((p_i - p_{i-1}) / 2), s_0 = 0, pv:132-141) lands on
    // [s_i + delta_i, s_{i+1} + delta_i).  With q_i = s_i + min(delta_i, delta_{i-1}) and
    // T_i = s_i + max(delta_i, delta_{i-1}) the destination axis is cut at the q_i, and inside [q_i, q_{i+1})
    //     d <  T_i :  contracting: Y[d] = X[d - delta_i] + X[d - delta_{i-1}]   (two regions overlap)
    //                 expanding:   Y[d] = 0                                      (gap between two regions)
    //     d >= T_i :  Y[d] = X[d - delta_i]
    // (pitch factors >= 0.75: never more than two regions on one bin).  A peak whose shifted position is
    // beyond nb (pv:127, only while expanding) has a huge delta: the same formulas give q = end of the previous
    // image and T = "never", i.e. zeros to the end, and later peaks fall outside [0, nb).
    //  A. the split writes X as four planes of words (the first stale level, bundle:394-438, behind bin M:
    //     all four operands of a stale slot sit in one thread) and |X|^2 as (ch0, ch1) pairs;
    //  B. every thread reads the magnitudes of its run of 16 bins (+ 2 on each side): peak masks, peak guard;
    //  C. every thread walks the peaks of its run and stores one 32-bit descriptor per peak
    //     (T' | c << TB | -delta << (TB + CB), c = delta_{i-1} - delta_i) at D[max(q_i, 0)], and copies of it at
    //     every bin = 0 or 1 (mod RW) inside (q_i, q_{i+1}), RW = lanes of a pair in one warp: every aligned
    //     group of RW destination bins then has a descriptor at its first and at its second bin;
    //  D. the Hermitian pre-pass takes Y[tp + KS j] and Y[M - tp - KS j] in step j: the RW lanes cover one
    //     aligned group (thread 0's bins are multiples of RW, where a descriptor always is), so the latest
    //     descriptor at or below a bin is one ballot and one shuffle away, nothing is carried from step to
    //     step, and the shifted spectrum goes from X to registers without ever existing in shared memory.
    // tests/ring_kernel_model.py::gather_rows_kernel is the numpy blueprint of these four steps.
    // =====================================================================================================
    constexpr int XPW = G::XPW, TB = G::TB, CB = G::CB, RW = 1 << G::LRW;
    float *xp = reinterpret_cast<float *>(mine);                                  // planes re0 | re1 | im0 | im1
    float2 *mg = reinterpret_cast<float2 *>(mine + G::A_BYTES);                   // |X|^2, bin k at unit k + (k >> 4) + 3
    int *D0 = reinterpret_cast<int *>(mine + G::A_BYTES), *D1 = D0 + G::DW;       // descriptors (over the magnitudes)
    const bool contract = p.pitch_factor < 1.0f;
    const int klo = l0 ? KS / 2 : tp, khi = l0 ? -4 * KS : tp;                    // bin of slot j: klo / khi + KS j
    const int tlo = klo, thi = khi;
    {
        // ---- A: real split in registers -> X planes, magnitudes, stale bins -----------------------------------
        float *xk_lo = xp + klo, *xk_hi = xp + khi;                               // + KS j
        float *xm_lo = xp + (M - klo), *xm_hi = xp + (M - khi);                   // - KS j
        float2 *uk_lo = mg + (klo + (klo >> 4) + 3), *uk_hi = mg + (khi + (khi >> 4) + 3);           // + SS j
        float2 *um_lo = mg + ((M - klo) + ((M - klo) >> 4) + 3), *um_hi = mg + ((M - khi) + ((M - khi) >> 4) + 3);   // - SS j
        auto put = [&](float *xw, float2 *mu, const cpx2 &v) {
            xw[0] = v.re.x; xw[XPW] = v.re.y; xw[2 * XPW] = v.im.x; xw[3 * XPW] = v.im.y;
            *mu = fma2(v.re, v.re, mul2(v.im, v.im));                             // pv:82-92, float32
        };
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const int j4 = jj + 4;
            cpx2 k0, m0v, k4, m4;
            {
                const cpx2 za = sel(l0, b[jj], a[jj]);
                ring_split_regs(za, b[7 - jj], PVB_TWH(twh + tlo + KS * jj), k0, m0v);
            }
            {
                const cpx2 za = sel(l0, a[jj], a[j4]);
                const cpx2 zb = sel(l0, a[(12 - j4) & 7], b[7 - j4]);
                ring_split_regs(za, zb, PVB_TWH(twh + thi + KS * j4), k4, m4);
            }
            put(xk_lo + KS * jj, uk_lo + SS * jj, k0);
            put(xm_lo - KS * jj, um_lo - SS * jj, m0v);
            put(xk_hi + KS * j4, uk_hi + SS * j4, k4);
            put(xm_hi - KS * j4, um_hi - SS * j4, m4);
            if (contract) {
                // first stale level (what _realTransform4 leaves in slot N/2 + q, rebuilt from the valid half):
                // S[q] = ((X[q] - X[N/4 + q]) + conj(X[M - q] - X[N/4 - q])) conj(W_N^{2q}) / 4.  Slots jj and
                // jj + 4 of a thread hold bins k, k + N/4, M - k, N/4 - k: q = tp, tp + KS (from k) and
                // 2 KS - tp, KS - tp (from M - k).  Thread 0's bins pair differently (see below).
                const bool fromk = jj < 2;
                const int q = (jj == 0) ? tp : (jj == 1) ? tp + KS : (jj == 2) ? 2 * KS - tp : KS - tp;
                const cpx2 &A = fromk ? k0 : m4, &Bv = fromk ? k4 : m0v, &Cv = fromk ? m0v : k4, &Dv = fromk ? m4 : k0;
                const float2 sr = add2(sub2(A.re, Bv.re), sub2(Cv.re, Dv.re));
                const float2 si = sub2(sub2(A.im, Bv.im), sub2(Cv.im, Dv.im));
                const float2 w = PVB_TWH(twh + 2 * q);
                const cpx2 sv = cmul_s(cpx2{sr, si}, 0.25f * w.x, -0.25f * w.y);
                if (!l0) {
                    float *xe = xp + M + q;
                    xe[0] = sv.re.x; xe[XPW] = sv.re.y; xe[2 * XPW] = sv.im.x; xe[3 * XPW] = sv.im.y;
                }
            }
        }
        if (l0) {
            cpx2 hk, hm;
            ring_split_regs(a[4], a[4], PVB_TWH(twh + M / 2), hk, hm);
            put(xp + M / 2, mg + (M / 2 + M / 32 + 3), hm);
        }
    }
    pair_sync<TP>(pin);

    uint32_t mask0, mask1;
    if (!(xskip & 1)) {
        // ---- B: squared magnitudes of the run (bins 16 tp - 2 .. 16 tp + 17), peak masks, peak guard ---------
        int m0[20], m1[20];
        float2 esum = make_float2(0.f, 0.f);
        {
            const float2 *mrun = mg + 17 * tp;                                    // unit of bin 16 tp - 2
#pragma unroll
            for (int i = 0; i < 20; i++) {
                const float2 v = mrun[(i < 2) ? i : (i < 18) ? i + 1 : i + 2];
                m0[i] = __float_as_int(v.x);
                m1[i] = __float_as_int(v.y);
                if (i >= 2 && i < 18) esum = add2(esum, v);                       // own run: energy of the frame
            }
        }
        ring_masks<N, PCH>(p, m0, m1, esum, mask0, mask1, mine, pair, tp, pin, lane, FULL, has1, t);
        // thread 0 rebuilds the three stale bins whose operands sit in its registers in another order
        // (q = KS/2, KS, 3 KS/2); threads 1 .. 3 do it from the planes instead
        if (contract && tp >= 1 && tp <= 3) {
            const int q = (KS / 2) * tp;
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                const float *re = xp + ch * XPW, *im = xp + (2 + ch) * XPW;
                const float sr = (re[q] - re[N / 4 + q]) + (re[M - q] - re[N / 4 - q]);
                const float si = (im[q] - im[N / 4 + q]) - (im[M - q] - im[N / 4 - q]);
                const float2 w = PVB_TWH(twh + 2 * q);
                const float wr = 0.25f * w.x, wi = -0.25f * w.y;
                // (same operation order as cmul_s)
                xp[ch * XPW + M + q] = fmaf(sr, wr, -(si * wi));
                xp[(2 + ch) * XPW + M + q] = fmaf(sr, wi, si * wr);
            }
        }
    } else {
        mask0 = mask1 = 0;
    }
    pair_sync<TP>(pin);          // everyone has read the magnitudes: the descriptors take their place
    {
        // ---- C: descriptors -------------------------------------------------------------------------------
        {
            int4 *dz = reinterpret_cast<int4 *>(D0);
#pragma unroll
            for (int i = 0; i < (2 * G::DW / 4 + TP - 1) / TP; i++)
                if (tp + TP * i < 2 * G::DW / 4) dz[tp + TP * i] = make_int4(0, 0, 0, 0);
        }
        const int *krun = ktab + 20 * tp;                                         // keys of bins 16 tp .. + 15
        int pk0, nk0, pk1, nk1;
        bool any0, any1;
        {
            // keys of the nearest peaks below / above this thread's run (0: none below)
            const int il0 = (31 - __clz(mask0)) & 15, if0 = (__ffs(mask0) - 1) & 15;
            const int il1 = (31 - __clz(mask1)) & 15, if1 = (__ffs(mask1) - 1) & 15;
            const int ol0 = krun[il0], of0 = krun[if0], ol1 = krun[il1], of1 = krun[if1];
            const int none_above = ((2 * 8190) << 16) | 32768;                    // "peak" at +6142 with delta 0
            if constexpr (TP <= 32) {
                constexpr uint32_t PM = (TP == 32) ? 0xFFFFFFFFu : (TP == 16) ? 0xFFFFu : 0xFFu;
                const int hb = (TP == 32) ? 0 : int(threadIdx.x & (32 - TP));
                const uint32_t nz0 = (__ballot_sync(FULL, mask0 != 0) >> hb) & PM;
                const uint32_t nz1 = (__ballot_sync(FULL, mask1 != 0) >> hb) & PM;
                const uint32_t lt = (1u << tp) - 1u, gt = ~((2u << tp) - 1u) & PM;
                pk0 = __shfl_sync(FULL, ol0, hb + ((31 - __clz(nz0 & lt)) & (TP - 1)));
                nk0 = __shfl_sync(FULL, of0, hb + ((__ffs(nz0 & gt) - 1) & (TP - 1)));
                pk1 = __shfl_sync(FULL, ol1, hb + ((31 - __clz(nz1 & lt)) & (TP - 1)));
                nk1 = __shfl_sync(FULL, of1, hb + ((__ffs(nz1 & gt) - 1) & (TP - 1)));
                if (!(nz0 & lt)) pk0 = 0;
                if (!(nz0 & gt)) nk0 = none_above;
                if (!(nz1 & lt)) pk1 = 0;
                if (!(nz1 & gt)) nk1 = none_above;
                any0 = nz0 != 0;
                any1 = nz1 != 0;
            } else {
                constexpr int WPP = G::WPP;
                int *scr = reinterpret_cast<int *>(mine + G::BUF_SLOTS * 16);     // [4][TP] keys, [2][WPP] ballots
                scr[tp] = ol0; scr[TP + tp] = of0; scr[2 * TP + tp] = ol1; scr[3 * TP + tp] = of1;
                const uint32_t bl0 = __ballot_sync(FULL, mask0 != 0), bl1 = __ballot_sync(FULL, mask1 != 0);
                if (lane == 0) { scr[4 * TP + (tp >> 5)] = int(bl0); scr[4 * TP + WPP + (tp >> 5)] = int(bl1); }
                pair_sync<TP>(pin);
                const int *bal0 = scr + 4 * TP, *bal1 = bal0 + WPP;
                const int tb0 = ring_thread_below<WPP>(bal0, tp), ta0 = ring_thread_above<WPP>(bal0, tp);
                const int tb1 = ring_thread_below<WPP>(bal1, tp), ta1 = ring_thread_above<WPP>(bal1, tp);
                pk0 = (tb0 >= 0) ? scr[tb0] : 0;
                nk0 = (ta0 >= 0) ? scr[TP + ta0] : none_above;
                pk1 = (tb1 >= 0) ? scr[2 * TP + tb1] : 0;
                nk1 = (ta1 >= 0) ? scr[3 * TP + ta1] : none_above;
                any0 = ring_thread_last<WPP>(bal0) >= 0;
                any1 = ring_thread_last<WPP>(bal1) >= 0;
            }
        }
        pair_sync<TP>(pin);      // the descriptor arrays are zero
        ring_emit_descriptors<NB, TB, CB, G::LRW, G::INVALID_DELTA>(mask0, 16 * tp, pk0, nk0, krun, contract, D0);
        ring_emit_descriptors<NB, TB, CB, G::LRW, G::INVALID_DELTA>(mask1, 16 * tp, pk1, nk1, krun, contract, D1);
        // a channel without peaks: the shifted spectrum is zero (pv:121).  Rare (silence): its planes are
        // zeroed and every group gets a descriptor that reads them.
        if (!any0 || !any1) {
            pair_sync<TP>(pin);
#pragma unroll 1
            for (int ch = 0; ch < 2; ch++) {
                if (ch ? any1 : any0) continue;
                for (int i = tp; i < XPW; i += TP) xp[ch * XPW + i] = xp[(2 + ch) * XPW + i] = 0.f;
                int *Dc = ch ? D1 : D0;
                for (int x = RW * tp; x <= M; x += RW * TP) Dc[x] = Dc[x + 1] = 1;
            }
        }
    }
    pair_sync<TP>(pin);

    // ---- D: gather + Hermitian C2R pre-pass in registers (mirror of the split) -----------------------------
    if (contract) ring_gather_unsplit<N, true>(mine, twh, tp, lane, FULL, a, b);
    else ring_gather_unsplit<N, false>(mine, twh, tp, lane, FULL, a, b);
    } else {
    // ---- real split in registers -> XQ (2x scaled) --------------------------------------------------
    // slot j pairs (a[j], b[7-j]) at k = tp + KS j.  thread 0: j < 4: (b[j], b[7-j]) at k = KS/2 + KS j;
    // j >= 4: (a[j-4], a[(12-j)&7]) at k = KS (j-4); plus the self pair k = M/2 (a[4]).
    const int gA = tp + (tp >> 4);                                    // slot of bin tp
    const int gB = G::SM - tp - ((tp + 15) >> 4);                     // slot of bin M - tp
    const int sAlo = l0 ? KS / 2 + KS / 32 : gA, sAhi = l0 ? -4 * SS : gA;
    const int sBlo = l0 ? G::SM - KS / 2 - ((KS / 2 + 15) >> 4) : gB, sBhi = l0 ? G::SM + 4 * SS : gB;
    const int tlo = l0 ? KS / 2 : tp, thi = l0 ? -4 * KS : tp;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const cpx2 za = sel(l0, b[j], a[j]);
        ring_split(za, b[7 - j], PVB_TWH(twh + tlo + KS * j), XQ + sAlo + SS * j, XQ + sBlo - SS * j);
    }
#pragma unroll
    for (int j = 4; j < 8; j++) {
        const cpx2 za = sel(l0, a[j - 4], a[j]);
        const cpx2 zb = sel(l0, a[(12 - j) & 7], b[7 - j]);
        ring_split(za, zb, PVB_TWH(twh + thi + KS * j), XQ + sAhi + SS * j, XQ + sBhi - SS * j);
    }
    if (l0) ring_split(a[4], a[4], PVB_TWH(twh + M / 2), XQ + M / 2 + M / 32, XQ + M / 2 + M / 32);
    pair_sync<TP>(pin);

    // ---- peaks, regions of influence, shift (pv:95-173) -------------------------------------------------
    // X lives in float4 slots (both channels per bin); the shifted spectrum Y is written over it as
    // four planes of floats (re0 | re1 | im0 | im1, bin d at word d + (d >> YSHIFT)): the 32-bit scatter of
    // threads that own runs 16 bins apart then spreads over all banks.
    if (!(xskip & 1))
    {
        // contracting shifts (pitch factor < 1) read stale upper bins and let regions collide; per channel
        // with per-channel pitch factors, and the pair takes the extra steps if either channel needs them
        const bool contract0 = pfc0 < 1.0f, contract1 = pfc1 < 1.0f;
        const bool contract = contract0 | contract1;
        const float4 *runp = XQ + 17 * tp;                            // slot of bin 16 tp
        uint32_t mask0, mask1;
        {
            const float4 *hlo = tp ? runp - 3 : XQ;                   // bins 16 tp - 2, - 1 (thread 0: unused)
            int m0[20], m1[20];
            float2 esum = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 20; i++) {
                const float4 v = (i < 2) ? hlo[i] : (i < 18) ? runp[i - 2] : runp[i - 1];   // i >= 18: bins 16 tp + 16, + 17 (slot 16 is padding)
                const float2 re = make_float2(v.x, v.y), im = make_float2(v.z, v.w);
                const float2 mg = fma2(re, re, mul2(im, im));         // pv:82-92, float32
                m0[i] = __float_as_int(mg.x);
                m1[i] = __float_as_int(mg.y);
                if (i >= 2 && i < 18) esum = add2(esum, mg);          // own run: energy of the frame
            }
            ring_masks<N, PCH>(p, m0, m1, esum, mask0, mask1, mine, pair, tp, pin, lane, FULL, has1, t);
        }

        int dst0[16], dst1[16];
        int dl0 = 0, dl1 = 0;
        bool any0, any1;
        {
            const int *krun = ktab + 20 * tp;                         // keys of bins 16 tp .. + 15
            int rk[16], rk1[16];                                      // (rk1: PCH only)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int4 kv = *reinterpret_cast<const int4 *>(krun + 4 * i);
                rk[4 * i] = kv.x; rk[4 * i + 1] = kv.y; rk[4 * i + 2] = kv.z; rk[4 * i + 3] = kv.w;
            }
            // PCH: table word = delta field of channel 0 | channel 1; key = position of the bin | delta field
            const int posk = (2 * (16 * tp + 2048)) << 16;            // position part of the key of bin 16 tp
            if constexpr (PCH) {
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const int w = rk[e];
                    rk[e] = (w & 0xFFFF) + posk + (e << 17);
                    rk1[e] = int(unsigned(w) >> 16) + posk + (e << 17);
                }
            }
            // keys of the nearest peaks below / above this thread's run, and of the last peak
            const int il0 = (31 - __clz(mask0)) & 15, if0 = (__ffs(mask0) - 1) & 15;
            const int il1 = (31 - __clz(mask1)) & 15, if1 = (__ffs(mask1) - 1) & 15;
            int ol0 = krun[il0], of0 = krun[if0], ol1 = krun[il1], of1 = krun[if1];
            if constexpr (PCH) {
                ol0 = (ol0 & 0xFFFF) + posk + (il0 << 17);
                of0 = (of0 & 0xFFFF) + posk + (if0 << 17);
                ol1 = int(unsigned(ol1) >> 16) + posk + (il1 << 17);
                of1 = int(unsigned(of1) >> 16) + posk + (if1 << 17);
            }
            int pk0, nk0, lk0, pk1, nk1, lk1;
            const int none_above = (2 * 8190) << 16;                  // "peak" at +6142; below: key 0 = "peak" at -2048
            if constexpr (TP <= 32) {
                // one ballot bit per thread of the pair (frame 512: the pair's half of the warp)
                constexpr uint32_t PM = (TP == 32) ? 0xFFFFFFFFu : (TP == 16) ? 0xFFFFu : 0xFFu;
                const int hb = (TP == 32) ? 0 : int(threadIdx.x & (32 - TP));
                const uint32_t nz0 = (__ballot_sync(FULL, mask0 != 0) >> hb) & PM;
                const uint32_t nz1 = (__ballot_sync(FULL, mask1 != 0) >> hb) & PM;
                const uint32_t lt = (1u << tp) - 1u, gt = ~((2u << tp) - 1u) & PM;
                pk0 = __shfl_sync(FULL, ol0, hb + ((31 - __clz(nz0 & lt)) & (TP - 1)));
                nk0 = __shfl_sync(FULL, of0, hb + ((__ffs(nz0 & gt) - 1) & (TP - 1)));
                lk0 = __shfl_sync(FULL, ol0, hb + ((31 - __clz(nz0)) & (TP - 1)));
                pk1 = __shfl_sync(FULL, ol1, hb + ((31 - __clz(nz1 & lt)) & (TP - 1)));
                nk1 = __shfl_sync(FULL, of1, hb + ((__ffs(nz1 & gt) - 1) & (TP - 1)));
                lk1 = __shfl_sync(FULL, ol1, hb + ((31 - __clz(nz1)) & (TP - 1)));
                if (!(nz0 & lt)) pk0 = 0;
                if (!(nz0 & gt)) nk0 = none_above;
                if (!(nz1 & lt)) pk1 = 0;
                if (!(nz1 & gt)) nk1 = none_above;
                any0 = nz0 != 0;
                any1 = nz1 != 0;
            } else {
                // several warps per pair: exchange through the pair's scratch area
                constexpr int WPP = G::WPP;
                int *scr = reinterpret_cast<int *>(mine + G::BUF_SLOTS * 16);     // [4][TP] keys, [2][WPP] ballots
                scr[tp] = ol0; scr[TP + tp] = of0; scr[2 * TP + tp] = ol1; scr[3 * TP + tp] = of1;
                const uint32_t bl0 = __ballot_sync(FULL, mask0 != 0), bl1 = __ballot_sync(FULL, mask1 != 0);
                if (lane == 0) { scr[4 * TP + (tp >> 5)] = int(bl0); scr[4 * TP + WPP + (tp >> 5)] = int(bl1); }
                pair_sync<TP>(pin);
                const int *bal0 = scr + 4 * TP, *bal1 = bal0 + WPP;
                const int tb0 = ring_thread_below<WPP>(bal0, tp), ta0 = ring_thread_above<WPP>(bal0, tp);
                const int tl0 = ring_thread_last<WPP>(bal0);
                const int tb1 = ring_thread_below<WPP>(bal1, tp), ta1 = ring_thread_above<WPP>(bal1, tp);
                const int tl1 = ring_thread_last<WPP>(bal1);
                pk0 = (tb0 >= 0) ? scr[tb0] : 0;
                nk0 = (ta0 >= 0) ? scr[TP + ta0] : none_above;
                lk0 = scr[tl0 & (TP - 1)];
                pk1 = (tb1 >= 0) ? scr[2 * TP + tb1] : 0;
                nk1 = (ta1 >= 0) ? scr[3 * TP + ta1] : none_above;
                lk1 = scr[2 * TP + (tl1 & (TP - 1))];
                any0 = tl0 >= 0;
                any1 = tl1 >= 0;
            }
            dl0 = (lk0 & 0xFFFF) - 32768;
            dl1 = (lk1 & 0xFFFF) - 32768;
            int col0 = 0, col1 = 0;                                   // DEEP: (peaks below this thread's run - 1) mod 3
            if constexpr (COLOURS) {
                // peaks below this thread's run, both channels in one word (at most N/6 peaks per channel)
                constexpr int W = (TP < 32) ? TP : 32;
                const int cnt = __popc(mask0) | (__popc(mask1) << 16);
                int inc = cnt;
#pragma unroll
                for (int d = 1; d < W; d <<= 1) {
                    const int v = __shfl_up_sync(FULL, inc, d, W);
                    if ((lane & (W - 1)) >= d) inc += v;
                }
                int below = inc - cnt;
                if constexpr (G::WPP > 1) {
                    int *wtot = reinterpret_cast<int *>(mine + G::BUF_SLOTS * 16) + 4 * TP + 8 + 20;
                    if (lane == 31) wtot[tp >> 5] = inc;
                    pair_sync<TP>(pin);
#pragma unroll
                    for (int w = 0; w < G::WPP - 1; w++)
                        if (w < (tp >> 5)) below += wtot[w];
                }
                col0 = ((below & 0xFFFF) + 2) % 3;
                col1 = ((below >> 16) + 2) % 3;
            }
            // both scans run unconditionally (a channel without peaks ends up with every bin on the
            // dump slot): two independent instruction streams the scheduler can interleave
            ring_owner_scan<G::XQ_SLOTS - 1, G::YS, COLOURS>(mask0, 16 * tp, pk0, nk0, rk, (contract0 && !DEEP) ? int(0x80000000u) : 0, dst0, col0);
            if constexpr (PCH)
                ring_owner_scan<G::XQ_SLOTS - 1, G::YS, COLOURS>(mask1, 16 * tp, pk1, nk1, rk1, (contract1 && !DEEP) ? int(0x80000000u) : 0, dst1, col1);
            else
                ring_owner_scan<G::XQ_SLOTS - 1, G::YS, COLOURS>(mask1, 16 * tp, pk1, nk1, rk, (contract1 && !DEEP) ? int(0x80000000u) : 0, dst1, col1);
        }

        // sources into registers: own run, bin M and the first stale level (what _realTransform4
        // leaves in slots N/2 + q, bundle:394-438, rebuilt from the valid half); bins M + tp + TP i
        float4 xv[16];
#pragma unroll
        for (int e = 0; e < 16; e++) xv[e] = runp[e];
        float4 ext[4];
        ext[0] = ext[1] = ext[2] = ext[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l0) ext[0] = XQ[G::SM];
        float4 ext4 = make_float4(0.f, 0.f, 0.f, 0.f);               // DEEP: slot N/2 + N/8 (thread 0)
        (void)ext4;
        if (contract && !(xskip & 32)) {
            constexpr int QO = N / 4 + N / 64;                        // slots between bins k and k + N/4
            auto level1 = [&](int qq) {
                const int sq = qq + (qq >> 4);                        // slot of bin q
                const int sm = G::SM - qq - ((qq + 15) >> 4);         // slot of bin M - q
                const cpx2 A = unpack4(XQ[sq]), Bv = unpack4(XQ[sq + QO]);         // bins q, N/4 + q
                const cpx2 Cv = unpack4(XQ[sm]), D = unpack4(XQ[sm - QO]);         // bins M - q, N/4 - q
                const float2 sr = add2(sub2(A.re, Bv.re), sub2(Cv.re, D.re));
                const float2 si = sub2(sub2(A.im, Bv.im), sub2(Cv.im, D.im));
                const float2 w = PVB_TWH(twh + 2 * qq);
                return pack4(cmul_s(cpx2{sr, si}, 0.25f * w.x, -0.25f * w.y));
            };
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int q = tp + TP * i;
                const float4 sv = level1(q ? q : 1);
                if (q) ext[i] = sv;
            }
            if constexpr (DEEP) { if (l0) ext4 = level1(N / 8); }
        }
        // DEEP: the second stale level while X is still there.  Slot N/2 + q, q = N/8 + tp + TP i: for even i
        // (every thread) and for odd i (thread 0) the block-tree walk stops at L = N/16, r = 16, s = 2 + 4 sb,
        // sb = 2 + i / 2, o = q - sb N/16 <= N/32, and DFT_L(xw[16 m + s])[o] = 1/16 sum_u W_N^{-s (o + u L)} X[o + u L]
        // (X extended Hermitian above N/2): 16 terms from the spectrum instead of N/16 frame samples.  In ring
        // order nothing changes: (o + u L) - (N/2 + q) is a multiple of L and t of 16.
        float4 ext2[4];
        (void)ext2;
        if constexpr (DEEP) {
#pragma unroll
            for (int ip = 0; ip < 4; ip++) {
                ext2[ip] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int q = N / 8 + tp + TP * ip;
                const bool lvl2 = ((ip & 1) == 0 || l0) && q != N / 8;
                const bool need = (any0 && M + q + dl0 < NB) || (any1 && M + q + dl1 < NB);
                if (lvl2 && need) {
                    constexpr int L16 = N / 16;
                    const int sb = 2 + (ip >> 1), o = q - sb * L16, sx = 2 + 4 * sb;
                    cpx2 acc;
                    acc.re = acc.im = make_float2(0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const int idx = o + u * L16;
                        const bool mir = (u > 8) || (u == 8 && o != 0);                  // bins above N/2: conj X[N - idx]
                        const int b = mir ? N - idx : idx;
                        const cpx2 xv = unpack4(XQ[b + (b >> 4)]);
                        const float2 xi = mir ? make_float2(-xv.im.x, -xv.im.y) : xv.im;
                        const int k = (sx * idx) & (N - 1);
                        float2 w = PVB_TWH(twh + (k & (M - 1)));
                        if (k & M) w = make_float2(-w.x, -w.y);
                        // + (xr + j xi) conj(w)
                        acc.re = fma2(xv.re, bc2(w.x), acc.re);
                        acc.re = fma2(xi, bc2(w.y), acc.re);
                        acc.im = fma2(xi, bc2(w.x), acc.im);
                        acc.im = fma2(xv.re, bc2(-w.y), acc.im);
                    }
                    ext2[ip] = make_float4(0.0625f * acc.re.x, 0.0625f * acc.re.y, 0.0625f * acc.im.x, 0.0625f * acc.im.y);
                }
            }
        }
        // DEEP, pitch factors below 0.5 (down to 0.33: the reference's demo divides a pitch in [0.5, 1.5] by a speed
        // in [0.5, 1.5], main.js:82,93): the last region reaches slots N/2 + q up to q = N/3, i.e. into the quarter
        // 3N/4 + o of the output, which holds DFT_{N/4}(xw[4 m + 3])[o] for o <= N/8 (bundle:394-438):
        //   1/4 W_N^{-3 o} (X[o] - j X[N/4 + o] - conj X[N/2 - o] + j conj X[N/4 - o]);  q = N/4 + tp + TP i, i < 3
        constexpr bool wide = WIDE;
        float4 ext3[3];
        (void)ext3;
        if constexpr (WIDE) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                ext3[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int o = tp + TP * i, q = N / 4 + o;
                const bool need = (any0 && M + q + dl0 < NB) || (any1 && M + q + dl1 < NB);
                if (wide && need) {
                    constexpr int QO = N / 4 + N / 64;                // slots between bins k and k + N/4
                    const int so = o + (o >> 4);                      // slot of bin o
                    const int sm = G::SM - o - ((o + 15) >> 4);       // slot of bin M - o
                    const cpx2 A = unpack4(XQ[so]), Bv = unpack4(XQ[so + QO]);         // bins o, N/4 + o
                    const cpx2 Cv = unpack4(XQ[sm]), D = unpack4(XQ[sm - QO]);         // bins M - o, N/4 - o
                    // (A - conj C) + j (conj D - B)
                    const float2 sr = add2(sub2(A.re, Cv.re), add2(Bv.im, D.im));
                    const float2 si = add2(add2(A.im, Cv.im), sub2(D.re, Bv.re));
                    const float2 w = PVB_TWH(twh + 3 * o);
                    ext3[i] = pack4(cmul_s(cpx2{sr, si}, 0.25f * w.x, -0.25f * w.y));
                }
            }
        }
        pair_sync<TP>(pin);      // every thread holds its sources: the buffer becomes Y
        // (PVB_RING_EXACT: while contracting every bin of [0, nb) is stored exactly once in the first
        // sub-step, provided both channels have peaks)
        if ((DEEP || !(PVB_RING_EXACT && contract && any0 && any1)) && !(xskip & 8)) {
#pragma unroll
            for (int i = 0; i < 17; i++) XQ[tp + TP * i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tp < 2) XQ[17 * TP + tp] = make_float4(0.f, 0.f, 0.f, 0.f);
            pair_sync<TP>(pin);
        }

        using Y = RingY<G::XQ_SLOTS>;
#if PVB_RING_NO_DUMP
#define PVB_NOT_DUMP(d) (((d) & 0x7fffffff) != Y::SB * (G::XQ_SLOTS - 1))
#else
#define PVB_NOT_DUMP(d) true
#endif
        if constexpr (DEEP) {
            // up to three regions (i, i + 1, i + 2) land on one bin: one sub-step per colour = ordinal of the owning
            // peak mod 3 (low bits of dst), each with pairwise disjoint destinations (loads of a batch issued before
            // its first store).
            constexpr int DUMPB = Y::SB * (G::XQ_SLOTS - 1);
#pragma unroll
            for (int e = 0; e < 16; e++) {
                if (dst0[e] == DUMPB + (dst0[e] & 3)) dst0[e] = -1;   // outside [0, nb): no colour matches
                if (dst1[e] == DUMPB + (dst1[e] & 3)) dst1[e] = -1;
            }
            // the last region's sources from bin N/2 up to the end of the first and all of the second stale level go first (plain stores
            // onto the zero-filled planes: nothing else has been written yet), which ends their live ranges
            {
                auto put_ext = [&](int q, const float4 &v) {
                    const int d0 = M + q + dl0, d1 = M + q + dl1;
                    if (any0 && unsigned(d0) < unsigned(NB)) Y::store(mine + Y::SB * (d0 + (d0 >> G::YS)), 0, v.x, v.z);
                    if (any1 && unsigned(d1) < unsigned(NB)) Y::store(mine + Y::SB * (d1 + (d1 >> G::YS)), 1, v.y, v.w);
                };
#pragma unroll
                for (int i = 0; i < 4; i++) put_ext(tp + TP * i, ext[i]);
                if (l0) put_ext(N / 8, ext4);
#pragma unroll
                for (int ip = 0; ip < 4; ip++)                       // second level (zeros where not needed / not level 2)
                    if (((ip & 1) == 0 || l0) && !(ip == 0 && l0)) put_ext(N / 8 + tp + TP * ip, ext2[ip]);
                if constexpr (WIDE) {
#pragma unroll
                    for (int i = 0; i < 3; i++) put_ext(N / 4 + tp + TP * i, ext3[i]);
                }
            }
            if constexpr (WIDE) {
                // below 0.5 any number of regions can land on one bin (a wide region's right half reaches over
                // several narrow neighbours): shared-memory atomic adds on top of the stores above
                pair_sync<TP>(pin);
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    if (dst0[e] >= 0) Y::atomic_add(mine + (dst0[e] & ~3), 0, xv[e].x, xv[e].z);
                    if (dst1[e] >= 0) Y::atomic_add(mine + (dst1[e] & ~3), 1, xv[e].y, xv[e].w);
                }
            } else
#pragma unroll 1
            for (int col = 0; col < 3; col++) {
                pair_sync<TP>(pin);
                unsigned char *minec = mine - col;                    // cancels the colour bits of dst
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float2 o0[4], o1[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int e = 4 * g + i;
                        o0[i] = o1[i] = make_float2(0.f, 0.f);
                        if ((dst0[e] & 3) == col) o0[i] = Y::load(minec + dst0[e], 0);
                        if ((dst1[e] & 3) == col) o1[i] = Y::load(minec + dst1[e], 1);
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int e = 4 * g + i;
                        if ((dst0[e] & 3) == col) Y::store(minec + dst0[e], 0, o0[i].x + xv[e].x, o0[i].y + xv[e].z);
                        if ((dst1[e] & 3) == col) Y::store(minec + dst1[e], 1, o1[i].x + xv[e].y, o1[i].y + xv[e].w);
                    }
                }
            }
            // the last region's sources beyond the first stale level (distinct destinations, everything else is complete)
            pair_sync<TP>(pin);
            auto add_ext = [&](int q, const float4 &v) {
                const int d0 = M + q + dl0, d1 = M + q + dl1;
                if (any0 && unsigned(d0) < unsigned(NB)) {
                    unsigned char *at = mine + Y::SB * (d0 + (d0 >> G::YS));
                    const float2 o = Y::load(at, 0);
                    Y::store(at, 0, o.x + v.x, o.y + v.z);
                }
                if (any1 && unsigned(d1) < unsigned(NB)) {
                    unsigned char *at = mine + Y::SB * (d1 + (d1 >> G::YS));
                    const float2 o = Y::load(at, 1);
                    Y::store(at, 1, o.x + v.y, o.y + v.w);
                }
            };
            // slots N/2 + q, N/8 < q < N/4
            constexpr int LOG2N = (N == 256) ? 8 : (N == 512) ? 9 : (N == 1024) ? 10 : (N == 2048) ? 11 : 12;
            constexpr int L0 = (LOG2N & 1) ? 2 : 4;                  // smallest block of the radix-4 recursion
#pragma unroll 1
            for (int i = 4; i < 8; i++) {
                const int q = tp + TP * i;
                const bool need = (any0 && M + q + dl0 < NB) || (any1 && M + q + dl1 < NB);
                if (!need) continue;
                if (q == N / 8) continue;                            // last slot of the first level: stored above
                if ((i & 1) == 0 || l0) continue;                    // second level: from the spectrum, stored above
                int lg = LOG2N, r = 1, sq_ = 0, o = M + q;
                while ((1 << lg) > L0 && o > (1 << (lg - 1))) {
                    const int sb = o >> (lg - 2);
                    o -= sb << (lg - 2);
                    sq_ += r * sb;
                    r <<= 2;
                    lg -= 2;
                }
                const float2 *sp_ = dsc + 2 * (sq_ >> 4) + ((sq_ >> 2) & 1);      // xw[r m + s], m at stride r / 8
                // W_L^{o m} = W_N^{16 j}, j = (o r / 16) m mod N/16
                const float2 *t16 = dsc + N / 8;
                const int stride = r >> 3, kstep = (o * (r >> 4)) & (N / 16 - 1), L = 1 << lg;
                cpx2 acc;
                acc.re = acc.im = make_float2(0.f, 0.f);
                int k = 0;
#pragma unroll 4
                for (int m = 0; m < L; m++) {
                    float2 xs = sp_[m * stride];
                    const float2 w = t16[G::T16_HALF ? (k & (TP - 1)) : k];
                    if (G::T16_HALF && (k & TP)) xs = make_float2(-xs.x, -xs.y);
                    acc.re = fma2(xs, bc2(w.x), acc.re);
                    acc.im = fma2(xs, bc2(w.y), acc.im);
                    k = (k + kstep) & (N / 16 - 1);
                }
                // frame order -> ring order: U[b] = X[b] W_N^{b t}, b = N/2 + q (t is a multiple of 64); the
                // spectrum in shared memory is 2 U (the real split leaves the factor to the synthesis window)
                const int kr = (q * t) & (N - 1);
                float2 wr = PVB_TWH(twh + (kr & (M - 1)));
                if (kr & M) wr = make_float2(-wr.x, -wr.y);
                add_ext(q, pack4(cmul_s(acc, 2.0f * wr.x, 2.0f * wr.y)));
            }
        } else {
        // first sub-step: plain stores (pairwise disjoint destinations)
        if (!(xskip & 16))
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (dst0[e] >= 0 && PVB_NOT_DUMP(dst0[e])) Y::store(mine + dst0[e], 0, xv[e].x, xv[e].z);
            if (dst1[e] >= 0 && PVB_NOT_DUMP(dst1[e])) Y::store(mine + dst1[e], 1, xv[e].y, xv[e].w);
        }
        if (any0) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int d = M + tp + TP * i + dl0;
                if (unsigned(d) < unsigned(NB)) Y::store(mine + Y::SB * (d + (d >> G::YS)), 0, ext[i].x, ext[i].z);
            }
        }
        if (any1) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int d = M + tp + TP * i + dl1;
                if (unsigned(d) < unsigned(NB)) Y::store(mine + Y::SB * (d + (d >> G::YS)), 1, ext[i].y, ext[i].w);
            }
        }
        if (contract && !(xskip & 4)) {
            // second sub-step: left halves add on top (pairwise disjoint among themselves, so the
            // loads of a batch can all be issued before the first store)
            pair_sync<TP>(pin);
            unsigned char *mine2 = mine + 0x80000000u;                // cancels the flag bit of dst
#pragma unroll
            for (int g = 0; g < 2; g++) {
                float2 o0[8], o1[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int e = 8 * g + i;
                    o0[i] = o1[i] = make_float2(0.f, 0.f);
                    if (dst0[e] < 0 && PVB_NOT_DUMP(dst0[e])) o0[i] = Y::load(mine2 + dst0[e], 0);
                    if (dst1[e] < 0 && PVB_NOT_DUMP(dst1[e])) o1[i] = Y::load(mine2 + dst1[e], 1);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int e = 8 * g + i;
                    if (dst0[e] < 0 && PVB_NOT_DUMP(dst0[e])) Y::store(mine2 + dst0[e], 0, o0[i].x + xv[e].x, o0[i].y + xv[e].z);
                    if (dst1[e] < 0 && PVB_NOT_DUMP(dst1[e])) Y::store(mine2 + dst1[e], 1, o1[i].x + xv[e].y, o1[i].y + xv[e].w);
                }
            }
        }
        }   // !DEEP
    }
    pair_sync<TP>(pin);

    // ---- Hermitian C2R pre-pass in registers (mirror of the split) -------------------------------------
    {
        // words of the same bins in the Y planes (pad every 2^YS bins instead of every 16)
        constexpr int YS = G::YS, YP = 1 << YS;
        constexpr int YSS = KS + (KS >> YS), YSM = M + (M >> YS);
        static_assert(KS % YP == 0 && YS >= 4, "Y-plane padding must divide the butterfly stride and fit the buffer");
        const int yA = tp + (tp >> YS), yB = YSM - tp - ((tp + YP - 1) >> YS);
        const int yAlo = l0 ? KS / 2 + ((KS / 2) >> YS) : yA, yAhi = l0 ? -4 * YSS : yA;
        const int yBlo = l0 ? YSM - KS / 2 - ((KS / 2 + YP - 1) >> YS) : yB, yBhi = l0 ? YSM + 4 * YSS : yB;
        cpx2 zk[8], zmk[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int sa = (j < 4 ? yAlo : yAhi) + YSS * j, sb = (j < 4 ? yBlo : yBhi) - YSS * j;
            cpx2 yk = RingY<G::XQ_SLOTS>::load_bin(mine, sa), ym = RingY<G::XQ_SLOTS>::load_bin(mine, sb);
            if (j == 4) {                        // thread 0: k == 0, bins 0 and N/2 enter with their real part only
                yk.im = make_float2(l0 ? 0.f : yk.im.x, l0 ? 0.f : yk.im.y);
                ym.im = make_float2(l0 ? 0.f : ym.im.x, l0 ? 0.f : ym.im.y);
            }
            const float2 w = PVB_TWH(twh + (j < 4 ? tlo : thi) + KS * j);
            ring_unsplit(yk, ym, w, zk[j], zmk[j]);
        }
        cpx2 zh, dummy;
        {
            const cpx2 y = RingY<G::XQ_SLOTS>::load_bin(mine, M / 2 + ((M / 2) >> YS));
            ring_unsplit(y, y, PVB_TWH(twh + M / 2), zh, dummy);
        }
        a[0] = sel(l0, zk[4], zk[0]);
        a[1] = sel(l0, zk[5], zk[1]);
        a[2] = sel(l0, zk[6], zk[2]);
        a[3] = sel(l0, zk[7], zk[3]);
        a[4] = sel(l0, zh, zk[4]);
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
    }   // scatter middle
    pair_sync<TP>(pin);  // everyone has read Y: the exchange slots may overwrite it

    // ---- inverse pass 1 (DIT): butterflies A and B over k3, twiddle conj(W_64^{k2 m3}) -------------------
    dft8<true>(a);
    dft8<true>(b);
    {
        const int k2a = tp >> G::LR1, k2b = kB >> G::LR1;
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
    pair_sync<TP>(pin);

    // accumulator values (L2 hits thanks to the prefetch) are requested before the last exchange so
    // that their latency hides behind inverse pass 2; the tail slot starts from zero (ola:134)
    float4 *al = p.acc2 + size_t(pair) * (N / 2) + (HB ? 0 : cn);
    float4 q[16];
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int h = PVB_FH(e), fr = PVB_FR(e), f = PVB_FB(fr);
        q[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (PVB_ROLE(h, fr) < NEWFROM) q[e] = al[PVB_RING_IDX(h, f)];
    }

    // ---- inverse pass 2: butterflies (k1, m3) over k2, twiddle conj(W_M^{k1 (m3 + 8 m2 + 64 toff)}) -------
    if constexpr (NSPL == 2) {
        // mirror of forward pass 2: with Y[k1] the twiddled outputs of rows k' and k' + 16,
        // row (0, k') = (Y[k'] + Y[k'+16]) and row (1, k') = (Y[k'] - Y[k'+16]) conj(W_32^{k'}); the
        // conjugates of the two forward tables carry W_M^{k' (n + 64 toff)} (and W_32^{k'})
        const int kp = tp >> 3;
        float4 *b0 = ex + G::RS * kp + m3l, *b1 = b0 + G::RS * 16;
        const float2 *t0 = tw1 + G::TW1_ROW * kp + m3l, *t1 = t0 + G::TW1_ROW * 16;
        const float sgn = (toff & 1) ? -1.f : 1.f;
        cpx2 x0[8], x1[8];
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            x0[k2] = unpack4(b0[G::G8 * k2]);
            x1[k2] = unpack4(b1[G::G8 * k2]);
        }
        dft8<true>(x0);
        dft8<true>(x1);
#pragma unroll
        for (int m2 = 0; m2 < 8; m2++) {
            const float2 w = w128[m3l + 8 * m2];
            const cpx2 v1 = cmul_s(x1[m2], sgn * w.x, -sgn * w.y);    // conj(W_128^n) (-1)^toff
            const float2 wa = t0[8 * m2], wb = t1[8 * m2];
            b0[G::G8 * m2] = pack4(cmul_s(cadd(x0[m2], v1), wa.x, -wa.y));
            b1[G::G8 * m2] = pack4(cmul_s(csub(x0[m2], v1), wb.x, -wb.y));
        }
    } else if (!(xskip & 2))
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k1 = (tp >> 3) + (R1 / 2) * h;
        float4 *bp = ex + G::RS * k1 + m3l;
        const float2 *twp = tw1 + G::TW1_ROW * k1 + m3l;
        cpx2 x[8];
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) x[k2] = unpack4(bp[G::G8 * k2]);
        dft8<true>(x);
#pragma unroll
        for (int m2 = 0; m2 < 8; m2++) {
            const float2 w = twp[8 * m2];
            bp[G::G8 * m2] = pack4(cmul_s(x[m2], w.x, -w.y));
        }
    }
    pair_sync<TP>(pin);

    // ---- inverse pass 3: butterflies n over k1 -> ring samples; window, overlap-add, emit ------------------
    {
        float *o0 = p.out + (size_t(hopi) * p.num_channels + c0) * hop + 2 * cn;
        const float *wol = swout + 2 * cn;
        const int row0 = RT * sp;
#pragma unroll
        for (int h = 0; h < G::NB1; h++) {
            const int nl = PVB_COL(h);
            cpx2 x[RT];
#pragma unroll
            for (int k1 = 0; k1 < RT; k1++) x[k1] = unpack4(ex[G::RS * (row0 + k1) + nl + (G::G8 - 8) * (nl >> 3)]);
            dft_r<RT, true>(x);
#pragma unroll
            for (int j = 0; j < RT; j++) {
                // window_out = hannWindow / (2 N R): fromComplexArray, applyHannWindow and the division
                // by nbOverlaps (pv:65-67, ola:153) in one multiply (the scales are powers of two)
                const int fb = PVB_FB(j);
                float2 wo = PVB_TLD2(reinterpret_cast<const float2 *>(wol + 2 * TPH * h + 128 * fb));
                if constexpr (G::WOUT_DERIVED) wo = make_float2(wo.x * WOUT_SCALE, wo.y * WOUT_SCALE);
                const float4 qv = q[RT * h + j];
                const float2 y0 = fma2(x[j].re, bc2(wo.x), make_float2(qv.x, qv.y));
                const float2 y1 = fma2(x[j].im, bc2(wo.y), make_float2(qv.z, qv.w));
                if (PVB_ROLE(h, j) < HEADTO) {                        // head: emit (ola:111-118)
                    *reinterpret_cast<float2 *>(o0 + 2 * TPH * h + 128 * fb) = make_float2(y0.x, y1.x);
                    if (has1) *reinterpret_cast<float2 *>(o0 + hop + 2 * TPH * h + 128 * fb) = make_float2(y0.y, y1.y);
                } else {
                    al[PVB_RING_IDX(h, fb)] = make_float4(y0.x, y0.y, y1.x, y1.y);
                }
            }
        }
    }
    return true;
#undef PVB_NOT_DUMP
#undef PVB_RING_OFF
#undef PVB_FR
#undef PVB_FH
#undef PVB_FB
#undef PVB_ROLE
#undef PVB_RING_IDX
#undef PVB_COL
#undef PVB_TLD2
#undef PVB_TWH
}

// N = frame size (512: half a warp per pair, 1024: one warp, 2048: two warps, 4096: four warps),
// NBLK = hop / 128 (see ring_one_call).
//
// MULTI: the launch does p.num_hops consecutive process() calls (pvb_process_many): every pair loops over
// the calls, its state goes back and forth through L1 / L2 instead of HBM (one DRAM round trip of the
// rings per launch instead of per call), completion flags are taken and released once, and only the
// first-pass twiddle table is re-staged per call.  Bit-identical to num_hops single launches.
// DEEP: 1 = pitch factors down to 0.5, 2 = down to 0.33 (see ring_one_call); scalar (then below 0.75) or per
// channel (PCH: every channel in [0.5, 64] / [0.33, 64]), one call per launch.
template <int N, int NBLK, bool PCH = false, bool MULTI = false, int DEEP = 0>
__global__ void __launch_bounds__((DEEP ? RingGeoT<N, PCH>::DEEP_PAIRS : MULTI ? RingGeoT<N, PCH>::MULTI_PAIRS : RingGeoT<N, PCH>::MAX_PAIRS) * RingGeoT<N, PCH>::TP)
__maxnreg__((DEEP ? RingGeoT<N, PCH>::DEEP_REGS : MULTI ? RingGeoT<N, PCH>::BIG_REGS : 128))     // two CTAs per SM either way
pv_process_ring_kernel(const RingParams p) {
    constexpr int TP = RingGeoT<N, PCH>::TP;
    const int tp = threadIdx.x % TP, pin = threadIdx.x / TP;
    const int pair = blockIdx.x * (blockDim.x / TP) + pin;
    bool live = 2 * pair < p.num_channels;
    using G_ = RingGeoT<N, PCH>;
    if constexpr (G_::TMEMX) {
        // 128 columns of tensor memory per CTA (two CTAs per SM: 256 of 512): one warp allocates, the base
        // address reaches the others through shared memory behind the CTA barrier that follows the frame loads
        extern __shared__ __align__(16) unsigned char smem_k[];
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;"
                         ::"r"(unsigned(__cvta_generic_to_shared(smem_k + G_::OFF_MBAR + 8))) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
#ifdef PVB_EXPERIMENTS
    if (p.stamps && threadIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMin(p.stamps, now);
    }
#endif
    if constexpr (MULTI) {
        live = ring_one_call<N, NBLK, PCH, 1>(p, 0, live);
#pragma unroll 1
        for (int hopi = 1; hopi < p.num_hops; hopi++) live = ring_one_call<N, NBLK, PCH, 2>(p, hopi, live);
    } else {
        live = ring_one_call<N, NBLK, PCH, 0, DEEP>(p, 0, live);
    }
#ifdef PVB_EXPERIMENTS
    if (p.stamps && threadIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMax(p.stamps + 1, now);
    }
#endif
    if (live) {
        // release: state and output of this pair are complete for call my_seq.  The pair barrier orders
        // every thread's stores before thread 0's release store, which is cumulative at gpu scope;
        // PVB_RING_LANE_FENCE=1 additionally fences in every thread (the first version of this code).
#if PVB_RING_LANE_FENCE
        __threadfence();
#endif
        pair_sync<TP>(pin);
        if (tp == 0)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.done + pair), "r"(p.my_seq) : "memory");
        // flag mode skipped the wait at the top: take it here, where the previous grid is long gone, so
        // that completion stays transitive along the stream
        if (p.flag_mode) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if constexpr (G_::TMEMX) {
        // every warp of the CTA is done with its window: the allocating warp frees the columns
        extern __shared__ __align__(16) unsigned char smem_k[];
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tb = *reinterpret_cast<const volatile uint32_t *>(smem_k + G_::OFF_MBAR + 8);
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tb) : "memory");
        }
    }
}

// tables the ring-order kernel copies into shared memory: NJ first-pass twiddle tables
// tw1[toff][k1][n] = W_M^{(n + 64 toff) k1} (rows of TW1_ROW), then w64[a][b] = W_64^{ab}, twh[k] = W_N^k
template <int N>
constexpr int ring_host_table_bytes() { return RingGeoT<N>::NTAB * RingGeoT<N>::TW1_BYTES + RingGeoT<N>::W64_BYTES + RingGeoT<N>::TWH_BYTES + RingGeoT<N>::W128_BYTES; }

template <int N>
inline void ring_host_tables(const float2 *tw /* [N] W_N^j */, float2 *out /* ring_host_table_bytes / 8 */) {
    using G = RingGeoT<N>;
    float2 *w64 = out + G::NTAB * (G::TW1_BYTES / 8), *w128 = w64 + G::W64_BYTES / 8, *twh = w128 + G::W128_BYTES / 8;
    for (int i = 0; i < ring_host_table_bytes<N>() / 8; i++) out[i] = make_float2(0.f, 0.f);
    for (int u = 0; u < G::NTAB; u++)
        for (int row = 0; row < G::R1; row++)
            for (int n = 0; n < 64; n++) {
                // frame 4096: row = 16 s + k' holds W_M^{(n + 64 toff) k'} W_32^{s k'}
                // frame 256: table u = 2 toff + half; ring columns n < 32 of a half-rotated ring belong
                // to the next ring block: W_M^{n k1} W_R1^{(toff + carry) k1}
                const int k1 = row % G::RT, sp = row / G::RT;
                const int toff = G::HB ? (u >> 1) : u;
                const int carry = (G::HB && (u & 1) && n < 32) ? 1 : 0;
                out[u * (G::TW1_BYTES / 8) + G::TW1_ROW * row + n] =
                    tw[(2 * (n + 64 * (toff + carry)) * k1 + (N / 32) * sp * k1) & (N - 1)];
            }
    if (G::NSPL == 2)
        for (int n = 0; n < 64; n++) w128[n] = tw[((N / 128) * n) & (N - 1)];
    for (int a = 0; a < 8; a++)
        for (int b = 0; b < 8; b++) w64[8 * a + b] = tw[((N / 64) * a * b) & (N - 1)];
    for (int k = 0; k <= G::M; k++) twh[k] = tw[k];
}

// planar [C'][N] rings (pv_kernel.cuh conventions: frame sample n at hist[(n + rb + hop) mod N],
// accumulator sample k at acc[(k + rb) mod N]) <-> paired rings aligned to t
static __global__ void pv_ring_convert_kernel(float *hist_planar, float *acc_planar, float2 *hist2, float2 *acc2,
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
