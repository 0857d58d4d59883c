// pv_api.cu — host side of the C ABI declared in include/phaze_b200.h.
//
// Owns the per-handle device state (history ring, overlap-add ring, tables), the
// stream, and the launch of the fused kernel in pv_kernel.cuh.  No compute happens on
// the host: if there is no CUDA device every entry point reports PVB_ERR_CUDA.
#include "../../include/phaze_b200.h"
#include "pv_kernel.cuh"
#include "pv_kernel_warp.cuh"
#include "pv_kernel_cta.cuh"
#include "pv_kernel_ring.cuh"
#include "pv_ring_launch.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <map>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

struct pvb_processor {
    int n = 0, hop = 0, overlaps = 0, channels = 0, device = 0;
    uint64_t ring_calls = 0;     // process() calls since the rings were last zeroed / set
    uint64_t cursor_calls = 0;   // timeCursor / hop (pv:31,71)
    // layout of d_hist / d_acc: planar rows (pv_kernel.cuh), pairs of channels interleaved and
    // aligned to the time cursor (pv_kernel_ring.cuh), or all zero (either reading is valid)
    enum Layout { ZERO, PLANAR, PAIRED };
    Layout layout = ZERO;
    float *d_hist = nullptr, *d_acc = nullptr, *d_window = nullptr, *d_window_out = nullptr;
    int num_sms = 148;
    float2 *d_tw = nullptr;
    float4 *d_ring_tab = nullptr;    // tables of the ring-order kernel (frames 1024 and 2048)
    unsigned *d_done = nullptr;      // [pairs]: per-pair completion flags of the ring-order kernel
    unsigned ring_seq = 0;           // sequence number of this handle's last ring-order launch
    // sticky device-error word: pinned, mapped host memory the kernels write only when a completion flag
    // is still missing after the whole previous grid has drained (h_err[0]: pairs lost so far)
    unsigned *h_err = nullptr, *d_err = nullptr;
    // peak guard (pv_kernel_ring.cuh): shared per (device, frame size), see acquire_exact_pool
    struct ExactPool *xpool = nullptr;
    unsigned long long *d_xcount = nullptr;   // channel frames re-decided in float64 so far
    // per-channel pitch factors (pvb_process_pf): last array uploaded, its device copy, its range
    std::vector<float> h_pf;
    float *d_pf = nullptr;
    bool pf_fast = false;            // every channel inside the ring-order kernel's range [0.75, 64]
    bool pf_deep = false;            // every channel inside [0.33, 64]: the ring-order kernel's DEEP instances
    bool pf_half = false;            // every channel inside [0.5, 64]: the DEEP instances with ordered sub-steps
    // pvb_set_option
    int opt_kernel = 0, opt_launch_mode = 0, opt_inputs_ready = 0, opt_peak_guard = 0, opt_many_mode = 0;
    cudaStream_t last_stream = nullptr;   // stream of the most recent submission (state entry points wait for it)
    float *d_in = nullptr, *d_out = nullptr;   // staging for the host-buffer entry points
    size_t staging_floats = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;    // copy streams of the pipelined host path
    std::vector<cudaEvent_t> ev_in, ev_done;
    int64_t launches = 0;
    char err[256] = "";
};

namespace {

thread_local char g_create_err[256] = "";
// Which handle launched the library's most recent kernel on each stream.  The ring-order kernel uses
// it to decide how much of its state is provably older than the kernel in front of it (see
// RingParams::early); every launch path records itself here.
std::mutex g_stream_mu;
std::unordered_map<cudaStream_t, const void *> g_last_on_stream;
// Caller buffers of the library's flag-mode launches per stream since the last launch that waited for
// the whole stream (grid mode / plain): flag mode lets launches overlap beyond their immediate
// predecessor, so a new call whose input aliases a recent output (handles chained through a buffer)
// or whose output aliases a recent input falls back to grid mode.  How far back can a launch
// overlap?  Every flag-mode kernel ends with griddepcontrol.wait, so its CTAs stay resident until the
// launch before it has completed: launches k .. k + d can only be in flight together while all their
// CTAs fit the device, i.e. d <= (2 CTAs x SMs) / (CTAs per launch) <= 2 x SMs.  The window keeps
// that many entries; when it is full the next launch runs in grid mode (waits for everything) and the
// window restarts.
struct RecentIo { const char *in_lo, *in_hi, *out_lo, *out_hi; const void *handle; };
std::unordered_map<cudaStream_t, std::vector<RecentIo>> g_recent_io;
// whether the library's previous launch on the stream ran in flag mode: such a kernel releases its
// dependents without waiting, so "older than the kernel in front of us" no longer means "complete"
// and a grid-mode launch behind it must not load anything before its griddepcontrol.wait
std::unordered_map<cudaStream_t, bool> g_last_flag_mode;

// Timing experiments of profiles/*.sh: compiled in only with -DPVB_EXPERIMENTS (make variants), read
// from the environment once per process.  The default build has none of them.
#ifdef PVB_EXPERIMENTS
struct Experiments {
    int skip = 0;          // PVB_SKIP: phase-skipping timing experiments of the ring kernel (WRONG results)
    int early = -1;        // PVB_EARLY=0/1/2: cap on the ring kernel's pre-wait state loads
    int ring_pad_kb = 0;   // PVB_RING_PAD_KB: extra dynamic shared memory per CTA (occupancy experiment)
    int ring_wpc = 0;      // PVB_RING_WPC: pairs per CTA of the ring kernel (0: default)
    int stagger_ns = 0;    // PVB_STAGGER_NS: start offset between warps sharing an SM
    bool no_aligned = false;   // PVB_NO_ALIGNED=1: warp kernel without the hop % 128 == 0 specialisation
    int trace = 0;         // PVB_TRACE=n: %globaltimer start / end stamps of the first n ring launches of the process
    Experiments() {
        auto geti = [](const char *name, int dflt) { const char *e = std::getenv(name); return e ? std::atoi(e) : dflt; };
        skip = geti("PVB_SKIP", 0);
        early = geti("PVB_EARLY", -1);
        ring_pad_kb = geti("PVB_RING_PAD_KB", 0);
        ring_wpc = geti("PVB_RING_WPC", 0);
        stagger_ns = geti("PVB_STAGGER_NS", 0);
        no_aligned = geti("PVB_NO_ALIGNED", 0) == 1;
        trace = geti("PVB_TRACE", 0);
    }
};
const Experiments &experiments() { static const Experiments e; return e; }
unsigned long long *g_trace_buf = nullptr;
int g_trace_next = 0;
#else
struct Experiments {
    static constexpr int skip = 0, early = -1, ring_pad_kb = 0, ring_wpc = 0, stagger_ns = 0;
    static constexpr bool no_aligned = false;
    static constexpr int trace = 0;
};
constexpr Experiments experiments() { return Experiments(); }
#endif

// pitch_factor == mant * 2^-shift exactly; shift outside [1, 62] -> 0 (kernel uses float64)
void split_pitch_factor(float pf, int *mant, int *shift) {
    *mant = 0;
    *shift = 0;
    if (!std::isfinite(pf) || pf == 0.0f) return;
    int e = 0;
    const float fr = std::frexp(pf, &e);               // pf = fr * 2^e, |fr| in [0.5, 1)
    const int m = int(std::ldexp(fr, 24));             // exact: 24-bit mantissa
    const int sh = 24 - e;
    if (sh < 1 || sh > 62) return;
    *mant = m;
    *shift = sh;
}

int fail(pvb_processor *p, int code, const char *fmt, ...) {
    char *dst = p ? p->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    return code;
}

#define PVB_CUDA(p, call)                                                              \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess)                                                         \
            return fail((p), PVB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

}  // namespace

// Device resources of the peak guard's exact path, shared by all handles of one frame size on one
// device: fft.js's own float64 tables (bundle:13-17, 31-38) and a pool of scratch slots of 2N doubles
// (one per channel pair that can be resident on the device at a time; L2-resident in practice).
struct ExactPool {
    double *pool = nullptr, *tw = nullptr;
    unsigned *locks = nullptr;
    int *rev = nullptr;
    int slots = 0, refs = 0;
};

namespace {

std::mutex g_pool_mu;
std::map<std::pair<int, int>, ExactPool> g_pools;       // (device, frame size)

int ring_max_pairs(int n) {
    switch (n) {
        case 256: return pvb::RingGeoT<256>::MAX_PAIRS;
        case 512: return pvb::RingGeoT<512>::MAX_PAIRS;
        case 1024: return pvb::RingGeoT<1024>::MAX_PAIRS;
        case 2048: return pvb::RingGeoT<2048>::MAX_PAIRS;
    }
    return pvb::RingGeoT<4096>::MAX_PAIRS;
}

ExactPool *acquire_exact_pool(int dev, int n, int num_sms) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    ExactPool &e = g_pools[{dev, n}];
    if (e.refs++ > 0) return &e;
    // function FFT(size), bundle:4-43, in the same float64 arithmetic (libm cos / sin like the JS Math object)
    std::vector<double> tw(2 * size_t(n));
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < 2 * n; i += 2) {
        const double angle = pi * i / n;
        tw[i] = std::cos(angle);
        tw[i + 1] = -std::sin(angle);
    }
    int power = 0;
    for (int t = 1; n > t; t <<= 1) power++;
    const int width = (power % 2 == 0) ? power - 1 : power;
    std::vector<int> rev(size_t(1) << width);
    for (int j = 0; j < (1 << width); j++) {
        int32_t r = 0;
        for (int shift = 0; shift < width; shift += 2) {
            const int back = width - shift - 2;                     // may be -1: JS masks shift counts to 5 bits
            r |= int32_t(uint32_t((j >> shift) & 3) << (back & 31));
        }
        rev[j] = r;
    }
    e.slots = 2 * num_sms * ring_max_pairs(n) + 8;
    const size_t pool_bytes = size_t(e.slots) * 2 * size_t(n) * sizeof(double);
    bool ok = cudaMalloc(&e.pool, pool_bytes) == cudaSuccess &&
              cudaMalloc(&e.locks, size_t(e.slots) * sizeof(unsigned)) == cudaSuccess &&
              cudaMalloc(&e.tw, tw.size() * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&e.rev, rev.size() * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMemset(e.locks, 0, size_t(e.slots) * sizeof(unsigned)) == cudaSuccess &&
         cudaMemcpy(e.tw, tw.data(), tw.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(e.rev, rev.data(), rev.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        cudaFree(e.pool); cudaFree(e.locks); cudaFree(e.tw); cudaFree(e.rev);
        g_pools.erase({dev, n});
        return nullptr;
    }
    return &e;
}

void release_exact_pool(int dev, int n) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pools.find({dev, n});
    if (it == g_pools.end() || --it->second.refs > 0) return;
    cudaFree(it->second.pool); cudaFree(it->second.locks); cudaFree(it->second.tw); cudaFree(it->second.rev);
    g_pools.erase(it);
}

bool valid_frame(int n) { return n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096; }

template <int N>
cudaError_t launch_n(const pvb::FrameParams &fp, cudaStream_t s) {
    using G = pvb::Geo<N>;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pvb::pv_process_kernel<N>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(G::SMEM_BYTES));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int pairs = (fp.num_channels + 1) / 2;
    const int grid = (pairs + G::G - 1) / G::G;
    if (grid == 0) return cudaSuccess;
    pvb::pv_process_kernel<N><<<grid, G::THREADS, G::SMEM_BYTES, s>>>(fp);
    return cudaGetLastError();
}

// in-place shift kernels (pv_kernel_warp.cuh, pv_kernel_pair.cuh, pv_kernel_cta.cuh): pitch factors
// in [0.75, 64] (first stale level only, at most two sources per bin), R <= 32
bool fast_range(const pvb::FrameParams &fp) {
    return fp.pf_shift >= 1 && fp.pitch_factor >= 0.75f && fp.pitch_factor <= 64.0f && fp.overlaps <= 32;
}
bool warp_kernel_applies(int n, const pvb::FrameParams &fp) { return n == 1024 && fast_range(fp); }
// ring-order kernel, DEEP instances (pv_kernel_ring.cuh): scalar pitch factors in [0.33, 0.75) -- stale slots up
// to N/2 + N/3 rebuilt from the spectrum / the windowed frame; colliding regions in three ordered sub-steps
// (from 0.5) or through shared-memory atomics (below)
bool deep_range(int n, const pvb::FrameParams &fp) {
    (void)n;
    return !fp.pf_ch && fp.pf_shift >= 1 && fp.pitch_factor >= 0.33f && fp.pitch_factor < 0.75f && fp.overlaps <= 32;
}

template <int N>
cudaError_t launch_cta_n(const pvb::FrameParams &fp, const float *window_out, cudaStream_t s) {
    using G = pvb::CtaGeo<N>;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pvb::pv_process_cta_kernel<N>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(G::SMEM_BYTES));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int pairs = (fp.num_channels + 1) / 2;
    const int grid = (pairs + G::G - 1) / G::G;
    if (grid == 0) return cudaSuccess;
    pvb::pv_process_cta_kernel<N><<<grid, G::THREADS, G::SMEM_BYTES, s>>>(fp, window_out);
    return cudaGetLastError();
}

cudaError_t launch_cta(int n, const pvb::FrameParams &fp, const float *window_out, cudaStream_t s) {
    switch (n) {
        case 256: return launch_cta_n<256>(fp, window_out, s);
        case 512: return launch_cta_n<512>(fp, window_out, s);
        case 1024: return launch_cta_n<1024>(fp, window_out, s);
        case 2048: return launch_cta_n<2048>(fp, window_out, s);
        case 4096: return launch_cta_n<4096>(fp, window_out, s);
    }
    return cudaErrorInvalidValue;
}

// warps (channel pairs) per CTA: the choice that leaves the most even load per SM
int pick_warps_per_cta(int pairs, int num_sms) {
    int best = 7;
    long best_cost = -1;
    for (int w = 4; w <= pvb::WarpGeo::MAX_WARPS; w++) {
        const long ctas = (pairs + w - 1) / w;
        const long cost = ((ctas + num_sms - 1) / num_sms) * w;     // warps on the busiest SM
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && w > best)) {
            best = w;
            best_cost = cost;
        }
    }
    return best;
}

cudaError_t launch_warp(const pvb::FrameParams &fp, const float *window_out, int num_sms,
                        cudaStream_t s) {
    using W = pvb::WarpGeo;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pvb::pv_process_warp_kernel<true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(W::MAX_WARPS * W::WARP_BYTES));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(pvb::pv_process_warp_kernel<false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(W::MAX_WARPS * W::WARP_BYTES));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int pairs = (fp.num_channels + 1) / 2;
    if (pairs == 0) return cudaSuccess;
    const int wpc = pick_warps_per_cta(pairs, num_sms);
    const int grid = (pairs + wpc - 1) / wpc;
    pvb::WarpParams wp;
    wp.f = fp;
    wp.window_out = window_out;
    wp.num_sms = num_sms;
    wp.stagger_ns = (grid <= 2 * num_sms) ? experiments().stagger_ns : 0;
    if (fp.hop % 128 == 0 && !experiments().no_aligned)
        pvb::pv_process_warp_kernel<true><<<grid, wpc * 32, size_t(wpc) * W::WARP_BYTES, s>>>(wp);
    else
        pvb::pv_process_warp_kernel<false><<<grid, wpc * 32, size_t(wpc) * W::WARP_BYTES, s>>>(wp);
    return cudaGetLastError();
}

// ring-order kernel (pv_kernel_ring.cuh): paired state layout aligned to the time cursor
bool ring_geometry_ok(const pvb_processor *h) {
    // frame 4096 keeps the frame blocks of one parity per thread: the hop must be a multiple of 256
    // (frame 256 keeps roles in units of 64 samples: hop 64 or 128)
    return h->hop % (h->n == 4096 ? 256 : h->n == 256 ? 64 : 128) == 0 && h->hop <= h->n / 2;
}

enum KernelFamily { K_RING = 1, K_WARP = 2, K_CTA = 3, K_GENERIC = 4 };

// the kernel family a call with these parameters runs on (PVB_OPT_KERNEL: first family to try)
KernelFamily pick_kernel(const pvb_processor *h, const pvb::FrameParams &fp) {
    const int first = h->opt_kernel ? h->opt_kernel : K_RING;
    if (fp.pf_ch) {
        // per-channel pitch factors: the ring-order kernel when every channel is in its range, else the
        // generic kernel (the only other family that reads pf_ch)
        if (first <= K_RING && (h->pf_fast || h->pf_deep) && fp.overlaps <= 32 && ring_geometry_ok(h)) return K_RING;
        return K_GENERIC;
    }
    if (first <= K_RING && (fast_range(fp) || deep_range(h->n, fp)) && ring_geometry_ok(h)) return K_RING;
    if (first <= K_WARP && warp_kernel_applies(h->n, fp)) return K_WARP;
    if (first <= K_CTA && fast_range(fp)) return K_CTA;
    return K_GENERIC;
}
bool ring_kernel_applies(const pvb_processor *h, const pvb::FrameParams &fp) { return pick_kernel(h, fp) == K_RING; }

size_t state_rows(int channels);

// everything of a ring-order launch that does not depend on the frame size (one call per launch:
// it records the caller's buffers and decides between flag mode and grid mode).
// `input_ready`: the caller's input is known to be complete and visible before this launch can start
// (see PVB_OPT_INPUTS_READY; always true for the library's own staging buffers and for the second and
// later launches of one submission, which start behind a launch that waited for the whole stream or
// behind one that already had the guarantee).
pvb::RingParams make_ring_params(const pvb_processor *h, const pvb::FrameParams &fp, cudaStream_t s,
                                 bool input_ready, int num_hops) {
    pvb::RingParams rp;
    rp.num_hops = num_hops;
    rp.in = fp.in;
    rp.out = fp.out;
    rp.hist2 = reinterpret_cast<float4 *>(fp.hist);
    rp.acc2 = reinterpret_cast<float4 *>(fp.acc);
    rp.window2 = h->d_window;
    rp.window_out2 = h->d_window_out;
    rp.gtab = h->d_ring_tab;
    rp.num_channels = fp.num_channels;
    rp.hop = fp.hop;
    rp.tmod = fp.step_mod_r * fp.hop;
    rp.pitch_factor = fp.pitch_factor;
    rp.pf_mant = fp.pf_mant;
    rp.pf_shift = fp.pf_shift;
    rp.pf_ch = fp.pf_ch;
    rp.stagger_ns = experiments().stagger_ns;
    rp.skip = experiments().skip;
    const bool pdl = h->opt_launch_mode != 2;
    {
        // early state loads: 2 when another handle's kernel (which passed its own wait before it let
        // us launch) sits between this handle's previous call and this one, else 1
        std::lock_guard<std::mutex> lk(g_stream_mu);
        auto it = g_last_on_stream.find(s);
        rp.early = (it != g_last_on_stream.end() && it->second != h) ? 2 : 1;
        if (!pdl) rp.early = 0;
        if (experiments().early >= 0 && rp.early > experiments().early) rp.early = experiments().early;
        // flag mode needs the caller's buffers to be disjoint from those of every launch it may overlap
        const size_t io_bytes = size_t(num_hops) * size_t(fp.num_channels) * size_t(fp.hop) * sizeof(float);
        RecentIo io;
        io.in_lo = reinterpret_cast<const char *>(fp.in);
        io.in_hi = fp.in ? io.in_lo + io_bytes : io.in_lo;
        io.out_lo = reinterpret_cast<const char *>(fp.out);
        io.out_hi = io.out_lo + io_bytes;
        io.handle = h;
        auto &recent = g_recent_io[s];
        bool safe = pdl && h->opt_launch_mode == 0 && input_ready;
        for (const RecentIo &r : recent) {
            if (io.in_lo < r.out_hi && r.out_lo < io.in_hi) safe = false;      // reads what a recent call writes
            if (io.out_lo < r.in_hi && r.in_lo < io.out_hi) safe = false;      // writes what a recent call reads
            // writes what a recent call writes: ordered only when it is the same handle writing the same
            // rows (every pair waits for its own previous call)
            if (io.out_lo < r.out_hi && r.out_lo < io.out_hi && !(r.handle == h && r.out_lo == io.out_lo))
                safe = false;
        }
        if (recent.size() >= size_t(2 * h->num_sms + 8)) safe = false;         // window full (see g_recent_io)
        if (!safe) recent.clear();          // a grid-mode launch waits for everything before it
        recent.push_back(io);
        rp.flag_mode = safe ? 1 : 0;
        bool &last_flag = g_last_flag_mode[s];
        if (last_flag) rp.early = 0;
        last_flag = safe;
    }
    // PVB_OPT_PEAK_GUARD: 0 auto (five or more uncertain comparisons: one natural near-tie between two
    // neighbouring bins makes two), 1 off, 2 always, 3 strict (one or more)
    rp.guard_min = h->opt_peak_guard == 1 ? 0x7fffffff : h->opt_peak_guard == 2 ? 0 : h->opt_peak_guard == 3 ? 1 : 5;
    rp.xtw = h->xpool->tw;
    rp.xrev = h->xpool->rev;
    rp.xpool = h->xpool->pool;
    rp.xlocks = h->xpool->locks;
    rp.xslots = h->xpool->slots;
    rp.xcount = h->d_xcount;
    rp.stamps = nullptr;
#ifdef PVB_EXPERIMENTS
    if (experiments().trace > 0) {
        // one (start, end) pair per launch, in launch order; pvb_trace_dump() prints them
        std::lock_guard<std::mutex> lk(g_stream_mu);
        if (!g_trace_buf) {
            cudaMalloc(&g_trace_buf, size_t(experiments().trace) * 2 * sizeof(unsigned long long));
            std::vector<unsigned long long> init(size_t(experiments().trace) * 2);
            for (size_t i = 0; i < init.size(); i += 2) { init[i] = ~0ull; init[i + 1] = 0; }
            cudaMemcpy(g_trace_buf, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
        }
        if (g_trace_next < experiments().trace) rp.stamps = g_trace_buf + 2 * size_t(g_trace_next++);
    }
#endif
    rp.done = h->d_done;
    rp.err = h->d_err;
    rp.wait_seq = h->ring_seq;
    rp.my_seq = h->ring_seq + 1;
    return rp;
}

// The ring-order kernel instances live in pv_ring_inst.cu (one translation unit per frame size).
// Pairs per CTA: frame 1024 balances one wave (4..7 warps); the other sizes fill their CTAs (frame 256:
// 32 quarter-warp pairs, 512: 16 half-warp pairs, 2048: 4 pairs of two warps, 4096: 2 pairs of four).
cudaError_t launch_ring(const pvb_processor *h, const pvb::FrameParams &fp, cudaStream_t s, bool input_ready,
                        int num_hops) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = pvb::ring_configure_256();
        if (e == cudaSuccess) e = pvb::ring_configure_512();
        if (e == cudaSuccess) e = pvb::ring_configure_1024();
        if (e == cudaSuccess) e = pvb::ring_configure_2048();
        if (e == cudaSuccess) e = pvb::ring_configure_4096();
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    pvb::RingLaunch l;
    l.pairs = (fp.num_channels + 1) / 2;
    if (l.pairs == 0) return cudaSuccess;
    // (builds with more than 7 pairs per CTA at frame 1024, -DPVB_RING_PAIRS_1024: always full CTAs)
    l.ppc = (h->n == 1024 && pvb::RingGeoT<1024>::MAX_PAIRS <= 7) ? pick_warps_per_cta(l.pairs, h->num_sms) : 0;
    if (experiments().ring_wpc > 0) l.ppc = experiments().ring_wpc;           // PVB_RING_WPC
    l.pad_kb = experiments().ring_pad_kb;                                     // PVB_RING_PAD_KB
    l.pdl = h->opt_launch_mode != 2;
    l.pch = fp.pf_ch != nullptr;
    l.multi = num_hops > 1;
    // pick_kernel admitted it: 1 = pitch factors down to 0.5, 2 = down to 0.33
    l.deep = fp.pf_ch ? (h->pf_fast ? 0 : h->pf_half ? 1 : 2) : (fast_range(fp) ? 0 : fp.pitch_factor >= 0.5f ? 1 : 2);                                   // pick_kernel admitted it: [0.5, 0.75)
    l.stream = s;
    pvb::RingParams rp = make_ring_params(h, fp, s, input_ready, num_hops);
    const_cast<pvb_processor *>(h)->ring_seq = rp.my_seq;       // the launch below stores it into done[]
    switch (h->n) {
        case 256: return pvb::ring_launch_256(rp, l);
        case 512: return pvb::ring_launch_512(rp, l);
        case 1024: return pvb::ring_launch_1024(rp, l);
        case 2048: return pvb::ring_launch_2048(rp, l);
        case 4096: return pvb::ring_launch_4096(rp, l);
    }
    return cudaErrorInvalidValue;
}

// num_hops > 1 (several consecutive calls in one launch) only with the ring-order kernel
cudaError_t launch_any(const pvb_processor *h, const pvb::FrameParams &fp, cudaStream_t s, bool input_ready,
                       int num_hops) {
    const int n = h->n;
    switch (pick_kernel(h, fp)) {
        case K_RING: return launch_ring(h, fp, s, input_ready, num_hops);
        case K_WARP: return launch_warp(fp, h->d_window_out, h->num_sms, s);
        case K_CTA: return launch_cta(n, fp, h->d_window_out, s);
        case K_GENERIC: break;
    }
    switch (n) {
        case 256: return launch_n<256>(fp, s);
        case 512: return launch_n<512>(fp, s);
        case 1024: return launch_n<1024>(fp, s);
        case 2048: return launch_n<2048>(fp, s);
        case 4096: return launch_n<4096>(fp, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch(const pvb_processor *h, const pvb::FrameParams &fp, cudaStream_t s, bool input_ready,
                   int num_hops = 1) {
    const cudaError_t e = launch_any(h, fp, s, input_ready, num_hops);
    std::lock_guard<std::mutex> lk(g_stream_mu);
    if (!ring_kernel_applies(h, fp)) {       // the other kernels are plain launches: they wait for everything
        g_last_flag_mode[s] = false;
        g_recent_io[s].clear();
    }
    if (g_last_on_stream.size() > 4096) {    // streams come and go; unknown == conservative
        g_last_on_stream.clear();
        g_recent_io.clear();
        g_last_flag_mode.clear();
    }
    g_last_on_stream[s] = h;
    return e;
}

// number of source bins that can land inside [0, nb): for pitchFactor >= 1 only bins
// 0..N/2 (delta >= 0); below 1 the last region reaches up to nb + nb*(1-pf) (pv:133,150).
int source_limit(int n, float pitch_factor) {
    const int nb = n / 2 + 1;
    if (!(pitch_factor < 1.0f)) return nb;
    double lim = double(nb) + std::ceil(double(nb) * (1.0 - double(pitch_factor))) + 2.0;
    if (!(lim < double(n))) return n;      // also catches NaN / negative factors
    if (lim < nb) lim = nb;
    return int(lim);
}

size_t state_rows(int channels) { return size_t(channels > 0 ? ((channels + 1) & ~1) : 2); }

int alloc_state(pvb_processor *p, int channels) {
    cudaFree(p->d_hist);
    cudaFree(p->d_acc);
    cudaFree(p->d_done);
    p->d_hist = p->d_acc = nullptr;
    p->d_done = nullptr;
    p->ring_seq = 0;
    p->channels = channels;
    // rows are padded to an even channel count: the kernels process channels in pairs
    const size_t bytes = state_rows(channels) * size_t(p->n) * sizeof(float);
    if (cudaMalloc(&p->d_hist, bytes) != cudaSuccess || cudaMalloc(&p->d_acc, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(p, PVB_ERR_NOMEM, "cudaMalloc of %zu state bytes failed", 2 * bytes);
    }
    const size_t flag_bytes = (state_rows(channels) / 2) * sizeof(unsigned);
    if (p->h_err) *p->h_err = 0;
    if (cudaMalloc(&p->d_done, flag_bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(p, PVB_ERR_NOMEM, "cudaMalloc of the completion flags failed");
    }
    PVB_CUDA(p, cudaMemsetAsync(p->d_hist, 0, bytes, p->stream));
    PVB_CUDA(p, cudaMemsetAsync(p->d_acc, 0, bytes, p->stream));
    PVB_CUDA(p, cudaMemsetAsync(p->d_done, 0, flag_bytes, p->stream));
    p->ring_calls = 0;
    p->layout = pvb_processor::ZERO;
    return PVB_OK;
}

// Re-lay the state for the other kernel family (rare: the pitch factor left / entered the range
// of the ring-order kernel, or the host touches the state).  Out of place, on stream s.
int ensure_layout(pvb_processor *p, pvb_processor::Layout want, cudaStream_t s) {
    if (p->layout == want || p->channels == 0) return PVB_OK;
    if (p->layout == pvb_processor::ZERO) {
        p->layout = want;
        return PVB_OK;
    }
    const size_t rows = state_rows(p->channels);
    const size_t bytes = rows * size_t(p->n) * sizeof(float);
    float *nh = nullptr, *na = nullptr;
    if (cudaMalloc(&nh, bytes) != cudaSuccess || cudaMalloc(&na, bytes) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(nh);
        return fail(p, PVB_ERR_NOMEM, "cudaMalloc of %zu bytes for the state re-layout failed", 2 * bytes);
    }
    const int rb = int(p->ring_calls % uint64_t(p->overlaps)) * p->hop;
    const int tmod = int(p->cursor_calls % uint64_t(p->overlaps)) * p->hop;
    const int pairs = int(rows / 2);
    const bool to_paired = want == pvb_processor::PAIRED;
    float *planar_h = to_paired ? p->d_hist : nh, *planar_a = to_paired ? p->d_acc : na;
    float *paired_h = to_paired ? nh : p->d_hist, *paired_a = to_paired ? na : p->d_acc;
    pvb::pv_ring_convert_kernel<<<p->num_sms * 8, 256, 0, s>>>(
        planar_h, planar_a, reinterpret_cast<float2 *>(paired_h), reinterpret_cast<float2 *>(paired_a),
        pairs, p->n, p->hop, rb, tmod, to_paired ? 1 : 0);
    PVB_CUDA(p, cudaGetLastError());
    PVB_CUDA(p, cudaStreamSynchronize(s));
    cudaFree(p->d_hist);
    cudaFree(p->d_acc);
    p->d_hist = nh;
    p->d_acc = na;
    p->layout = want;
    return PVB_OK;
}

int ensure_staging(pvb_processor *p, size_t floats) {
    if (floats <= p->staging_floats) return PVB_OK;
    PVB_CUDA(p, cudaStreamSynchronize(p->stream));
    cudaFree(p->d_in);
    cudaFree(p->d_out);
    p->d_in = p->d_out = nullptr;
    p->staging_floats = 0;
    if (cudaMalloc(&p->d_in, floats * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&p->d_out, floats * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return fail(p, PVB_ERR_NOMEM, "cudaMalloc of staging buffers failed");
    }
    p->staging_floats = floats;
    return PVB_OK;
}

// wait for everything this handle has submitted: its own stream and the caller stream of the most
// recent device submission
cudaError_t sync_all(pvb_processor *p) {
    cudaError_t e = cudaStreamSynchronize(p->stream);
    if (e == cudaSuccess && p->last_stream && p->last_stream != p->stream) e = cudaStreamSynchronize(p->last_stream);
    return e;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// hooks of the pipelined host path: run before / after the launch of call k
struct CallHooks {
    virtual void before(int k, cudaStream_t s) = 0;
    virtual void after(int k, cudaStream_t s) = 0;
    // copy groups: calls [first[k], last[k]) share their copies (they may share a launch too)
    std::vector<int> first, last;
    int left_in_group(int k) const { return last[k] - k; }
};

// sticky device error (a completion flag that never arrived): the handle refuses further work
int check_device_error(pvb_processor *p) {
    if (p->h_err && *reinterpret_cast<volatile unsigned *>(p->h_err) != 0)
        return fail(p, PVB_ERR_CUDA, "ring-order kernel: %u channel pair(s) never saw the completion flag of their "
                    "previous call and were not processed; state is stale (pvb_reset / pvb_resize to recover)",
                    *reinterpret_cast<volatile unsigned *>(p->h_err));
    return PVB_OK;
}

// Per-channel pitch factors: keep the last array on the device; a changed array is copied on the
// launch stream (stream-ordered behind every launch that still reads the old one).
int upload_pitch_factors(pvb_processor *p, const float *pf_host, cudaStream_t s) {
    const size_t c = size_t(p->channels);
    if (c == 0) return PVB_OK;
    if (p->d_pf && p->h_pf.size() == c && std::memcmp(p->h_pf.data(), pf_host, c * sizeof(float)) == 0) return PVB_OK;
    if (!p->d_pf || p->h_pf.size() != c) {
        PVB_CUDA(p, sync_all(p));
        cudaFree(p->d_pf);
        p->d_pf = nullptr;
        if (cudaMalloc(&p->d_pf, state_rows(p->channels) * sizeof(float)) != cudaSuccess) {
            cudaGetLastError();
            return fail(p, PVB_ERR_NOMEM, "cudaMalloc of the pitch-factor array failed");
        }
    }
    p->h_pf.assign(pf_host, pf_host + c);
    p->pf_fast = p->pf_deep = p->pf_half = true;
    for (size_t i = 0; i < c; i++) {
        int m = 0, sh = 0;
        split_pitch_factor(pf_host[i], &m, &sh);
        if (!(sh >= 1 && pf_host[i] >= 0.75f && pf_host[i] <= 64.0f)) p->pf_fast = false;
        if (!(sh >= 1 && pf_host[i] >= 0.33f && pf_host[i] <= 64.0f)) p->pf_deep = false;
        if (!(pf_host[i] >= 0.5f)) p->pf_half = false;
    }
    // (pageable source: the runtime stages it before returning, so h_pf may change right after)
    PVB_CUDA(p, cudaMemcpyAsync(p->d_pf, p->h_pf.data(), c * sizeof(float), cudaMemcpyHostToDevice, s));
    std::lock_guard<std::mutex> lk(g_stream_mu);
    g_last_flag_mode[s] = false;        // the copy waited for everything before it; so does what follows
    g_recent_io[s].clear();
    return PVB_OK;
}

// `inputs_ready`: the input of the FIRST launch is known to be complete before it can start (the
// library's own staging buffers, or PVB_OPT_INPUTS_READY); later launches of the submission start
// behind the first one and inherit the guarantee.  `pf_host`: per-channel pitch factors or nullptr.
int submit(pvb_processor *p, const float *in_dev, float *out_dev, int num_calls, float pf,
           cudaStream_t s, bool inputs_ready, CallHooks *hooks = nullptr, const float *pf_host = nullptr) {
    const size_t block = size_t(p->channels) * size_t(p->hop);
    {
        const int rc = check_device_error(p);
        if (rc != PVB_OK) return rc;
    }
    p->last_stream = s;
    if (pf_host) {
        const int rc = upload_pitch_factors(p, pf_host, s);
        if (rc != PVB_OK) return rc;
        pf = 1.0f;
    }
    // most calls per launch of the MULTI ring-order kernels (a launch of 64 hops at 4096 channels runs ~1 ms)
    constexpr int MAX_HOPS_PER_LAUNCH = 64;
    for (int k = 0; k < num_calls;) {
        pvb::FrameParams fp;
        fp.in = in_dev ? in_dev + size_t(k) * block : nullptr;
        fp.out = out_dev + size_t(k) * block;
        fp.hist = p->d_hist;
        fp.acc = p->d_acc;
        fp.window = p->d_window;
        fp.tw = p->d_tw;
        fp.num_channels = p->channels;
        fp.hop = p->hop;
        fp.overlaps = p->overlaps;
        fp.ring_base = int(p->ring_calls % uint64_t(p->overlaps)) * p->hop;
        fp.step_mod_r = int(p->cursor_calls % uint64_t(p->overlaps));
        fp.src_limit = source_limit(p->n, pf);
        fp.pitch_factor = pf;
        fp.pf_ch = (pf_host && p->channels > 0) ? p->d_pf : nullptr;
        split_pitch_factor(pf, &fp.pf_mant, &fp.pf_shift);
        // how many consecutive calls this launch does: all that remain (up to the end of their copy
        // group on the pipelined host path) when the ring-order kernel takes them, else one
        int hops = 1;
        if (p->channels > 0 && !pf_host && p->opt_many_mode == 1 && ring_kernel_applies(p, fp) && fast_range(fp)) {
            hops = num_calls - k;
            if (hooks && hops > hooks->left_in_group(k)) hops = hooks->left_in_group(k);
            if (hops > MAX_HOPS_PER_LAUNCH) hops = MAX_HOPS_PER_LAUNCH;
        }
        if (hooks) for (int j = k; j < k + hops; j++) hooks->before(j, s);
        if (p->channels > 0) {
            const int rc = ensure_layout(p, ring_kernel_applies(p, fp) ? pvb_processor::PAIRED
                                                                       : pvb_processor::PLANAR, s);
            if (rc != PVB_OK) return rc;
            fp.hist = p->d_hist;
            fp.acc = p->d_acc;
            PVB_CUDA(p, launch(p, fp, s, inputs_ready || k > 0, hops));
            p->launches++;
        }
        p->ring_calls += hops;
        p->cursor_calls += hops;   // pv:71, once per call for all channels
        if (hooks) for (int j = k; j < k + hops; j++) hooks->after(j, s);
        k += hops;
    }
    return PVB_OK;
}

}  // namespace

extern "C" {

int32_t pvb_version(void) { return PVB_VERSION; }

const char *pvb_error_string(int32_t code) {
    switch (code) {
        case PVB_OK: return "ok";
        case PVB_ERR_BAD_SIZE: return "FFT size must be a power of two in [256, 4096] and hop must divide it";
        case PVB_ERR_BAD_ARG: return "bad argument";
        case PVB_ERR_CUDA: return "CUDA error (no CPU fallback exists)";
        case PVB_ERR_NOMEM: return "out of memory";
    }
    return "unknown error";
}

const char *pvb_last_error(const pvb_processor *p) { return p ? p->err : g_create_err; }

int32_t pvb_create(const pvb_config *cfg, pvb_processor **out) {
    if (!cfg || !out) return fail(nullptr, PVB_ERR_BAD_ARG, "pvb_create: NULL argument");
    *out = nullptr;
    const int n = cfg->frame_size ? cfg->frame_size : 2048;    // pv:6
    const int hop = cfg->hop_size ? cfg->hop_size : 128;       // ola:3
    if (!valid_frame(n))
        return fail(nullptr, PVB_ERR_BAD_SIZE, "FFT size must be a power of two in [256, 4096], got %d", n);
    if (hop < 4 || hop > n || n % hop != 0 || hop % 4 != 0)
        return fail(nullptr, PVB_ERR_BAD_SIZE, "hop %d must divide frame %d and be a multiple of 4", hop, n);
    if (cfg->num_channels < 0) return fail(nullptr, PVB_ERR_BAD_ARG, "negative channel count");

    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, PVB_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    int dev = cfg->device;
    if (dev < 0) cudaGetDevice(&dev);
    if (dev >= count) return fail(nullptr, PVB_ERR_BAD_ARG, "device %d out of range (%d devices)", dev, count);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major < 10)
        return fail(nullptr, PVB_ERR_CUDA, "device %d is not sm_100 class (kernels are sm_100a only)", dev);

    pvb_processor *p = new (std::nothrow) pvb_processor();
    if (!p) return fail(nullptr, PVB_ERR_NOMEM, "host allocation failed");
    p->n = n;
    p->hop = hop;
    p->overlaps = n / hop;     // ola:17
    p->device = dev;
    DeviceGuard guard(dev);

    auto bail = [&](int code) { strncpy(g_create_err, p->err, 255); pvb_destroy(p); return code; };
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) {
        fail(p, PVB_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
        return bail(PVB_ERR_CUDA);
    }

    if (cudaHostAlloc(reinterpret_cast<void **>(&p->h_err), sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void **>(&p->d_err), p->h_err, 0) != cudaSuccess) {
        fail(p, PVB_ERR_NOMEM, "allocation of the device-error word failed: %s", cudaGetErrorString(cudaGetLastError()));
        return bail(PVB_ERR_NOMEM);
    }
    *p->h_err = 0;

    // tables, computed in double like the JS and rounded once to float32
    std::vector<float> win(2 * n), win_out(2 * n);    // stored twice: the ring-order kernel reads them rotated
    std::vector<float2> tw(n);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < n; i++) {
        win[i] = float(0.5 * (1 - std::cos(2 * pi * i / n)));                          // pv:10-12
        // synthesis window with 1/N (inverseTransform), the two folded 1/2 of the real-split and
        // 1/nbOverlaps (ola:153) folded in; all powers of two, so the product is exact
        win_out[i] = win[i] * (1.0f / float(2 * n)) * (1.0f / float(p->overlaps));
    }
    for (int i = 0; i < n; i++) {
        win[n + i] = win[i];
        win_out[n + i] = win_out[i];
    }
    for (int j = 0; j < n; j++) {
        double c = std::cos(2 * pi * j / n), s = -std::sin(2 * pi * j / n);
        if (j % (n / 4) == 0) {    // exact quarter turns
            const int qd = j / (n / 4);
            c = (qd == 0) ? 1 : (qd == 2) ? -1 : 0;
            s = (qd == 1) ? -1 : (qd == 3) ? 1 : 0;
        }
        tw[j] = make_float2(float(c), float(s));
    }
    cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaMalloc(&p->d_window, 2 * n * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&p->d_window_out, 2 * n * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&p->d_tw, n * sizeof(float2)) != cudaSuccess) {
        fail(p, PVB_ERR_NOMEM, "cudaMalloc of tables failed");
        return bail(PVB_ERR_NOMEM);
    }
    if (cudaMemcpy(p->d_window, win.data(), 2 * n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->d_window_out, win_out.data(), 2 * n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->d_tw, tw.data(), n * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
        fail(p, PVB_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return bail(PVB_ERR_CUDA);
    }
    if (n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096) {
        const size_t tab_bytes = n == 256 ? pvb::ring_host_table_bytes<256>()
                                 : n == 512 ? pvb::ring_host_table_bytes<512>()
                                 : n == 1024 ? pvb::ring_host_table_bytes<1024>()
                                 : n == 2048 ? pvb::ring_host_table_bytes<2048>() : pvb::ring_host_table_bytes<4096>();
        std::vector<float2> rt(tab_bytes / sizeof(float2));
        if (n == 256) pvb::ring_host_tables<256>(tw.data(), rt.data());
        else if (n == 512) pvb::ring_host_tables<512>(tw.data(), rt.data());
        else if (n == 1024) pvb::ring_host_tables<1024>(tw.data(), rt.data());
        else if (n == 2048) pvb::ring_host_tables<2048>(tw.data(), rt.data());
        else pvb::ring_host_tables<4096>(tw.data(), rt.data());
        if (cudaMalloc(&p->d_ring_tab, tab_bytes) != cudaSuccess ||
            cudaMemcpy(p->d_ring_tab, rt.data(), tab_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
            fail(p, PVB_ERR_CUDA, "ring table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            return bail(PVB_ERR_CUDA);
        }
    }
    p->xpool = acquire_exact_pool(dev, n, p->num_sms);
    if (!p->xpool || cudaMalloc(&p->d_xcount, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(p->d_xcount, 0, sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        fail(p, PVB_ERR_NOMEM, "allocation of the peak-guard scratch pool failed");
        return bail(PVB_ERR_NOMEM);
    }
    int rc = alloc_state(p, cfg->num_channels);
    if (rc != PVB_OK) return bail(rc);
    if (cudaStreamSynchronize(p->stream) != cudaSuccess) {
        fail(p, PVB_ERR_CUDA, "state initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return bail(PVB_ERR_CUDA);
    }
    *out = p;
    return PVB_OK;
}

void pvb_destroy(pvb_processor *p) {
    if (!p) return;
    DeviceGuard guard(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->last_stream && p->last_stream != p->stream) cudaStreamSynchronize(p->last_stream);
    {
        // forget this handle's address: a new handle may be allocated at the same one
        std::lock_guard<std::mutex> lk(g_stream_mu);
        for (auto &kv : g_last_on_stream) if (kv.second == p) kv.second = nullptr;
        for (auto &kv : g_recent_io) for (RecentIo &r : kv.second) if (r.handle == p) r.handle = nullptr;
    }
    cudaFree(p->d_hist);
    cudaFree(p->d_acc);
    cudaFree(p->d_window);
    cudaFree(p->d_window_out);
    cudaFree(p->d_tw);
    cudaFree(p->d_ring_tab);
    cudaFree(p->d_done);
    cudaFree(p->d_in);
    cudaFree(p->d_out);
    if (p->h_err) cudaFreeHost(p->h_err);
    cudaFree(p->d_xcount);
    cudaFree(p->d_pf);
    if (p->xpool) release_exact_pool(p->device, p->n);
    for (cudaEvent_t e : p->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : p->ev_done) cudaEventDestroy(e);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

int32_t pvb_process_many_device(pvb_processor *p, const float *in_dev, float *out_dev,
                                int32_t num_calls, float pitch_factor, void *stream) {
    if (!p) return PVB_ERR_BAD_ARG;
    if (!out_dev || num_calls < 0) return fail(p, PVB_ERR_BAD_ARG, "pvb_process: bad argument");
    DeviceGuard guard(p->device);
    return submit(p, in_dev, out_dev, num_calls, pitch_factor,
                  stream ? static_cast<cudaStream_t>(stream) : p->stream, p->opt_inputs_ready != 0);
}

int32_t pvb_process_device(pvb_processor *p, const float *in_dev, float *out_dev,
                           float pitch_factor, void *stream) {
    return pvb_process_many_device(p, in_dev, out_dev, 1, pitch_factor, stream);
}

int32_t pvb_process_many(pvb_processor *p, const float *in, float *out, int32_t num_calls,
                         float pitch_factor) {
    if (!p) return PVB_ERR_BAD_ARG;
    if (!out || num_calls < 0) return fail(p, PVB_ERR_BAD_ARG, "pvb_process: bad argument");
    DeviceGuard guard(p->device);
    const size_t block = size_t(p->channels) * size_t(p->hop);
    const size_t floats = block * size_t(num_calls);
    if (floats == 0) {
        p->ring_calls += num_calls;
        p->cursor_calls += num_calls;
        return PVB_OK;
    }
    int rc = ensure_staging(p, floats);
    if (rc != PVB_OK) return rc;
    if (num_calls == 1) {
        if (in) PVB_CUDA(p, cudaMemcpyAsync(p->d_in, in, floats * sizeof(float), cudaMemcpyHostToDevice, p->stream));
        // (the input copy precedes the launch on the same stream: a copy engine operation, complete
        // before the kernel is eligible)
        rc = submit(p, in ? p->d_in : nullptr, p->d_out, 1, pitch_factor, p->stream, true);
        if (rc != PVB_OK) return rc;
        PVB_CUDA(p, cudaMemcpyAsync(out, p->d_out, floats * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
        PVB_CUDA(p, cudaStreamSynchronize(p->stream));
        return check_device_error(p);
    }
    // several calls: the input copy of call k+1, the kernel of call k and the output copy of
    // call k-1 overlap (three streams chained by events); results are those of K single calls
    if (!p->s_in) {
        PVB_CUDA(p, cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
        PVB_CUDA(p, cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    }
    while (int(p->ev_in.size()) < num_calls) {
        cudaEvent_t a, b;
        PVB_CUDA(p, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        PVB_CUDA(p, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        p->ev_in.push_back(a);
        p->ev_done.push_back(b);
    }
    // copies move `group` calls at a time (16 MiB chunks at the default workload run the PCIe
    // link ~5 % faster than 4 MiB ones)
    struct Pipe : CallHooks {
        pvb_processor *p; const float *in; float *out; size_t block; int calls;
        cudaError_t err = cudaSuccess;
        void note(cudaError_t e) { if (err == cudaSuccess) err = e; }
        void before(int k, cudaStream_t s) override {
            if (in && first[k] == k) {
                const int n = last[k] - k;
                note(cudaMemcpyAsync(p->d_in + k * block, in + k * block, n * block * sizeof(float),
                                     cudaMemcpyHostToDevice, p->s_in));
                note(cudaEventRecord(p->ev_in[k], p->s_in));
                note(cudaStreamWaitEvent(s, p->ev_in[k], 0));
            }
        }
        void after(int k, cudaStream_t s) override {
            if (k + 1 == last[k]) {
                const int f = first[k];
                note(cudaEventRecord(p->ev_done[k], s));
                note(cudaStreamWaitEvent(p->s_out, p->ev_done[k], 0));
                note(cudaMemcpyAsync(out + f * block, p->d_out + f * block,
                                     (k + 1 - f) * block * sizeof(float), cudaMemcpyDeviceToHost, p->s_out));
            }
        }
    } pipe;
    pipe.p = p; pipe.in = in; pipe.out = out; pipe.block = block; pipe.calls = num_calls;
    int group = (block * sizeof(float) >= (size_t(16) << 20)) ? 1
                : int((size_t(16) << 20) / (block * sizeof(float)));
    if (group > 8) group = 8;
    {
        // the pipeline fills with the first input group and drains with the last output group: those ramp
        // (1, 1, 2, ... calls) so that only a single call's copy is exposed at either end of a submission
        std::vector<int> ramp, sizes;
        int ramp_sum = 0;
        for (int r = 1, i = 0; r < group; i++) {           // 1, 1, 2, 4, ... (< group)
            ramp.push_back(r);
            ramp_sum += r;
            if (i >= 1) r *= 2;
        }
        if (num_calls < 2 * ramp_sum + 2 * group) { ramp.clear(); ramp_sum = 0; }
        sizes = ramp;
        for (int left = num_calls - 2 * ramp_sum; left > 0; left -= group) sizes.push_back(left < group ? left : group);
        for (size_t i = ramp.size(); i-- > 0;) sizes.push_back(ramp[i]);
        pipe.first.resize(num_calls);
        pipe.last.resize(num_calls);
        int k0 = 0;
        for (int n : sizes) {
            for (int j = k0; j < k0 + n; j++) { pipe.first[j] = k0; pipe.last[j] = k0 + n; }
            k0 += n;
        }
    }
    // (every launch waits for the event of its input copy before it becomes eligible)
    rc = submit(p, in ? p->d_in : nullptr, p->d_out, num_calls, pitch_factor, p->stream, true, &pipe);
    if (rc != PVB_OK) return rc;
    PVB_CUDA(p, pipe.err);
    PVB_CUDA(p, cudaStreamSynchronize(p->s_out));
    PVB_CUDA(p, cudaStreamSynchronize(p->stream));
    return check_device_error(p);
}

int32_t pvb_process(pvb_processor *p, const float *in, float *out, float pitch_factor) {
    return pvb_process_many(p, in, out, 1, pitch_factor);
}

int32_t pvb_process_pf_device(pvb_processor *p, const float *in_dev, float *out_dev,
                              const float *pitch_factors, void *stream) {
    if (!p) return PVB_ERR_BAD_ARG;
    if (!out_dev || !pitch_factors) return fail(p, PVB_ERR_BAD_ARG, "pvb_process_pf: bad argument");
    DeviceGuard guard(p->device);
    return submit(p, in_dev, out_dev, 1, 1.0f, stream ? static_cast<cudaStream_t>(stream) : p->stream,
                  p->opt_inputs_ready != 0, nullptr, pitch_factors);
}

int32_t pvb_process_pf(pvb_processor *p, const float *in, float *out, const float *pitch_factors) {
    if (!p) return PVB_ERR_BAD_ARG;
    if (!out || !pitch_factors) return fail(p, PVB_ERR_BAD_ARG, "pvb_process_pf: bad argument");
    DeviceGuard guard(p->device);
    const size_t floats = size_t(p->channels) * size_t(p->hop);
    if (floats == 0) {
        p->ring_calls++;
        p->cursor_calls++;
        return PVB_OK;
    }
    int rc = ensure_staging(p, floats);
    if (rc != PVB_OK) return rc;
    if (in) PVB_CUDA(p, cudaMemcpyAsync(p->d_in, in, floats * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    rc = submit(p, in ? p->d_in : nullptr, p->d_out, 1, 1.0f, p->stream, true, nullptr, pitch_factors);
    if (rc != PVB_OK) return rc;
    PVB_CUDA(p, cudaMemcpyAsync(out, p->d_out, floats * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
    PVB_CUDA(p, cudaStreamSynchronize(p->stream));
    return check_device_error(p);
}

int32_t pvb_sync(pvb_processor *p) {
    if (!p) return PVB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    PVB_CUDA(p, sync_all(p));
    return check_device_error(p);
}

int32_t pvb_resize(pvb_processor *p, int32_t num_channels) {
    if (!p) return PVB_ERR_BAD_ARG;
    if (num_channels < 0) return fail(p, PVB_ERR_BAD_ARG, "negative channel count");
    DeviceGuard guard(p->device);
    PVB_CUDA(p, sync_all(p));
    cudaFree(p->d_in);
    cudaFree(p->d_out);
    p->d_in = p->d_out = nullptr;
    p->staging_floats = 0;
    cudaFree(p->d_pf);
    p->d_pf = nullptr;
    p->h_pf.clear();
    int rc = alloc_state(p, num_channels);     // ola:54-88: fresh zeroed buffers (also clears the error word)
    if (rc != PVB_OK) return rc;
    PVB_CUDA(p, cudaStreamSynchronize(p->stream));
    return PVB_OK;
}

int32_t pvb_reset(pvb_processor *p) {
    if (!p) return PVB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    const size_t bytes = state_rows(p->channels) * size_t(p->n) * sizeof(float);
    PVB_CUDA(p, sync_all(p));
    PVB_CUDA(p, cudaMemsetAsync(p->d_hist, 0, bytes, p->stream));
    PVB_CUDA(p, cudaMemsetAsync(p->d_acc, 0, bytes, p->stream));
    PVB_CUDA(p, cudaMemsetAsync(p->d_done, 0, (state_rows(p->channels) / 2) * sizeof(unsigned), p->stream));
    PVB_CUDA(p, cudaStreamSynchronize(p->stream));
    p->ring_seq = 0;
    if (p->h_err) *p->h_err = 0;
    p->ring_calls = 0;
    p->cursor_calls = 0;
    p->layout = pvb_processor::ZERO;
    return PVB_OK;
}

int32_t pvb_frame_size(const pvb_processor *p) { return p ? p->n : PVB_ERR_BAD_ARG; }
int32_t pvb_hop_size(const pvb_processor *p) { return p ? p->hop : PVB_ERR_BAD_ARG; }
int32_t pvb_num_channels(const pvb_processor *p) { return p ? p->channels : PVB_ERR_BAD_ARG; }
double pvb_time_cursor(const pvb_processor *p) { return p ? double(p->cursor_calls) * p->hop : 0.0; }
int64_t pvb_kernel_launches(const pvb_processor *p) { return p ? p->launches : 0; }

int64_t pvb_ring_stuck_count(pvb_processor *p) {
    if (!p || !p->h_err) return 0;
    DeviceGuard guard(p->device);
    if (sync_all(p) != cudaSuccess) return -1;
    return int64_t(*reinterpret_cast<volatile unsigned *>(p->h_err));
}

int64_t pvb_peak_guard_count(pvb_processor *p) {
    if (!p || !p->d_xcount) return 0;
    DeviceGuard guard(p->device);
    unsigned long long v = 0;
    if (sync_all(p) != cudaSuccess) return -1;
    if (cudaMemcpy(&v, p->d_xcount, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return int64_t(v);
}

#ifdef PVB_EXPERIMENTS
// experiment builds only: copy the launch stamps out (ns; start = earliest CTA start, end = latest CTA end)
PVB_API int32_t pvb_trace_dump(unsigned long long *out, int32_t capacity) {
    cudaDeviceSynchronize();
    const int n = g_trace_next < capacity ? g_trace_next : capacity;
    if (n > 0) cudaMemcpy(out, g_trace_buf, size_t(n) * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return n;
}
#endif

int32_t pvb_set_option(pvb_processor *p, int32_t option, int64_t value) {
    if (!p) return PVB_ERR_BAD_ARG;
    switch (option) {
        case PVB_OPT_KERNEL:
            if (value < 0 || value > 4) break;
            p->opt_kernel = int(value);
            return PVB_OK;
        case PVB_OPT_LAUNCH_MODE:
            if (value < 0 || value > 2) break;
            p->opt_launch_mode = int(value);
            return PVB_OK;
        case PVB_OPT_INPUTS_READY:
            if (value < 0 || value > 1) break;
            p->opt_inputs_ready = int(value);
            return PVB_OK;
        case PVB_OPT_PEAK_GUARD:
            if (value < 0 || value > 3) break;
            p->opt_peak_guard = int(value);
            return PVB_OK;
        case PVB_OPT_MANY_MODE:
            if (value < 0 || value > 1) break;
            p->opt_many_mode = int(value);
            return PVB_OK;
    }
    return fail(p, PVB_ERR_BAD_ARG, "pvb_set_option: unknown option %d or bad value %lld", int(option), (long long)value);
}

int64_t pvb_get_option(const pvb_processor *p, int32_t option) {
    if (!p) return PVB_ERR_BAD_ARG;
    switch (option) {
        case PVB_OPT_KERNEL: return p->opt_kernel;
        case PVB_OPT_LAUNCH_MODE: return p->opt_launch_mode;
        case PVB_OPT_INPUTS_READY: return p->opt_inputs_ready;
        case PVB_OPT_PEAK_GUARD: return p->opt_peak_guard;
        case PVB_OPT_MANY_MODE: return p->opt_many_mode;
    }
    return PVB_ERR_BAD_ARG;
}

const char *pvb_kernel_name(const pvb_processor *p, float pitch_factor) {
    if (!p) return "";
    pvb::FrameParams fp{};
    fp.pf_ch = nullptr;
    fp.pitch_factor = pitch_factor;
    fp.overlaps = p->overlaps;
    split_pitch_factor(pitch_factor, &fp.pf_mant, &fp.pf_shift);
    static const char *cta[] = {"pvb::pv_process_cta_kernel<256>", "pvb::pv_process_cta_kernel<512>",
                                "pvb::pv_process_cta_kernel<1024>", "pvb::pv_process_cta_kernel<2048>",
                                "pvb::pv_process_cta_kernel<4096>"};
    static const char *gen[] = {"pvb::pv_process_kernel<256>", "pvb::pv_process_kernel<512>",
                                "pvb::pv_process_kernel<1024>", "pvb::pv_process_kernel<2048>",
                                "pvb::pv_process_kernel<4096>"};
    const int idx = p->n == 256 ? 0 : p->n == 512 ? 1 : p->n == 1024 ? 2 : p->n == 2048 ? 3 : 4;
    switch (pick_kernel(p, fp)) {
        case K_RING: return fast_range(fp) ? "pvb::pv_process_ring_kernel"
                            : pitch_factor >= 0.5f ? "pvb::pv_process_ring_kernel (deep)" : "pvb::pv_process_ring_kernel (deep, atomics)";
        case K_WARP: return "pvb::pv_process_warp_kernel";
        case K_CTA: return cta[idx];
        case K_GENERIC: break;
    }
    return gen[idx];
}

int32_t pvb_set_time_cursor(pvb_processor *p, double samples) {
    if (!p) return PVB_ERR_BAD_ARG;
    // the reference only ever holds multiples of hopSize here (pv:71)
    const double calls = samples / p->hop;
    if (!(samples >= 0) || calls != std::floor(calls) || calls > 9.0e15)
        return fail(p, PVB_ERR_BAD_ARG, "timeCursor must be a non-negative multiple of the hop size");
    if (p->layout == pvb_processor::PAIRED) {       // the paired rings are aligned to the cursor
        DeviceGuard guard(p->device);
        PVB_CUDA(p, sync_all(p));
        const int rc = ensure_layout(p, pvb_processor::PLANAR, p->stream);
        if (rc != PVB_OK) return rc;
    }
    p->cursor_calls = uint64_t(calls);
    return PVB_OK;
}

size_t pvb_state_bytes(const pvb_processor *p) {
    return p ? size_t(2) * size_t(p->channels) * size_t(p->n) * sizeof(float) : 0;
}

int32_t pvb_get_state(pvb_processor *p, float *blob) {
    if (!p) return PVB_ERR_BAD_ARG;
    if (!blob) return fail(p, PVB_ERR_BAD_ARG, "NULL state blob");
    DeviceGuard guard(p->device);
    const size_t cn = size_t(p->channels) * size_t(p->n);
    if (cn == 0) return PVB_OK;
    std::vector<float> ring(2 * cn);
    PVB_CUDA(p, sync_all(p));
    {
        int rc = check_device_error(p);
        if (rc != PVB_OK) return rc;
        rc = ensure_layout(p, pvb_processor::PLANAR, p->stream);
        if (rc != PVB_OK) return rc;
    }
    PVB_CUDA(p, cudaMemcpy(ring.data(), p->d_hist, cn * sizeof(float), cudaMemcpyDeviceToHost));
    PVB_CUDA(p, cudaMemcpy(ring.data() + cn, p->d_acc, cn * sizeof(float), cudaMemcpyDeviceToHost));
    const int n = p->n, hop = p->hop;
    const int rb = int(p->ring_calls % uint64_t(p->overlaps)) * hop;
    for (int c = 0; c < p->channels; c++) {
        const float *h = ring.data() + size_t(c) * n, *a = ring.data() + cn + size_t(c) * n;
        float *ho = blob + size_t(c) * n, *ao = blob + cn + size_t(c) * n;
        for (int j = 0; j < n; j++) {
            ho[j] = h[(j + rb) & (n - 1)];
            ao[j] = (j < n - hop) ? a[(j + rb) & (n - 1)] : 0.0f;    // ola:134: the tail is zero
        }
    }
    return PVB_OK;
}

int32_t pvb_set_state(pvb_processor *p, const float *blob) {
    if (!p) return PVB_ERR_BAD_ARG;
    if (!blob) return fail(p, PVB_ERR_BAD_ARG, "NULL state blob");
    DeviceGuard guard(p->device);
    const size_t cn = size_t(p->channels) * size_t(p->n);
    if (cn == 0) return PVB_OK;
    std::vector<float> ring(2 * cn);
    const int n = p->n, hop = p->hop;
    const int rb = int(p->ring_calls % uint64_t(p->overlaps)) * hop;
    for (int c = 0; c < p->channels; c++) {
        float *h = ring.data() + size_t(c) * n, *a = ring.data() + cn + size_t(c) * n;
        const float *hi = blob + size_t(c) * n, *ai = blob + cn + size_t(c) * n;
        for (int j = 0; j < n; j++) {
            h[(j + rb) & (n - 1)] = hi[j];
            a[(j + rb) & (n - 1)] = (j < n - hop) ? ai[j] : 0.0f;
        }
    }
    PVB_CUDA(p, sync_all(p));
    p->layout = pvb_processor::PLANAR;      // whatever was there is replaced
    PVB_CUDA(p, cudaMemcpy(p->d_hist, ring.data(), cn * sizeof(float), cudaMemcpyHostToDevice));
    PVB_CUDA(p, cudaMemcpy(p->d_acc, ring.data() + cn, cn * sizeof(float), cudaMemcpyHostToDevice));
    return PVB_OK;
}

void *pvb_alloc_host(size_t bytes) {
    void *ptr = nullptr;
    if (cudaMallocHost(&ptr, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return ptr;
}

void pvb_free_host(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
