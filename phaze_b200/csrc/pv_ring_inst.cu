// pv_ring_inst.cu — instantiates the ring-order kernel (pv_kernel_ring.cuh) for ONE frame size:
// compiled five times with -DPVB_RING_INST_N=256 / 512 / 1024 / 2048 / 4096 (see the Makefile).
#include "pv_kernel_ring.cuh"
#include "pv_ring_launch.h"

#include <initializer_list>

#ifndef PVB_RING_INST_N
#error "compile with -DPVB_RING_INST_N=<frame size>"
#endif

namespace pvb {
namespace {

// NBLK counts role units of RingGeoT<N>::UNIT samples (64 at frame 256, else 128)
template <int N, bool PCH, bool MULTI, int DEEP, int... NBLKS>
cudaError_t launch_t(const RingParams &rp, const RingLaunch &l) {
    using G = RingGeoT<N, PCH>;
    constexpr int CAP = DEEP ? G::DEEP_PAIRS : MULTI ? G::MULTI_PAIRS : G::MAX_PAIRS;   // what the kernel's launch bounds (and two CTAs per SM) allow
    int ppc = l.ppc;
    if (ppc < G::MIN_PAIRS || ppc > CAP) ppc = CAP;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((l.pairs + ppc - 1) / ppc);
    cfg.blockDim = dim3(ppc * G::TP);
    cfg.dynamicSmemBytes = G::TAB_BYTES + size_t(ppc) * (G::PAIR_BYTES + (DEEP ? G::DEEP_BYTES : 0)) + size_t(l.pad_kb) * 1024;
    cfg.stream = l.stream;
    // programmatic dependent launch: CTAs of this launch may become resident and stage their tables while
    // the previous kernel on the stream drains; the kernel itself orders its accesses
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = l.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int nblk = rp.hop / G::UNIT;
    cudaError_t e = cudaErrorInvalidValue;                      // no instance for this hop
    (void)std::initializer_list<int>{
        (nblk == NBLKS ? (e = cudaLaunchKernelEx(&cfg, pv_process_ring_kernel<N, NBLKS, PCH, MULTI, DEEP>, rp), 0) : 0)...};
    return e;
}

template <int N, bool PCH, bool MULTI, int DEEP, int... NBLKS>
cudaError_t configure_t() {
    cudaError_t e = cudaSuccess;
    (void)std::initializer_list<int>{
        (e == cudaSuccess ? (e = cudaFuncSetAttribute(pv_process_ring_kernel<N, NBLKS, PCH, MULTI, DEEP>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), 0)
                          : 0)...};
    return e;
}

// (inside templates so that a frame size without the DEEP instance does not instantiate it)
template <int N, bool DEEPOK, int... NBLKS>
cudaError_t launch_deep_t(const RingParams &rp, const RingLaunch &l) {
    if constexpr (DEEPOK) {
        if (l.deep == 2) return l.pch ? launch_t<N, true, false, 2, NBLKS...>(rp, l) : launch_t<N, false, false, 2, NBLKS...>(rp, l);
        return l.pch ? launch_t<N, true, false, 1, NBLKS...>(rp, l) : launch_t<N, false, false, 1, NBLKS...>(rp, l);
    } else {
        return cudaErrorInvalidValue;
    }
}
template <int N, bool DEEPOK, int... NBLKS>
cudaError_t configure_deep_t() {
    if constexpr (DEEPOK) {
        cudaError_t e = configure_t<N, false, false, 1, NBLKS...>();
        if (e == cudaSuccess) e = configure_t<N, true, false, 1, NBLKS...>();
        if (e == cudaSuccess) e = configure_t<N, false, false, 2, NBLKS...>();
        if (e == cudaSuccess) e = configure_t<N, true, false, 2, NBLKS...>();
        return e;
    } else {
        return cudaSuccess;
    }
}

}  // namespace

// instances per (frame, hop): scalar pitch factor, per-channel pitch factors, several calls per launch, and
// the DEEP instances (pitch factors down to 0.5), scalar and per channel
#define PVB_RING_DEFINE(N, DEEPOK, ...)                                                             \
    cudaError_t ring_launch_##N(const RingParams &rp, const RingLaunch &l) {                        \
        if (l.deep)                                                                                 \
            return l.multi ? cudaErrorInvalidValue : launch_deep_t<N, DEEPOK, __VA_ARGS__>(rp, l);  \
        if (l.pch) return l.multi ? cudaErrorInvalidValue : launch_t<N, true, false, 0, __VA_ARGS__>(rp, l); \
        return l.multi ? launch_t<N, false, true, 0, __VA_ARGS__>(rp, l)                        \
                       : launch_t<N, false, false, 0, __VA_ARGS__>(rp, l);                      \
    }                                                                                               \
    cudaError_t ring_configure_##N() {                                                              \
        cudaError_t e = configure_t<N, false, false, 0, __VA_ARGS__>();                         \
        if (e == cudaSuccess) e = configure_t<N, true, false, 0, __VA_ARGS__>();                \
        if (e == cudaSuccess) e = configure_t<N, false, true, 0, __VA_ARGS__>();                \
        if (e == cudaSuccess) e = configure_deep_t<N, DEEPOK, __VA_ARGS__>();                       \
        return e;                                                                                   \
    }

#if PVB_RING_INST_N == 256
PVB_RING_DEFINE(256, true, 1, 2)
#elif PVB_RING_INST_N == 512
PVB_RING_DEFINE(512, true, 1, 2)
#elif PVB_RING_INST_N == 1024
PVB_RING_DEFINE(1024, true, 1, 2, 4)
#elif PVB_RING_INST_N == 2048
PVB_RING_DEFINE(2048, true, 1, 2, 4, 8)
#elif PVB_RING_INST_N == 4096
PVB_RING_DEFINE(4096, true, 2, 4, 8, 16)
#else
#error "PVB_RING_INST_N must be 256, 512, 1024, 2048 or 4096"
#endif

}  // namespace pvb
