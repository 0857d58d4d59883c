// pv_ring_launch.h — host-callable launchers of the ring-order kernel, one translation unit per frame
// size (pv_ring_inst.cu compiled with -DPVB_RING_INST_N=<frame>), so that the 30 kernel instances
// (frame x hop x scalar / per-channel pitch factor) compile in parallel.
#pragma once

#include <cuda_runtime.h>

namespace pvb {

struct RingParams;

struct RingLaunch {
    int pairs;            // channel pairs of the launch
    int ppc;              // pairs per CTA (0: the frame size's maximum)
    int pad_kb;           // extra dynamic shared memory per CTA (occupancy experiments)
    bool pdl;             // programmatic dependent launch allowed
    bool pch;             // per-channel pitch factors (RingParams::pf_ch)
    bool multi;           // several process() calls per launch (RingParams::num_hops; scalar pitch factor only)
    int deep;             // 0; 1: pitch factors down to 0.5; 2: down to 0.33 -- the DEEP instances (one call per launch)
    cudaStream_t stream;
};

// launch the instance for rp.hop; cudaErrorInvalidValue when the frame size has none for this hop
cudaError_t ring_launch_256(const RingParams &rp, const RingLaunch &l);
cudaError_t ring_launch_512(const RingParams &rp, const RingLaunch &l);
cudaError_t ring_launch_1024(const RingParams &rp, const RingLaunch &l);
cudaError_t ring_launch_2048(const RingParams &rp, const RingLaunch &l);
cudaError_t ring_launch_4096(const RingParams &rp, const RingLaunch &l);
// raise the dynamic shared memory limit of every instance (once per device)
cudaError_t ring_configure_256();
cudaError_t ring_configure_512();
cudaError_t ring_configure_1024();
cudaError_t ring_configure_2048();
cudaError_t ring_configure_4096();

}  // namespace pvb
