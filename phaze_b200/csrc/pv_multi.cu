// pv_multi.cu — one logical processor sharded over several GPUs of a node, inside ONE process, behind
// the C ABI (pvb_multi_* in include/phaze_b200.h).
//
// Channels are independent on this path (phase-vocoder.js:49-53; the only shared datum, timeCursor,
// advances identically everywhere), so shard i owns a contiguous, pair-aligned block of channels with its
// rings in its own HBM, and a call touches no other device's memory except for the exchange of the
// caller's blocks:
//   * host entry points: every shard copies ITS slab of the host block to its device and back (the
//     scatter / gather is the set of host<->device copies; nothing is staged through a root GPU);
//   * root entry points (the block lives in the HBM of device_ids[0]): slabs travel to / from the other
//     devices with cudaMemcpyPeerAsync -- copy engines over NVLink, no SM of any GPU is spent on the
//     exchange (the multi-process twin, phaze_b200/sharded.py, uses ncclSend / ncclRecv).
// Built only from the public single-handle entry points: sharding cannot change a bit of the result.
#include "../../include/phaze_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

struct pvb_multi {
    struct Shard {
        pvb_processor *h = nullptr;
        int device = 0, lo = 0, hi = 0;         // channels [lo, hi)
        cudaStream_t stream = nullptr;
        float *d_in = nullptr, *d_out = nullptr;
        size_t staging_floats = 0;
    };
    std::vector<Shard> shards;
    int channels = 0, n = 0, hop = 0;
    char err[256] = "";
};

namespace {

thread_local char g_multi_err[256] = "";

int mfail(pvb_multi *m, int code, const char *fmt, ...) {
    char *dst = m ? m->err : g_multi_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    return code;
}

#define PVM_CUDA(m, call)                                                               \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess)                                                          \
            return mfail((m), PVB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

struct DevGuard {
    int prev = -1;
    DevGuard() { cudaGetDevice(&prev); }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int ensure_staging(pvb_multi *m, pvb_multi::Shard &s, size_t floats) {
    if (floats <= s.staging_floats) return PVB_OK;
    PVM_CUDA(m, cudaStreamSynchronize(s.stream));
    cudaFree(s.d_in);
    cudaFree(s.d_out);
    s.d_in = s.d_out = nullptr;
    s.staging_floats = 0;
    if (cudaMalloc(&s.d_in, floats * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&s.d_out, floats * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return mfail(m, PVB_ERR_NOMEM, "cudaMalloc of the shard staging buffers failed");
    }
    s.staging_floats = floats;
    return PVB_OK;
}

int shard_error(pvb_multi *m, const pvb_multi::Shard &s, int rc) {
    return mfail(m, rc, "device %d: %s", s.device, pvb_last_error(s.h));
}

// common body: `root_device` < 0: in / out are host buffers; else they live on that device
int run(pvb_multi *m, const float *in, float *out, int num_calls, float pf, int root_device) {
    if (!m) return PVB_ERR_BAD_ARG;
    if (!out || num_calls < 0) return mfail(m, PVB_ERR_BAD_ARG, "pvb_multi_process: bad argument");
    DevGuard guard;
    const size_t hop = size_t(m->hop), C = size_t(m->channels);
    // submit everything on every device, then wait: the devices work concurrently
    for (pvb_multi::Shard &s : m->shards) {
        const size_t cl = size_t(s.hi - s.lo);
        PVM_CUDA(m, cudaSetDevice(s.device));
        if (cl == 0) {           // keeps timeCursor in step
            const int rc = pvb_process_many_device(s.h, nullptr, out, num_calls, pf, s.stream);   // (out is not touched)
            if (rc != PVB_OK) return shard_error(m, s, rc);
            continue;
        }
        const bool direct = root_device == s.device;      // the root's own slab: no copy at all for K == 1
        float *din = nullptr, *dout = nullptr;
        if (direct && num_calls == 1) {
            din = in ? const_cast<float *>(in) + size_t(s.lo) * hop : nullptr;
            dout = out + size_t(s.lo) * hop;
        } else {
            const int rc = ensure_staging(m, s, cl * hop * size_t(num_calls));
            if (rc != PVB_OK) return rc;
            din = in ? s.d_in : nullptr;
            dout = s.d_out;
            for (int k = 0; in && k < num_calls; k++) {
                const float *src = in + (size_t(k) * C + size_t(s.lo)) * hop;
                float *dst = s.d_in + size_t(k) * cl * hop;
                if (root_device < 0)
                    PVM_CUDA(m, cudaMemcpyAsync(dst, src, cl * hop * sizeof(float), cudaMemcpyHostToDevice, s.stream));
                else
                    PVM_CUDA(m, cudaMemcpyPeerAsync(dst, s.device, src, root_device, cl * hop * sizeof(float), s.stream));
            }
        }
        const int rc = pvb_process_many_device(s.h, din, dout, num_calls, pf, s.stream);
        if (rc != PVB_OK) return shard_error(m, s, rc);
        if (dout == s.d_out) {
            for (int k = 0; k < num_calls; k++) {
                float *dst = out + (size_t(k) * C + size_t(s.lo)) * hop;
                const float *src = s.d_out + size_t(k) * cl * hop;
                if (root_device < 0)
                    PVM_CUDA(m, cudaMemcpyAsync(dst, src, cl * hop * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
                else
                    PVM_CUDA(m, cudaMemcpyPeerAsync(dst, root_device, src, s.device, cl * hop * sizeof(float), s.stream));
            }
        }
    }
    for (pvb_multi::Shard &s : m->shards) {
        PVM_CUDA(m, cudaSetDevice(s.device));
        PVM_CUDA(m, cudaStreamSynchronize(s.stream));
        const int rc = pvb_sync(s.h);                     // also surfaces the shard's sticky device error
        if (rc != PVB_OK) return shard_error(m, s, rc);
    }
    return PVB_OK;
}

}  // namespace

extern "C" {

const char *pvb_multi_last_error(const pvb_multi *m) { return m ? m->err : g_multi_err; }

int32_t pvb_multi_create(const pvb_config *cfg, const int32_t *device_ids, int32_t num_devices, pvb_multi **out) {
    if (!cfg || !out || !device_ids || num_devices < 1)
        return mfail(nullptr, PVB_ERR_BAD_ARG, "pvb_multi_create: bad argument");
    *out = nullptr;
    if (cfg->num_channels < 0) return mfail(nullptr, PVB_ERR_BAD_ARG, "negative channel count");
    pvb_multi *m = new (std::nothrow) pvb_multi();
    if (!m) return mfail(nullptr, PVB_ERR_NOMEM, "host allocation failed");
    DevGuard guard;
    m->channels = cfg->num_channels;
    // contiguous blocks with boundaries on even channels: the kernels process channels in pairs, and the
    // pairing must not depend on the number of shards (same rule as phaze_b200/sharded.py shard_bounds)
    const long long pairs = (m->channels + 1) / 2;
    auto bail = [&](int code) { strncpy(g_multi_err, m->err, 255); pvb_multi_destroy(m); return code; };
    for (int i = 0; i < num_devices; i++) {
        pvb_multi::Shard s;
        s.device = device_ids[i];
        s.lo = int(std::min<long long>(2 * ((pairs * i) / num_devices), m->channels));
        s.hi = int(std::min<long long>(2 * ((pairs * (i + 1)) / num_devices), m->channels));
        pvb_config c = *cfg;
        c.num_channels = s.hi - s.lo;
        c.device = s.device;
        const int rc = pvb_create(&c, &s.h);
        if (rc != PVB_OK) {
            mfail(m, rc, "device %d: %s", s.device, pvb_last_error(nullptr));
            return bail(rc);
        }
        if (cudaSetDevice(s.device) != cudaSuccess ||
            cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) {
            mfail(m, PVB_ERR_CUDA, "device %d: stream creation failed: %s", s.device, cudaGetErrorString(cudaGetLastError()));
            m->shards.push_back(s);
            return bail(PVB_ERR_CUDA);
        }
        m->shards.push_back(s);
    }
    m->n = pvb_frame_size(m->shards[0].h);
    m->hop = pvb_hop_size(m->shards[0].h);
    // peer access between the root (device_ids[0]) and every other device, both directions, for the
    // root entry points; without it cudaMemcpyPeerAsync still works (staged through the host)
    for (size_t i = 1; i < m->shards.size(); i++) {
        const int a = m->shards[0].device, b = m->shards[i].device;
        if (a == b) continue;
        int ok = 0;
        if (cudaDeviceCanAccessPeer(&ok, a, b) == cudaSuccess && ok) {
            cudaSetDevice(a);
            if (cudaDeviceEnablePeerAccess(b, 0) != cudaSuccess) cudaGetLastError();    // already enabled is fine
            cudaSetDevice(b);
            if (cudaDeviceEnablePeerAccess(a, 0) != cudaSuccess) cudaGetLastError();
        }
    }
    *out = m;
    return PVB_OK;
}

void pvb_multi_destroy(pvb_multi *m) {
    if (!m) return;
    DevGuard guard;
    for (pvb_multi::Shard &s : m->shards) {
        cudaSetDevice(s.device);
        if (s.stream) cudaStreamSynchronize(s.stream);
        pvb_destroy(s.h);
        cudaFree(s.d_in);
        cudaFree(s.d_out);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete m;
}

int32_t pvb_multi_num_devices(const pvb_multi *m) { return m ? int32_t(m->shards.size()) : PVB_ERR_BAD_ARG; }
int32_t pvb_multi_num_channels(const pvb_multi *m) { return m ? m->channels : PVB_ERR_BAD_ARG; }

pvb_processor *pvb_multi_shard(pvb_multi *m, int32_t index, int32_t *first_channel, int32_t *num_channels) {
    if (!m || index < 0 || index >= int32_t(m->shards.size())) return nullptr;
    if (first_channel) *first_channel = m->shards[index].lo;
    if (num_channels) *num_channels = m->shards[index].hi - m->shards[index].lo;
    return m->shards[index].h;
}

int32_t pvb_multi_process(pvb_multi *m, const float *in, float *out, float pitch_factor) {
    return run(m, in, out, 1, pitch_factor, -1);
}

int32_t pvb_multi_process_many(pvb_multi *m, const float *in, float *out, int32_t num_calls, float pitch_factor) {
    return run(m, in, out, num_calls, pitch_factor, -1);
}

int32_t pvb_multi_process_root(pvb_multi *m, const float *in_dev, float *out_dev, int32_t num_calls,
                               float pitch_factor) {
    if (!m) return PVB_ERR_BAD_ARG;
    return run(m, in_dev, out_dev, num_calls, pitch_factor, m->shards[0].device);
}

int32_t pvb_multi_set_option(pvb_multi *m, int32_t option, int64_t value) {
    if (!m) return PVB_ERR_BAD_ARG;
    for (pvb_multi::Shard &s : m->shards) {
        const int rc = pvb_set_option(s.h, option, value);
        if (rc != PVB_OK) return shard_error(m, s, rc);
    }
    return PVB_OK;
}

}  // extern "C"
