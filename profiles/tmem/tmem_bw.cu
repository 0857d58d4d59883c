// tmem_bw.cu -- throughput of an FFT-exchange-sized round trip through tensor memory: every warp stores 64
// registers per thread (8 KB per warp, tcgen05.st.32x32b.x64) and loads them back transposed with two
// tcgen05.ld.16x256b.x8 (the 2-bit lane <-> register swap of tmem_probe.cu), 14 warps per SM like the ring kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define R16(m, b) m(b+0) m(b+1) m(b+2) m(b+3) m(b+4) m(b+5) m(b+6) m(b+7) m(b+8) m(b+9) m(b+10) m(b+11) m(b+12) m(b+13) m(b+14) m(b+15)

__device__ __forceinline__ void st_x64(uint32_t ta, const uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,"
        "%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63,%64};"
        ::"r"(ta),
        "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]),
        "r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]),
        "r"(v[32]),"r"(v[33]),"r"(v[34]),"r"(v[35]),"r"(v[36]),"r"(v[37]),"r"(v[38]),"r"(v[39]),"r"(v[40]),"r"(v[41]),"r"(v[42]),"r"(v[43]),"r"(v[44]),"r"(v[45]),"r"(v[46]),"r"(v[47]),
        "r"(v[48]),"r"(v[49]),"r"(v[50]),"r"(v[51]),"r"(v[52]),"r"(v[53]),"r"(v[54]),"r"(v[55]),"r"(v[56]),"r"(v[57]),"r"(v[58]),"r"(v[59]),"r"(v[60]),"r"(v[61]),"r"(v[62]),"r"(v[63])
        : "memory");
}
__device__ __forceinline__ void ld_x8(uint32_t ta, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
          "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
        : "r"(ta) : "memory");
}

// mode 0: TMEM round trips; mode 1: the same bytes through shared memory (16 STS.128 + 16 LDS.128 per thread)
__global__ void __launch_bounds__(224, 2) bw(int iters, int mode, uint32_t *sink, long long *cyc) {
    __shared__ uint32_t tbase_s;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w: lanes 32 (w % 4) .., columns 64 (w / 4) ..
    const uint32_t ta = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16) + 64 * (warp >> 2);
    uint32_t v[64];
#pragma unroll
    for (int j = 0; j < 64; j++) v[j] = threadIdx.x * 64 + j;
    float4 *ex = reinterpret_cast<float4 *>(dyn) + warp * 544;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (mode == 0) {
            st_x64(ta, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            ld_x8(ta, v);
            ld_x8(ta + (16u << 16), v + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) ex[65 * (j & 7) + lane + 32 * (j >> 3)] = make_float4(__uint_as_float(v[4*j]), __uint_as_float(v[4*j+1]), __uint_as_float(v[4*j+2]), __uint_as_float(v[4*j+3]));
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const float4 q = ex[65 * ((lane >> 3) + 4 * (j >> 3)) + (lane & 7) + 8 * (j & 7)];
                v[4*j] = __float_as_uint(q.x); v[4*j+1] = __float_as_uint(q.y); v[4*j+2] = __float_as_uint(q.z); v[4*j+3] = __float_as_uint(q.w);
            }
            __syncwarp();
        }
        v[0] += it;
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 64; j++) acc ^= v[j];
    if (acc == 0x12345678) sink[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase_s), "n"(128) : "memory");
}

int main() {
    uint32_t *sink; long long *cyc;
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8 * 296);
    const int iters = 2000, smem = 7 * 544 * 16;
    cudaFuncSetAttribute(bw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            bw<<<296, 224, smem>>>(iters, mode, sink, cyc);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            // 14 warps per SM, iters trips each
            printf("%s: %s, %.3f ms, %lld cycles per CTA; per warp-trip on an SM: %.1f cycles (SM-level: %.1f cycles per 8 KB trip)\n",
                   mode ? "shared memory (16 STS.128 + 16 LDS.128)" : "tensor memory (st 32x32b.x64 + 2 ld 16x256b.x8)",
                   cudaGetErrorString(e), ms, h[0], (double)h[0] / iters, (double)h[0] / iters / 14.0);
        }
    }
    return 0;
}
