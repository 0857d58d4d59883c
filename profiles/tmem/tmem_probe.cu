// tmem_probe.cu -- which (thread, register) of a warp ends up where when data goes registers -> tensor memory
// -> registers with different tcgen05.st / tcgen05.ld shapes (sm_100a).  Question behind it: can the FFT
// exchanges of pv_kernel_ring.cuh (register index <-> lane index transposes, 512 of the 1880 load/store-pipe
// cycles per channel pair) go through TMEM, whose data path is separate from the shared-memory pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NCOLS 64

__device__ __forceinline__ void st32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ld32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld16x64b_x4(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld16x128b_x2(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld16x256b_x1(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st16x256b_x2(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void st16x64b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x64b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// out[test][warp][lane][8]
__global__ void probe(uint32_t *out, long long *cycles) {
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)), "n"(NCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    // every warp works on its own 32 lanes (lane field = bits 31:16): warp w -> lanes 32 (w % 4) ..
    const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t v[8], r[8];
    auto dump = [&](int test) {
        for (int j = 0; j < 8; j++) out[((test * 4 + warp) * 32 + lane) * 8 + j] = r[j];
    };
    for (int j = 0; j < 8; j++) v[j] = (uint32_t)(threadIdx.x << 8) | j;
    // test 0: 32x32b store, 32x32b load (identity expected)
    st32x32b_x8(taddr, v); wait_st();
    for (int j = 0; j < 8; j++) r[j] = 0xdeadbeef;
    ld32x32b_x8(taddr, r); wait_ld(); dump(0);
    // tests 1..4: 32x32b store, loads of other shapes from lane offset 0 and 16
    for (int half = 0; half < 2; half++) {
        const uint32_t ta = taddr + ((uint32_t)(16 * half) << 16);
        for (int j = 0; j < 8; j++) r[j] = 0xdeadbeef;
        ld16x64b_x4(ta, r); wait_ld(); dump(1 + 4 * half);
        for (int j = 0; j < 8; j++) r[j] = 0xdeadbeef;
        ld16x128b_x2(ta, r); wait_ld(); dump(2 + 4 * half);
        for (int j = 0; j < 8; j++) r[j] = 0xdeadbeef;
        ld16x256b_x1(ta, r); wait_ld(); dump(3 + 4 * half);
        for (int j = 0; j < 8; j++) r[j] = 0xdeadbeef;
        ld16x256b_x2(ta, r); wait_ld(); dump(4 + 4 * half);
    }
    // tests 9, 10: stores of other shapes (to lanes 0..15 and 16..31 of the warp), 32x32b load
    for (int j = 0; j < 8; j++) r[j] = 0;
    st32x32b_x8(taddr, r); wait_st();                       // clear
    st16x256b_x2(taddr, v); wait_st();
    ld32x32b_x8(taddr, r); wait_ld(); dump(9);
    for (int j = 0; j < 8; j++) r[j] = 0;
    st32x32b_x8(taddr, r); wait_st();
    st16x64b_x8(taddr, v); wait_st();
    ld32x32b_x8(taddr, r); wait_ld(); dump(10);
    // timing: 64 round trips of 16 registers (st 32x32b.x8 twice, ld 16x256b.x2 twice) per warp
    __syncthreads();
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < 64; it++) {
        st32x32b_x8(taddr, v); st32x32b_x8(taddr + 8, v); wait_st();
        ld16x256b_x2(taddr, r); acc += r[0]; ld16x256b_x2(taddr + 8, r); wait_ld(); acc += r[1];
        v[0] += acc;
    }
    long long t1 = clock64();
    if (lane == 0) cycles[warp] = t1 - t0;
    if (acc == 0x12345) out[0] = acc;
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(NCOLS) : "memory");
}

int main() {
    const int ntest = 11, nwarp = 4;
    uint32_t *d; long long *dc;
    cudaMalloc(&d, ntest * nwarp * 32 * 8 * 4); cudaMemset(d, 0xff, ntest * nwarp * 32 * 8 * 4);
    cudaMalloc(&dc, 4 * 8);
    probe<<<1, 128>>>(d, dc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    static uint32_t h[11 * 4 * 32 * 8]; long long hc[4];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
    const char *names[ntest] = {"st32x32b.x8 -> ld32x32b.x8", "ld16x64b.x4 @lane+0", "ld16x128b.x2 @lane+0", "ld16x256b.x1 @lane+0", "ld16x256b.x2 @lane+0",
                                "ld16x64b.x4 @lane+16", "ld16x128b.x2 @lane+16", "ld16x256b.x1 @lane+16", "ld16x256b.x2 @lane+16",
                                "st16x256b.x2 -> ld32x32b.x8", "st16x64b.x8 -> ld32x32b.x8"};
    for (int t = 0; t < ntest; t++) {
        printf("== test %d: %s (warp 1; entries: src_thread.reg, -- = untouched)\n", t, names[t]);
        for (int l = 0; l < 32; l++) {
            printf("  T%02d:", l);
            for (int j = 0; j < 8; j++) {
                uint32_t x = h[((t * 4 + 1) * 32 + l) * 8 + j];
                if (x == 0xdeadbeef || x == 0) printf("   -- ");
                else printf(" %3u.%u", (x >> 8) - 32, x & 0xff);
            }
            printf("\n");
        }
    }
    printf("timing: 64 x (16 regs st 32x32b + 16 regs ld 16x256b) per warp, 4 warps concurrently: %lld %lld %lld %lld cycles\n", hc[0], hc[1], hc[2], hc[3]);
    return 0;
}
