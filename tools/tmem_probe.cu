// tmem_probe.cu — can tensor memory serve as a per-thread scratch for a non-MMA kernel on sm_100a?
// 16 warps per CTA; warp w owns TMEM lanes 32*(w%4).. and columns 128*(w/4)..  Checks st -> ld round trips
// (32x32b.x8) and times ld / st / LDS.128 loops per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(512, 1) probe(uint32_t* out, long long* cycles, int reps) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t tmem_base_sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_sm)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_sm;
    const uint32_t my = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * (warp >> 2));

    // ---- correctness: every thread writes 128 words, reads them back
    uint32_t bad = 0;
    for (int j = 0; j < 16; ++j) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = (blockIdx.x << 24) ^ (tid << 12) ^ (j * 8 + i) ^ 0x5a000000u;
        tm_st8(my + 8 * j, r);
    }
    tm_wait_st();
    for (int j = 0; j < 16; ++j) {
        uint32_t r[8];
        tm_ld8(my + 8 * j, r);
        tm_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) bad += (r[i] != ((blockIdx.x << 24) ^ (tid << 12) ^ (j * 8 + i) ^ 0x5a000000u));
    }
    // partial-warp divergence before a .sync.aligned op, then reconverge
    if (lane < 7) bad += (lane == 99);
    __syncwarp();
    {
        uint32_t r[8];
        tm_ld8(my + 8, r);
        tm_wait_ld();
        bad += (r[3] != ((blockIdx.x << 24) ^ (tid << 12) ^ (8 + 3) ^ 0x5a000000u));
    }
    out[blockIdx.x * 512 + tid] = bad;

    // ---- timing: ld loop
    __syncthreads();
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < reps; ++it) {
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            uint32_t r[8];
            tm_ld8(my + 8 * j, r);
            tm_wait_ld();
            acc += r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
        }
    }
    __syncthreads();
    long long t1 = clock64();
    // st loop
    for (int it = 0; it < reps; ++it) {
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = acc + i + it;
            tm_st8(my + 8 * j, r);
        }
        tm_wait_st();
    }
    __syncthreads();
    long long t2 = clock64();
    // LDS.128 loop of the same byte volume (2 x 16 B per 8 words)
    uint4* sm4 = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < 16 * 512; i += 512) sm4[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    long long t3 = clock64();
    for (int it = 0; it < reps; ++it) {
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            uint4 v;  // 16 distinct vectors per thread, each read twice
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"((uint32_t)__cvta_generic_to_shared(sm4 + (j & 15) * 512 + tid)));
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    __syncthreads();
    long long t4 = clock64();
    // ld with 2 loads in flight before the wait
    for (int it = 0; it < reps; ++it) {
#pragma unroll 2
        for (int j = 0; j < 16; j += 2) {
            uint32_t r[8], s[8];
            tm_ld8(my + 8 * j, r);
            tm_ld8(my + 8 * j + 8, s);
            tm_wait_ld();
            acc += r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7] ^ s[0] ^ s[1] ^ s[2] ^ s[3] ^ s[4] ^ s[5] ^ s[6] ^ s[7];
        }
    }
    __syncthreads();
    long long t5 = clock64();
    if (tid == 0) {
        cycles[blockIdx.x * 4 + 0] = t1 - t0;
        cycles[blockIdx.x * 4 + 1] = t2 - t1;
        cycles[blockIdx.x * 4 + 2] = t4 - t3;
        cycles[blockIdx.x * 4 + 3] = t5 - t4;
    }
    out[blockIdx.x * 512 + tid] += (acc == 0x12345678u);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
}

int main() {
    const int grid = 148, reps = 200;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, grid * 512 * 4);
    cudaMalloc(&cyc, grid * 4 * 8);
    cudaMemset(out, 0xff, grid * 512 * 4);
    printf("attr: %s\n", cudaGetErrorString(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 16)));
    probe<<<grid, 512, 16 * 512 * 16>>>(out, cyc, reps);
    printf("launch: %s\n", cudaGetErrorString(cudaGetLastError()));
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    static uint32_t h[148 * 512];
    long long hc[148 * 4];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < grid * 512; ++i) bad += h[i];
    printf("mismatches: %ld\n", bad);
    const double bytes = (double)reps * 16 * 8 * 4 * 512;      // per CTA per loop
    printf("tmem ld  (wait each)  : %.1f B/clk/SM\n", bytes / hc[0]);
    printf("tmem st               : %.1f B/clk/SM\n", bytes / hc[1]);
    printf("lds.128               : %.1f B/clk/SM\n", bytes / hc[2]);
    printf("tmem ld  (2 in flight): %.1f B/clk/SM\n", bytes / hc[3]);
    return bad != 0;
}
