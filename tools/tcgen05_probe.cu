// tcgen05_probe.cu — groundwork for the fused student lm_head: one hand-written tcgen05.mma tile.
// D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T with bf16 operands staged in shared memory in the canonical K-major
// no-swizzle ("interleave") layout: 8 x 16-byte core matrices, LBO = distance between the two K chunks of one MMA,
// SBO = distance between 8-row groups.  Descriptor bit layouts as in cute/arch/mma_sm100_desc.hpp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tcgen05_probe tools/tcgen05_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);                 // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // version = 1 (sm_100)
    return d;                                              // layout type 0 = no swizzle, base offset 0
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
    __shared__ __align__(1024) unsigned char sA[M * K * 2];
    __shared__ __align__(1024) unsigned char sB[N * K * 2];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_sm;
    const int tid = threadIdx.x, warp = tid >> 5;
    // core matrix (i = row / 8, kc = k / 8) at ((kc * rows/8 + i) * 128) bytes, row r % 8 at + 16 * (r % 8)
    for (int idx = tid; idx < M * (K / 8); idx += 128) {
        const int r = idx / (K / 8), kc = idx % (K / 8);
        const uint4 v = *reinterpret_cast<const uint4*>(A + r * K + kc * 8);
        *reinterpret_cast<uint4*>(sA + ((kc * (M / 8) + r / 8) * 128 + (r % 8) * 16)) = v;
    }
    for (int idx = tid; idx < N * (K / 8); idx += 128) {
        const int r = idx / (K / 8), kc = idx % (K / 8);
        const uint4 v = *reinterpret_cast<const uint4*>(B + r * K + kc * 8);
        *reinterpret_cast<uint4*>(sB + ((kc * (N / 8) + r / 8) * 128 + (r % 8) * 16)) = v;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base_sm)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the MMA's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sm;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t lboA = (M / 8) * 128, lboB = (N / 8) * 128, sbo = 128;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * lboA, lboA, sbo);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * lboB, lboB, sbo);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    // everybody waits for the MMAs to retire
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32w..32w+31 = rows of D, 64 fp32 columns each
    uint32_t r[64];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
#pragma unroll
    for (int c = 0; c < 64; c += 8)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[c]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]), "=r"(r[c + 6]), "=r"(r[c + 7])
                     : "r"(taddr + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < N; ++c) D[tid * N + c] = __uint_as_float(r[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem) : "memory");
}

int main() {
    static __nv_bfloat16 hA[M * K], hB[N * K];
    static float hD[M * N], ref[M * N];
    srand(1);
    for (int i = 0; i < M * K; ++i) hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f);
    for (int i = 0; i < N * K; ++i) hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.f);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float a = 0.f;
            for (int k = 0; k < K; ++k) a += __bfloat162float(hA[m * K + k]) * __bfloat162float(hB[n * K + k]);
            ref[m * N + n] = a;
        }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, sizeof(hD));
    probe<<<1, 128>>>(dA, dB, dD);
    printf("launch: %s\n", cudaGetErrorString(cudaGetLastError()));
    printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    for (int i = 0; i < M * N; ++i) {
        const double e = fabs((double)hD[i] - ref[i]);
        if (!(e <= 1e-3)) ++bad;
        if (e > maxerr || e != e) maxerr = e;
    }
    printf("D[0][0..3] = %g %g %g %g   ref %g %g %g %g\n", hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
    printf("mismatches %d of %d, max abs err %g\n", bad, M * N, maxerr);
    return bad != 0;
}
