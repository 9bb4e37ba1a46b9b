// tcgen05_gemm2_probe.cu — the CTA-PAIR form of tools/tcgen05_gemm_probe.cu: two CTAs of one TPC (cluster 2 x 1 x 1) compute one
// 256 x 256 output tile with tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16), issued by the even CTA only.  Each CTA loads
// its 128 rows of A and HALF of the B tile (128 of the 256 weight rows) — 32 KB per stage instead of 48 KB, six stages — with
// cp.async.bulk.tensor...cta_group::2, whose completion bytes land on the LEADER's full barrier; tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to both CTAs; each CTA drains its own 128 accumulator lanes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tcgen05_gemm2_probe tools/tcgen05_gemm2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int BM = 128, BN = 256, BNH = BN / 2, BK = 64, STAGES = 6;      // BM: rows per CTA (256 per pair)
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNH * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {      // the same shared-memory offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void commit_pair(uint32_t bar) {      // arrives on `bar` (same offset) in BOTH CTAs once the MMAs issued so far retire
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major operand tile with 128-byte swizzle: rows are 128 B, 8-row groups 1024 B apart (SBO); LBO unused (1)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;           // version (sm_100)
    d |= (uint64_t)2 << 61;           // SWIZZLE_128B
    return d;
}

constexpr int C_SLAB_BYTES = BM * 64 * 2;

// tile order: groups of `group_m` row-tiles walk the column tiles together, so one wave of CTAs touches
// ~group_m A tiles and ~148/group_m B tiles (both L2-resident) instead of every A tile
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int group_m, int& tile_m, int& tile_n) {
    const int per_group = group_m * tiles_n;
    const int g = t / per_group, first_m = g * group_m;
    const int gm = min(group_m, tiles_m - first_m);
    const int in_g = t - g * per_group;
    tile_m = first_m + in_g % gm;
    tile_n = in_g / gm;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapC,
            int M, int N, int K, int group_m, int no_store) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 4];
    __shared__ uint32_t tmem_base_sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + STAGES * 8;
    const uint32_t bar_acc_full = bar_full + 2 * STAGES * 8, bar_acc_empty = bar_acc_full + 2 * 8;
    const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + BN - 1) / BN, tiles = tiles_m * tiles_n;
    const int kblocks = K / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_acc_full + a * 8, 1); mbar_init(bar_acc_empty + a * 8, 8); }      // 4 warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_sm)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sm;

    if (warp == 0) {
        if (lane == 0) {          // ---------------- TMA producer (both CTAs) ----------------
            uint32_t it = 0;
            for (int t = pair; t < tiles; t += npairs) {
                int tile_m, tile_n;
                tile_coords(t, tiles_m, tiles_n, group_m, tile_m, tile_n);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bar_empty + s * 8, ph ^ 1u);
                    if (rank == 0) mbar_expect_tx(bar_full + s * 8, 2 * STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
                    const uint32_t leader_full = mapa_rank(bar_full + s * 8, 0);
                    const uint32_t dstA = smem_u32(smem + s * STAGE_BYTES), dstB = dstA + A_BYTES;
                    tma_load_2d_2sm(dstA, &mapA, kb * BK, tile_m * 2 * BM + (int)rank * BM, leader_full);
                    tma_load_2d_2sm(dstB, &mapB, kb * BK, tile_n * BN + (int)rank * BNH, leader_full);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {          // ---------------- MMA issuer (leader CTA only) ----------------
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
            uint32_t it = 0, i = 0;
            for (int t = pair; t < tiles; t += npairs, ++i) {
                const uint32_t as = i & 1, aph = (i >> 1) & 1;
                mbar_wait(bar_acc_empty + as * 8, aph ^ 1u);          // both CTAs' epilogues have drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + as * BN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bar_full + s * 8, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = make_desc_sw128(a0 + k * 32), db = make_desc_sw128(b0 + k * 32);
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                            :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                    }
                    commit_pair(bar_empty + s * 8);
                }
                commit_pair(bar_acc_full + as * 8);
            }
        }
    } else {                      // ---------------- epilogue (both CTAs): warps 2..5 <-> TMEM lane quarters ----------------
        const int q = warp & 3;
        const int r_in = q * 32 + lane;
        unsigned char* cstage = smem + STAGES * STAGE_BYTES;
        uint32_t i = 0, slab = 0;
        for (int t = pair; t < tiles; t += npairs, ++i) {
            int tile_m, tile_n;
            tile_coords(t, tiles_m, tiles_n, group_m, tile_m, tile_n);
            const uint32_t as = i & 1, aph = (i >> 1) & 1;
            mbar_wait(bar_acc_full + as * 8, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < BN; c += 64, ++slab) {
                uint32_t r[64];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + as * BN + (uint32_t)c;
                tmem_ld32(taddr, r);
                tmem_ld32(taddr + 32, r + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                unsigned char* buf = cstage + (slab & 1) * C_SLAB_BYTES;
                if (tid == 64) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[8 * j + 2 * u]), __uint_as_float(r[8 * j + 2 * u + 1]));
                        w[u] = *reinterpret_cast<uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(buf + r_in * 128 + ((j ^ (r_in & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tid == 64) {
                    if (!no_store)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                     :: "l"(&mapC), "r"(tile_n * BN + c), "r"(tile_m * 2 * BM + (int)rank * BM), "r"(smem_u32(buf)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(mapa_rank(bar_acc_empty + as * 8, 0)) : "memory");
        }
        if (tid == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiled enc, void* ptr, int rows, int cols, int box_rows, CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 8192, N = argc > 2 ? atoi(argv[2]) : 8192, K = argc > 3 ? atoi(argv[3]) : 896;
    if (N % 8 || K % BK) { printf("N %% 8, K %% 64 must be 0\n"); return 1; }
    std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
    srand(7);
    for (auto& v : hA) v = __float2bfloat16((rand() % 33 - 16) / 16.f);
    for (auto& v : hB) v = __float2bfloat16((rand() % 33 - 16) / 32.f);
    __nv_bfloat16 *dA, *dB, *dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, (size_t)M * N * 2);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, (size_t)M * N * 2);
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const CUtensorMap mapA = make_map(enc, dA, M, K, BM), mapB = make_map(enc, dB, N, K, BNH);
    const CUtensorMap mapC = make_map(enc, dC, M, N, BM, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    const int group_m = argc > 5 ? atoi(argv[5]) : 8, no_store = argc > 6 ? atoi(argv[6]) : 0;
    const int smem = STAGES * STAGE_BYTES + 2 * C_SLAB_BYTES + 1024;
    cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
    if (argc > 4 && atoi(argv[4]) > 0) sms = atoi(argv[4]);
    const int pairs = sms / 2 < tiles ? sms / 2 : tiles;
    dim3 grid(2 * pairs);
    gemm_kernel<<<grid, THREADS, smem>>>(mapA, mapB, mapC, M, N, K, group_m, no_store);
    printf("launch: %s\n", cudaGetErrorString(cudaGetLastError()));
    printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    // spot check 4096 random entries against the CPU product
    std::vector<__nv_bfloat16> hC((size_t)M * N);
    cudaMemcpy(hC.data(), dC, hC.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    double maxerr = 0;
    for (int t = 0; t < 4096; ++t) {
        const int m = rand() % M, n = rand() % N;
        double a = 0;
        for (int k = 0; k < K; ++k) a += (double)__bfloat162float(hA[(size_t)m * K + k]) * __bfloat162float(hB[(size_t)n * K + k]);
        const double got = __bfloat162float(hC[(size_t)m * N + n]);
        const double e = fabs(got - a);
        if (!(e <= 0.01 * fabs(a) + 0.05)) { if (bad < 5) printf("  C[%d][%d] = %g, ref %g\n", m, n, got, a); ++bad; }
        if (e > maxerr || e != e) maxerr = e;
    }
    printf("spot check: %d bad of 4096, max abs err %g\n", bad, maxerr);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) gemm_kernel<<<grid, THREADS, smem>>>(mapA, mapB, mapC, M, N, K, group_m, no_store);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) gemm_kernel<<<grid, THREADS, smem>>>(mapA, mapB, mapC, M, N, K, group_m, no_store);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    printf("M=%d N=%d K=%d ctas=%d group_m=%d%s: %.3f ms, %.1f TFLOP/s\n", M, N, K, (int)grid.x, group_m, no_store ? " NO-STORE" : "", ms, 2.0 * M * N * K / ms / 1e9);
    return bad != 0;
}
