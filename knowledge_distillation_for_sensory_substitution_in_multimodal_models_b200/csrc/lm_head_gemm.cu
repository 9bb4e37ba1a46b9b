// lm_head_gemm.cu — the student lm_head feeding the vocabulary-sharded logit KD (BASELINE configs[4]; the logits the
// reference takes from the HF student at P1:228-240 are `hidden @ W_lm_head^T`, 896 x 151 936 for Qwen2-0.5B).
//
//   logits[n_rows, V_local] = hidden[n_rows, H] * weight[V_local, H]^T        bf16/fp16 operands, fp32 accumulation, 16-bit out
//   partials[tile][row]     = (max, sum 2^((x - max) * log2e / T)) of the row's ROUNDED logits inside 128-column tile `tile`
//
// so that kd_vshard_stats does not read the student logits again (its teacher half is all that is left to stream).
//
// Tensor-core kernel for sm_100a, hand-written:
//   * persistent CTA PAIRS (cluster 2 x 1 x 1 = the two SMs of a TPC), one 256 x 256 output tile per pair at a time, walked in
//     groups of 8 row tiles (a wave touches ~8 x 256 hidden rows and ~9 weight tiles: both stay in L2);
//   * tcgen05.mma.cta_group::2.kind::f16 (M = 256 over the pair, N = 256, K = 16) issued by ONE lane of the even CTA; each CTA
//     feeds its 128 hidden rows and HALF of the weight tile (32 KB per 64-wide K block instead of 48 KB), so the shared-memory
//     fill per SM drops by a third — what matters on a power-bound GEMM;
//   * operands arrive by 2-D TMA (SWIZZLE_128B) into a 5-stage ring in each CTA; the loads carry .cta_group::2 and complete on
//     the LEADER's full barrier, tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs;
//   * tensor memory holds TWO 128 x 256 fp32 accumulators per CTA (all 512 columns): eight epilogue warps (two per TMEM lane
//     quarter, 128 columns each) drain tile i — tcgen05.ld with thread = row, rounding, online max / sum-exp, XOR-swizzled
//     shared-memory staging, TMA stores, two slabs in flight per half — while the MMAs of tile i + 1 run; they hand the
//     accumulator back on the leader's barrier (remote mbarrier arrive);
//   * ragged shapes: TMA zero-fills out-of-range rows / K columns on load and clips on store; the statistics mask columns
//     >= V_local themselves.
// Measured: DESIGN.md §4.3, profiles/r2_tcgen05_gemm_probe.txt (development path: tools/tcgen05_probe.cu -> tcgen05_gemm_probe.cu
// -> tcgen05_gemm2_probe.cu).
#include <cuda.h>
#include <math.h>

#include "kd_common.cuh"

namespace kd {

namespace {

// a CTA PAIR (cluster 2 x 1 x 1, the two SMs of a TPC) owns one 256 x 256 output tile: BM rows and half of the weight tile per CTA
constexpr int BM = 128, BN = 256, BNH = BN / 2, BK = 64, STAGES = 5;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNH * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SLAB = 64;                           // output columns per TMA store (one 128-byte swizzle atom)
constexpr int C_SLAB_BYTES = BM * SLAB * 2;
constexpr int EPI_HALVES = 2;                     // two epilogue warp quads, 128 of the tile's 256 columns each
constexpr int STAT_COLS = BN / EPI_HALVES;        // columns per statistics partial
constexpr int GEMM_THREADS = 64 + EPI_HALVES * 128;
constexpr int GROUP_M = 8;                        // in 256-row pair tiles
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + EPI_HALVES * 2 * C_SLAB_BYTES + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {      // the same shared-memory offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// 2-SM TMA load: the data lands in THIS CTA's shared memory, the completion bytes on the pair leader's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
                 : "memory");
}
// arrives on `bar` (same offset) in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major operand tile, 128-byte swizzle: 128-byte rows, 8-row groups 1024 B apart (SBO); LBO is unused for this layout
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;      // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int& tile_m, int& tile_n) {
    const int per_group = GROUP_M * tiles_n;
    const int g = t / per_group, first_m = g * GROUP_M;
    const int gm = min(GROUP_M, tiles_m - first_m);
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
        : "r"(taddr)
        : "memory");
}

template <typename T> struct Pack2;
template <> struct Pack2<__nv_bfloat16> {
    static constexpr uint32_t kFormat = 1;          // instruction-descriptor operand format: BF16
    // packs (a, b) and returns them as rounded
    static __device__ __forceinline__ uint32_t pack(float& a, float& b) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        const float2 f = __bfloat1622float2(h);
        a = f.x; b = f.y;
        return *reinterpret_cast<uint32_t*>(&h);
    }
};
template <> struct Pack2<__half> {
    static constexpr uint32_t kFormat = 0;          // F16
    static __device__ __forceinline__ uint32_t pack(float& a, float& b) {
        __half2 h = __floats2half2_rn(a, b);
        const float2 f = __half22float2(h);
        a = f.x; b = f.y;
        return *reinterpret_cast<uint32_t*>(&h);
    }
};

// STATS: 0 = product only, 1 = + the rows' (max, sum-exp) per 128-column half tile
template <typename T, int STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
lm_head_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ CUtensorMap mapC, int M, int N, int K, float k2, float2* __restrict__ partials) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 4];
    __shared__ uint32_t tmem_base_sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank;                                  // 0 = pair leader (issues the MMAs, owns the full / accumulator-empty barriers)
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int row_off = (int)rank * BM;             // this CTA's rows inside the pair's 256-row tile
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + STAGES * 8;
    const uint32_t bar_acc_full = bar_full + 2 * STAGES * 8, bar_acc_empty = bar_acc_full + 2 * 8;
    const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + BN - 1) / BN, tiles = tiles_m * tiles_n;
    const int kblocks = (K + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_acc_full + a * 8, 1); mbar_init(bar_acc_empty + a * 8, 2 * 4 * EPI_HALVES); }      // every epilogue warp of both CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sm)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                             // barriers initialised and tensor memory allocated in both CTAs
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sm;

    if (warp == 0) {
        if (lane == 0) {      // ---------------- TMA producer (both CTAs) ----------------
            uint32_t it = 0;
            for (int t = pair; t < tiles; t += npairs) {
                int tile_m, tile_n;
                tile_coords(t, tiles_m, tiles_n, tile_m, tile_n);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bar_empty + s * 8, ph ^ 1u);
                    if (rank == 0) mbar_expect_tx(bar_full + s * 8, 2 * STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
                    const uint32_t leader_full = mapa_rank(bar_full + s * 8, 0);
                    const uint32_t dstA = smem_u32(smem + s * STAGE_BYTES), dstB = dstA + A_BYTES;
                    tma_load_2d_2sm(dstA, &mapA, kb * BK, tile_m * 2 * BM + row_off, leader_full);
                    tma_load_2d_2sm(dstB, &mapB, kb * BK, tile_n * BN + (int)rank * BNH, leader_full);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {      // ---------------- MMA issuer (pair leader only) ----------------
            // D fp32 (bit 4), A/B format (bits 7-9 / 10-12), N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 over the pair); K-major
            const uint32_t idesc = (1u << 4) | (Pack2<T>::kFormat << 7) | (Pack2<T>::kFormat << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)((2 * BM) >> 4) << 24);
            uint32_t it = 0, i = 0;
            for (int t = pair; t < tiles; t += npairs, ++i) {
                const uint32_t as = i & 1, aph = (i >> 1) & 1;
                mbar_wait(bar_acc_empty + as * 8, aph ^ 1u);      // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + as * BN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bar_full + s * 8, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = umma_desc_sw128(a0 + k * 32), db = umma_desc_sw128(b0 + k * 32);
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                            "l"(da), "l"(db), "r"(idesc), "r"(acc)
                            : "memory");
                    }
                    commit_pair(bar_empty + s * 8);      // the stage returns to both producers once the MMAs that read it retire
                }
                commit_pair(bar_acc_full + as * 8);
            }
        }
    } else {                  // ---------------- epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ----------------
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int r_in = q * 32 + lane;
        const bool elected = (tid == 64 + half * 128);                 // issues this half's TMA stores
        unsigned char* cstage = smem + STAGES * STAGE_BYTES + half * 2 * C_SLAB_BYTES;
        const int n_stat = (N + STAT_COLS - 1) / STAT_COLS;
        uint32_t i = 0, slab = 0;
        for (int t = pair; t < tiles; t += npairs, ++i) {
            int tile_m, tile_n;
            tile_coords(t, tiles_m, tiles_n, tile_m, tile_n);
            const uint32_t as = i & 1, aph = (i >> 1) & 1;
            mbar_wait(bar_acc_full + as * 8, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            MaxSum ms;
            ms.init();
            const int row_g = tile_m * 2 * BM + row_off + r_in;
#pragma unroll 1
            for (int c = half * STAT_COLS; c < (half + 1) * STAT_COLS; c += SLAB, ++slab) {
                uint32_t r[SLAB];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + as * BN + (uint32_t)c;
                tmem_ld32(taddr, r);
                tmem_ld32(taddr + 32, r + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                unsigned char* buf = cstage + (slab & 1) * C_SLAB_BYTES;
                // the TMA store that read this staging buffer two slabs ago must be done with it
                if (elected) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                const int col_base = tile_n * BN + c;
#pragma unroll
                for (int j = 0; j < SLAB / 8; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float a = __uint_as_float(r[8 * j + 2 * u]), b = __uint_as_float(r[8 * j + 2 * u + 1]);
                        w[u] = Pack2<T>::pack(a, b);
                        if constexpr (STATS != 0) {
                            r[8 * j + 2 * u] = __float_as_uint(a);
                            r[8 * j + 2 * u + 1] = __float_as_uint(b);
                        }
                    }
                    // 16-byte chunk j of row r_in lands at chunk j ^ (r_in % 8): conflict-free, and what SWIZZLE_128B expects
                    *reinterpret_cast<uint4*>(buf + r_in * 128 + ((j ^ (r_in & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                if (elected) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&mapC), "r"(col_base),
                                 "r"(tile_m * 2 * BM + row_off), "r"(smem_u32(buf))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if constexpr (STATS != 0) {      // online (max, sum-exp) of this row over the slab's in-range columns
                    const int valid = min(SLAB, N - col_base);      // may be <= 0 in the last tile: the whole slab is out of range
                    float vmax = -INFINITY;
                    if (valid >= SLAB) {
#pragma unroll
                        for (int x = 0; x < SLAB; ++x) vmax = fmaxf(vmax, __uint_as_float(r[x]));
                    } else {
#pragma unroll
                        for (int x = 0; x < SLAB; ++x) {
                            if (x >= valid) r[x] = 0xff800000u;      // -inf
                            vmax = fmaxf(vmax, __uint_as_float(r[x]));
                        }
                    }
                    if (vmax > ms.m) {
                        ms.s = (ms.m == -INFINITY) ? 0.f : ms.s * ex2((ms.m - vmax) * k2);
                        ms.m = vmax;
                    }
                    if (ms.m != -INFINITY) {
                        const float mk = ms.m * k2;
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int x = 0; x < SLAB; x += 4) {
                            s0 += ex2(fmaf(__uint_as_float(r[x]), k2, -mk));
                            s1 += ex2(fmaf(__uint_as_float(r[x + 1]), k2, -mk));
                            s2 += ex2(fmaf(__uint_as_float(r[x + 2]), k2, -mk));
                            s3 += ex2(fmaf(__uint_as_float(r[x + 3]), k2, -mk));
                        }
                        ms.s += (s0 + s1) + (s2 + s3);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0)      // on the leader's barrier: its MMA warp reuses the accumulator of BOTH CTAs
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_rank(bar_acc_empty + as * 8, 0)) : "memory");
            if constexpr (STATS != 0) {
                const int st = tile_n * EPI_HALVES + half;
                if (row_g < M && st < n_stat) partials[(size_t)st * M + row_g] = make_float2(ms.m, ms.s);
            }
        }
        if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                             // neither CTA may leave (or free tensor memory) while its peer still uses the pair
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// merges the per-tile student (max, sum-exp) of each row into words 7, 8 of its kd_vshard_stats block
__global__ void lm_head_stats_merge_kernel(const float2* __restrict__ partials, int n_tiles, int64_t n_rows, float k2,
                                           float* __restrict__ stats, int stat_words) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    MaxSum ms;
    ms.init();
    for (int t = 0; t < n_tiles; ++t) {
        const float2 p = partials[(size_t)t * n_rows + r];
        ms.merge(p.x, p.y, k2);
    }
    stats[r * stat_words + 7] = ms.m;
    stats[r * stat_words + 8] = ms.s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// row-major [rows, cols] 16-bit matrix with row pitch `ld` elements; boxes of 64 columns x box_rows rows, 128-byte swizzle
bool make_map(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows,
              CUtensorMapL2promotion promo) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return encode_tiled()(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int gemm_sm_count() {
    static int sms = [] {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n > 0 ? n : 148;
    }();
    return sms;
}

template <typename T, int STATS>
int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, int M, int N, int K, float k2, float2* partials,
                cudaStream_t st) {
    auto kern = lm_head_gemm_kernel<T, STATS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);      // per device, cheap
    if (e != cudaSuccess) return cuda_fail(e, "lm_head_gemm_kernel: shared-memory opt-in");
    const int64_t tiles = (int64_t)((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
    const int64_t pairs = gemm_sm_count() / 2;
    const int grid = 2 * (int)(tiles < pairs ? tiles : pairs);      // persistent CTA pairs (cluster 2 x 1 x 1)
    kern<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(a, b, c, M, N, K, k2, partials);
    KD_CHECK_LAUNCH("lm_head_gemm_kernel");
    return KD_OK;
}

}  // namespace

}  // namespace kd

using namespace kd;

extern "C" {

int kd_lm_head_stat_tiles(int64_t V_local) { return V_local <= 0 ? 0 : (int)((V_local + STAT_COLS - 1) / STAT_COLS); }

int kd_lm_head_fwd(const void* hidden, int64_t ld_h, const void* weight, int64_t ld_w, void* logits, int64_t ld_c,
                   float* stat_partials, int64_t n_rows, int64_t V_local, int64_t H, int dtype, float T, void* stream) {
    KD_CHECK_ARG(n_rows >= 0 && V_local >= 0 && H >= 1, "kd_lm_head_fwd: bad shape");
    if (dtype != KD_DTYPE_BF16 && dtype != KD_DTYPE_F16) {
        set_error("kd_lm_head_fwd: bf16 / fp16 operands only (tensor-core path); got dtype %d", dtype);
        return KD_ERR_UNSUPPORTED;
    }
    if (n_rows == 0 || V_local == 0) return KD_OK;
    KD_CHECK_ARG(hidden && weight && logits, "kd_lm_head_fwd: null pointer");
    KD_CHECK_ARG(!stat_partials || T > 0.f, "kd_lm_head_fwd: T must be positive");
    KD_CHECK_ARG(ld_h >= H && ld_w >= H && ld_c >= V_local, "kd_lm_head_fwd: row pitch smaller than the row");
    const bool aligned = ld_h % 8 == 0 && ld_w % 8 == 0 && ld_c % 8 == 0 && V_local % 8 == 0 &&
                         ((reinterpret_cast<uintptr_t>(hidden) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(logits)) & 15u) == 0;
    if (!aligned || n_rows >= (int64_t(1) << 31) - 2 * BM || V_local >= (int64_t(1) << 31) - BN || H >= (int64_t(1) << 31) - BK) {
        set_error("kd_lm_head_fwd: TMA needs 16-byte aligned bases, row pitches and V_local that are multiples of 8 elements");
        return KD_ERR_UNSUPPORTED;
    }
    if (!encode_tiled()) {
        set_error("kd_lm_head_fwd: the driver does not export cuTensorMapEncodeTiled");
        return KD_ERR_CUDA;
    }
    const CUtensorMapDataType dt = dtype == KD_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap ma, mb, mc;
    if (!make_map(&ma, dt, hidden, n_rows, H, ld_h, BM, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
        !make_map(&mb, dt, weight, V_local, H, ld_w, BNH, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
        !make_map(&mc, dt, logits, n_rows, V_local, ld_c, BM, CU_TENSOR_MAP_L2_PROMOTION_NONE)) {
        set_error("kd_lm_head_fwd: cuTensorMapEncodeTiled rejected a matrix (n_rows %lld, V_local %lld, H %lld)", (long long)n_rows,
                  (long long)V_local, (long long)H);
        return KD_ERR_CUDA;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float k2 = stat_partials ? kLog2e / T : 0.f;
    float2* parts = reinterpret_cast<float2*>(stat_partials);
    const int M = (int)n_rows, N = (int)V_local, K = (int)H;
    if (dtype == KD_DTYPE_BF16)
        return parts ? launch_gemm<__nv_bfloat16, 1>(ma, mb, mc, M, N, K, k2, parts, st)
                     : launch_gemm<__nv_bfloat16, 0>(ma, mb, mc, M, N, K, k2, parts, st);
    return parts ? launch_gemm<__half, 1>(ma, mb, mc, M, N, K, k2, parts, st) : launch_gemm<__half, 0>(ma, mb, mc, M, N, K, k2, parts, st);
}

int kd_lm_head_stats_merge(const float* stat_partials, int64_t n_tiles, float* stats, int64_t n_rows, float T, void* stream) {
    KD_CHECK_ARG(n_rows >= 0 && n_tiles >= 1 && n_tiles < (1 << 30) && T > 0.f, "kd_lm_head_stats_merge: bad arguments");
    if (n_rows == 0) return KD_OK;
    KD_CHECK_ARG(stat_partials && stats, "kd_lm_head_stats_merge: null pointer");
    lm_head_stats_merge_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(stat_partials), (int)n_tiles, n_rows, kLog2e / T, stats, KD_VSHARD_STAT_WORDS);
    KD_CHECK_LAUNCH("lm_head_stats_merge_kernel");
    return KD_OK;
}

}  // extern "C"
