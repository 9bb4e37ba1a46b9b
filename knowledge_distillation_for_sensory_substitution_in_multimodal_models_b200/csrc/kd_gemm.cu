// kd_gemm.cu — the tensor-core GEMMs of the path: the student lm_head feeding the vocabulary-sharded logit KD (BASELINE
// configs[4]; the logits the reference takes from the HF student at P1:228-240 are `hidden @ W_lm_head^T`, 896 x 151 936 for
// Qwen2-0.5B), its two backward products, and the learned projections of BASELINE configs[2] / [3].
//
//   C[M, N] = A[M, K] * B[N, K]^T        bf16/fp16 operands, fp32 accumulation, 16-bit or fp32 out
//
// Either operand may be K-major (its K index contiguous in memory: A stored [M, K], B stored [N, K]) or MN-major (stored
// transposed: A as [K, M], B as [K, N]) — the tensor core reads both through its shared-memory descriptor, so no operand is
// ever transposed in HBM:   forward   logits = hidden * W^T                 A K-major, B K-major
//                           dH        = g * W           (B = W  as [K=V, N=H])   A K-major, B MN-major
//                           dW        = g^T * hidden    (A = g as [K=n, M=V], B = hidden as [K=n, N=H])   both MN-major
// Epilogues: plain store; + per-row (max, sum-exp) partials of the ROUNDED logits (so that kd_vshard_stats does not read the
// student logits again); or the feature-MSE epilogue (EV:160-165 on a projected tensor): the product is rounded like the
// library GEMM would, compared with the target tile, and what is stored is the GRADIENT w.r.t. the product.
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

#include <type_traits>

#include "kd_common.cuh"

namespace kd {

namespace {

// a CTA PAIR (cluster 2 x 1 x 1, the two SMs of a TPC) owns one 256 x 256 output tile: BM rows and half of the weight tile per CTA
constexpr int BM = 128, BN = 256, BNH = BN / 2, BK = 64, STAGES = 5;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNH * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SLAB = 64;                           // 16-bit output columns per TMA store (one 128-byte swizzle atom; fp32: 32)
constexpr int C_SLAB_BYTES = BM * 128;
constexpr int MN_GROUP_BYTES = BK * 128;           // MN-major operand: one 64-element MN group = BK rows of 128 bytes
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
// MN-major operand tile, 128-byte swizzle (canonical ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): a 128-byte row holds 64
// consecutive M (or N) elements of ONE k; 8 k-rows form the 1024-byte swizzle atom (SBO = 1024 B to the next 8 k), the next 64
// M / N elements start MN_GROUP_BYTES further (LBO).  One K = 16 MMA reads two atoms: its descriptor starts 2048 B after the last.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)(MN_GROUP_BYTES >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
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

// epilogues
constexpr int EPI_STORE = 0;      // C = round(acc)
constexpr int EPI_STATS = 1;      // + the rows' (max, sum-exp) per 128-column half tile (student lm_head)
constexpr int EPI_MSE = 2;        // C = upstream * 2 (round(acc) - target) / numel; partial sums of (round(acc) - target)^2
// contrastive_loss on the tensor cores (P1:406-414: logits = fs @ ft^T / temperature; cross_entropy against the diagonal)
constexpr int EPI_NCE_STATS = 3;  // nothing stored: the rows' (max, sum-exp) of the UNROUNDED similarities + the diagonal
constexpr int EPI_NCE_GRAD = 4;   // C = (softmax(acc / temperature) - I) * 2^14 from the rows' log-sum-exp (recomputed product)
constexpr float kNceGradScale = 16384.f;      // keeps softmax entries down to 4e-9 out of fp16's subnormal range; |C| <= 2^14

struct GemmExtra {                // what the non-plain epilogues need
    float2* partials;             // EPI_STATS: [stat tiles][M] (max, sum-exp)
    float k2;                     // EPI_STATS: log2(e) / T
    const void* target;           // EPI_MSE: [M, N] in the operand type, pitch ld_t elements
    int64_t ld_t;
    float gscale;                 // EPI_MSE: 2 / numel
    const float* upstream;        // EPI_MSE: device scalar or nullptr
    float* sq_partials;           // EPI_MSE: [tiles * 2 CTAs * EPI_HALVES * 4 warps] sums of squared differences
    float* diag;                  // EPI_NCE_STATS: [M] acc[i, i]
    const float* lse2;            // EPI_NCE_GRAD: [M] log2-sum-exp2 of the rows' acc * k2
    int max_ctas;                 // 0 = one CTA per SM; else at most this many (leaves SMs to a collective running beside the product)
};

template <typename OutT> struct OutPack;          // accumulator values -> the 16-byte chunks of a 128-byte staging row
template <> struct OutPack<float> { static constexpr int kCols = 32; };
template <> struct OutPack<__nv_bfloat16> { static constexpr int kCols = 64; };
template <> struct OutPack<__half> { static constexpr int kCols = 64; };

// T: operand type; EPI: epilogue; A_MN / B_MN: operand stored MN-major (transposed in memory); OutT: T or float
template <typename T, int EPI, bool A_MN, bool B_MN, typename OutT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
kd_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapT, int M, int N, int K, const GemmExtra ex) {
    constexpr int SLABC = OutPack<OutT>::kCols;     // output columns per 128-byte staging row
    constexpr bool OUT32 = sizeof(OutT) == 4;
    static_assert(EPI == EPI_STORE || !OUT32, "the statistics / MSE epilogues write the operand type");
    constexpr bool STATS = (EPI == EPI_STATS || EPI == EPI_NCE_STATS);
    constexpr bool STORE = (EPI != EPI_NCE_STATS);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 4 + EPI_HALVES];
    __shared__ uint32_t tmem_base_sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank;                                  // 0 = pair leader (issues the MMAs, owns the full / accumulator-empty barriers)
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int row_off = (int)rank * BM;             // this CTA's rows inside the pair's 256-row tile
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + STAGES * 8;
    const uint32_t bar_acc_full = bar_full + 2 * STAGES * 8, bar_acc_empty = bar_acc_full + 2 * 8;
    const uint32_t bar_tgt = bar_acc_empty + 2 * 8;            // EPI_MSE: "the target slab has landed in the staging buffer", one per column half
    const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + BN - 1) / BN, tiles = tiles_m * tiles_n;
    const int kblocks = (K + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_acc_full + a * 8, 1); mbar_init(bar_acc_empty + a * 8, 2 * 4 * EPI_HALVES); }      // every epilogue warp of both CTAs
        for (int h = 0; h < EPI_HALVES; ++h) mbar_init(bar_tgt + h * 8, 1);
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
                const int m0 = tile_m * 2 * BM + row_off, n0 = tile_n * BN + (int)rank * BNH;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bar_empty + s * 8, ph ^ 1u);
                    if (rank == 0) mbar_expect_tx(bar_full + s * 8, 2 * STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
                    const uint32_t leader_full = mapa_rank(bar_full + s * 8, 0);
                    const uint32_t dstA = smem_u32(smem + s * STAGE_BYTES), dstB = dstA + A_BYTES;
                    if constexpr (A_MN) {       // two boxes of {64 M elements, BK k-rows}: coordinates (m, k)
                        tma_load_2d_2sm(dstA, &mapA, m0, kb * BK, leader_full);
                        tma_load_2d_2sm(dstA + MN_GROUP_BYTES, &mapA, m0 + 64, kb * BK, leader_full);
                    } else {
                        tma_load_2d_2sm(dstA, &mapA, kb * BK, m0, leader_full);
                    }
                    if constexpr (B_MN) {
                        tma_load_2d_2sm(dstB, &mapB, n0, kb * BK, leader_full);
                        tma_load_2d_2sm(dstB + MN_GROUP_BYTES, &mapB, n0 + 64, kb * BK, leader_full);
                    } else {
                        tma_load_2d_2sm(dstB, &mapB, kb * BK, n0, leader_full);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {      // ---------------- MMA issuer (pair leader only) ----------------
            // D fp32 (bit 4), A/B format (bits 7-9 / 10-12), A / B MN-major (bits 15 / 16), N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | (Pack2<T>::kFormat << 7) | (Pack2<T>::kFormat << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
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
                        const uint64_t da = A_MN ? umma_desc_sw128_mn(a0 + k * 2048) : umma_desc_sw128(a0 + k * 32);
                        const uint64_t db = B_MN ? umma_desc_sw128_mn(b0 + k * 2048) : umma_desc_sw128(b0 + k * 32);
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
        float mse_g = 0.f;
        if constexpr (EPI == EPI_MSE) mse_g = ex.gscale * (ex.upstream ? *ex.upstream : 1.f);
        uint32_t i = 0, slab = 0;
        // EPI_MSE: the target slab of step `slab` is brought by TMA into the staging buffer the step will overwrite with its
        // output (same 128-byte swizzle: a thread finds its row's chunk j at j ^ (row % 8)), one step ahead
        auto fetch_target = [&](int tt, int c, uint32_t sl) {
            int tm, tn;
            tile_coords(tt, tiles_m, tiles_n, tm, tn);
            const uint32_t dst = smem_u32(cstage + (sl & 1) * C_SLAB_BYTES);
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");       // the store that last read this buffer is done with it
            mbar_expect_tx(bar_tgt + half * 8, C_SLAB_BYTES);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                         "l"(&mapT), "r"(tn * BN + c), "r"(tm * 2 * BM + row_off), "r"(bar_tgt + half * 8)
                         : "memory");
        };
        if constexpr (EPI == EPI_MSE) {
            if (elected && pair < tiles) fetch_target(pair, half * STAT_COLS, 0);
        }
        for (int t = pair; t < tiles; t += npairs, ++i) {
            int tile_m, tile_n;
            tile_coords(t, tiles_m, tiles_n, tile_m, tile_n);
            const uint32_t as = i & 1, aph = (i >> 1) & 1;
            mbar_wait(bar_acc_full + as * 8, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            MaxSum ms;
            ms.init();
            float sq = 0.f;
            const int row_g = tile_m * 2 * BM + row_off + r_in;
            [[maybe_unused]] float nce_nl = 0.f;      // EPI_NCE_GRAD: -log2-sum-exp2 of this thread's row
            if constexpr (EPI == EPI_NCE_GRAD) nce_nl = (row_g < M) ? -ex.lse2[row_g] : 0.f;
#pragma unroll 1
            for (int c = half * STAT_COLS; c < (half + 1) * STAT_COLS; c += SLABC, ++slab) {
                uint32_t r[SLABC];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + as * BN + (uint32_t)c;
                tmem_ld32(taddr, r);
                if constexpr (SLABC == 64) tmem_ld32(taddr + 32, r + 32);
                const int col_base = tile_n * BN + c;
                uint4 tg[EPI == EPI_MSE ? SLABC / 8 : 1];
                unsigned char* buf = cstage + (slab & 1) * C_SLAB_BYTES;
                if constexpr (EPI == EPI_MSE) {      // the target row's 128 bytes of this slab (TMA zero-fills outside the matrix)
                    mbar_wait(bar_tgt + half * 8, slab & 1);
#pragma unroll
                    for (int j = 0; j < SLABC / 8; ++j) tg[j] = *reinterpret_cast<const uint4*>(buf + r_in * 128 + ((j ^ (r_in & 7)) << 4));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");      // every thread holds its target chunks: the buffer may be overwritten
                } else if constexpr (!STORE) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // the TMA store that read this staging buffer two slabs ago must be done with it
                    if (elected) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                }
                if constexpr (EPI == EPI_NCE_STATS) {      // the diagonal entry of this thread's row, if the slab holds it
                    const int dcol = row_g - col_base;
                    if (dcol >= 0 && dcol < SLABC && row_g < M) {
                        float dv = 0.f;
#pragma unroll
                        for (int x = 0; x < SLABC; ++x) dv = (x == dcol) ? __uint_as_float(r[x]) : dv;
                        ex.diag[row_g] = dv;
                    }
                } else if constexpr (OUT32) {
#pragma unroll
                    for (int j = 0; j < SLABC / 4; ++j)
                        *reinterpret_cast<uint4*>(buf + r_in * 128 + ((j ^ (r_in & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < SLABC / 8; ++j) {
                        uint32_t w[4];
                        [[maybe_unused]] const uint32_t tw[4] = {tg[EPI == EPI_MSE ? j : 0].x, tg[EPI == EPI_MSE ? j : 0].y,
                                                                 tg[EPI == EPI_MSE ? j : 0].z, tg[EPI == EPI_MSE ? j : 0].w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float a = __uint_as_float(r[8 * j + 2 * u]), b = __uint_as_float(r[8 * j + 2 * u + 1]);
                            if constexpr (EPI == EPI_NCE_GRAD) {      // softmax from the row's log-sum-exp, minus the identity
                                const int cg = col_base + 8 * j + 2 * u;
                                a = ex2(fmaf(a, ex.k2, nce_nl));
                                b = ex2(fmaf(b, ex.k2, nce_nl));
                                if (cg == row_g) a -= 1.f;
                                if (cg + 1 == row_g) b -= 1.f;
                                a *= kNceGradScale;
                                b *= kNceGradScale;
                            }
                            w[u] = Pack2<T>::pack(a, b);             // a, b come back rounded to the operand type
                            if constexpr (EPI == EPI_STATS) {
                                r[8 * j + 2 * u] = __float_as_uint(a);
                                r[8 * j + 2 * u + 1] = __float_as_uint(b);
                            }
                            if constexpr (EPI == EPI_MSE) {           // (input.float() - target.float())^2 on the ROUNDED product
                                float ta, tb;
                                upk(Pair<T>::widen(tw[u]), ta, tb);
                                float da = a - ta, db = b - tb;
                                sq = fmaf(da, da, sq);
                                sq = fmaf(db, db, sq);
                                da *= mse_g;
                                db *= mse_g;
                                w[u] = Pack2<T>::pack(da, db);
                            }
                        }
                        // 16-byte chunk j of row r_in lands at chunk j ^ (r_in % 8): conflict-free, and what SWIZZLE_128B expects
                        *reinterpret_cast<uint4*>(buf + r_in * 128 + ((j ^ (r_in & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                if constexpr (STORE) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                }
                if (STORE && elected) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&mapC), "r"(col_base),
                                 "r"(tile_m * 2 * BM + row_off), "r"(smem_u32(buf))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if constexpr (EPI == EPI_MSE) {      // next step's target slab -> the other staging buffer
                        int nt = t, nc = c + SLABC;
                        if (nc >= (half + 1) * STAT_COLS) { nt = t + npairs; nc = half * STAT_COLS; }
                        if (nt < tiles) fetch_target(nt, nc, slab + 1);
                    }
                }
                if constexpr (STATS) {      // online (max, sum-exp) of this row over the slab's in-range columns
                    const float k2 = ex.k2;
                    const int valid = min(SLABC, N - col_base);      // may be <= 0 in the last tile: the whole slab is out of range
                    float vmax = -INFINITY;
                    if (valid >= SLABC) {
#pragma unroll
                        for (int x = 0; x < SLABC; ++x) vmax = fmaxf(vmax, __uint_as_float(r[x]));
                    } else {
#pragma unroll
                        for (int x = 0; x < SLABC; ++x) {
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
                        for (int x = 0; x < SLABC; x += 4) {
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
            if constexpr (STATS) {
                const int st = tile_n * EPI_HALVES + half;
                if (row_g < M && st < n_stat) ex.partials[(size_t)st * M + row_g] = make_float2(ms.m, ms.s);
            }
            if constexpr (EPI == EPI_MSE) {          // out-of-range rows / columns contributed (0 - 0)^2
                sq = warp_sum(sq);
                if (lane == 0) ex.sq_partials[((size_t)t * 2 + rank) * (EPI_HALVES * 4) + (warp - 2)] = sq;
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

// row-major [rows, cols] matrix of `esize`-byte elements with row pitch `ld` elements; boxes of box_cols x box_rows, 128-byte
// swizzle (box_cols * esize == 128)
bool make_map(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols,
              int box_rows, CUtensorMapL2promotion promo) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)esize};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
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

template <typename T, int EPI, bool A_MN, bool B_MN, typename OutT>
int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& tg, int M, int N, int K,
                const GemmExtra& ex, cudaStream_t st) {
    auto kern = kd_gemm_kernel<T, EPI, A_MN, B_MN, OutT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);      // per device, cheap
    if (e != cudaSuccess) return cuda_fail(e, "kd_gemm_kernel: shared-memory opt-in");
    const int64_t tiles = (int64_t)((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
    int64_t pairs = gemm_sm_count() / 2;
    if (ex.max_ctas > 0 && ex.max_ctas / 2 < pairs) pairs = ex.max_ctas / 2 > 0 ? ex.max_ctas / 2 : 1;
    const int grid = 2 * (int)(tiles < pairs ? tiles : pairs);      // persistent CTA pairs (cluster 2 x 1 x 1)
    kern<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(a, b, c, tg, M, N, K, ex);
    KD_CHECK_LAUNCH("kd_gemm_kernel");
    return KD_OK;
}

template <typename T>
int dispatch_gemm(int epi, bool a_mn, bool b_mn, bool out32, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                  const CUtensorMap& tg, int M, int N, int K, const GemmExtra& ex, cudaStream_t st) {
    if (epi == EPI_STATS) return launch_gemm<T, EPI_STATS, false, false, T>(a, b, c, tg, M, N, K, ex, st);
    if (epi == EPI_MSE) return launch_gemm<T, EPI_MSE, false, true, T>(a, b, c, tg, M, N, K, ex, st);
    if (epi == EPI_NCE_STATS || epi == EPI_NCE_GRAD) {      // fp16 operands only: unit vectors keep 11 mantissa bits there
        if constexpr (std::is_same<T, __half>::value) {
            if (epi == EPI_NCE_STATS) return launch_gemm<T, EPI_NCE_STATS, false, false, T>(a, b, c, tg, M, N, K, ex, st);
            return launch_gemm<T, EPI_NCE_GRAD, false, false, T>(a, b, c, tg, M, N, K, ex, st);
        } else {
            set_error("kd_gemm: the contrastive epilogues take fp16 operands");
            return KD_ERR_UNSUPPORTED;
        }
    }
    if (!a_mn && !b_mn) return out32 ? launch_gemm<T, EPI_STORE, false, false, float>(a, b, c, tg, M, N, K, ex, st)
                                     : launch_gemm<T, EPI_STORE, false, false, T>(a, b, c, tg, M, N, K, ex, st);
    if (!a_mn && b_mn) return out32 ? launch_gemm<T, EPI_STORE, false, true, float>(a, b, c, tg, M, N, K, ex, st)
                                    : launch_gemm<T, EPI_STORE, false, true, T>(a, b, c, tg, M, N, K, ex, st);
    if (a_mn && b_mn) return out32 ? launch_gemm<T, EPI_STORE, true, true, float>(a, b, c, tg, M, N, K, ex, st)
                                   : launch_gemm<T, EPI_STORE, true, true, T>(a, b, c, tg, M, N, K, ex, st);
    set_error("kd_gemm: A MN-major with B K-major is not built (transpose the roles: C^T = B A^T)");
    return KD_ERR_UNSUPPORTED;
}

// Shared front end: argument checks, tensor maps, dispatch.  A: logical [M, K]; B: logical [N, K]; *_mn = stored transposed.
int run_gemm(const char* who, int epi, const void* A, int64_t ld_a, bool a_mn, const void* B, int64_t ld_b, bool b_mn, void* C, int64_t ld_c,
             bool out32, int64_t M, int64_t N, int64_t K, int dtype, const GemmExtra& ex, cudaStream_t st) {
    if (dtype != KD_DTYPE_BF16 && dtype != KD_DTYPE_F16) {
        set_error("%s: bf16 / fp16 operands only (tensor-core path); got dtype %d", who, dtype);
        return KD_ERR_UNSUPPORTED;
    }
    if (M == 0 || N == 0) return KD_OK;
    KD_CHECK_ARG(A && B && C, "%s: null pointer", who);
    KD_CHECK_ARG(ld_a >= (a_mn ? M : K) && ld_b >= (b_mn ? N : K) && ld_c >= N, "%s: row pitch smaller than the row", who);
    const int64_t c_mult = out32 ? 4 : 8;
    const bool aligned = ld_a % 8 == 0 && ld_b % 8 == 0 && ld_c % c_mult == 0 && N % 8 == 0 && (!a_mn || M % 8 == 0) &&
                         ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15u) == 0;
    if (!aligned || M >= (int64_t(1) << 31) - 2 * BM || N >= (int64_t(1) << 31) - BN || K >= (int64_t(1) << 31) - BK) {
        set_error("%s: TMA needs 16-byte aligned bases and row pitches, and N (and M of an MN-major A) a multiple of 8 elements", who);
        return KD_ERR_UNSUPPORTED;
    }
    if (!encode_tiled()) {
        set_error("%s: the driver does not export cuTensorMapEncodeTiled", who);
        return KD_ERR_CUDA;
    }
    const CUtensorMapDataType dt = dtype == KD_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const auto promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUtensorMap ma, mb, mc;
    const bool ok_a = a_mn ? make_map(&ma, dt, 2, A, K, M, ld_a, 64, BK, promo) : make_map(&ma, dt, 2, A, M, K, ld_a, BK, BM, promo);
    const bool ok_b = b_mn ? make_map(&mb, dt, 2, B, K, N, ld_b, 64, BK, promo) : make_map(&mb, dt, 2, B, N, K, ld_b, BK, BNH, promo);
    const bool ok_c = out32 ? make_map(&mc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C, M, N, ld_c, 32, BM, CU_TENSOR_MAP_L2_PROMOTION_NONE)
                            : make_map(&mc, dt, 2, C, M, N, ld_c, 64, BM, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    CUtensorMap mt = mc;                      // EPI_MSE: the target, tiled like the output
    const bool ok_t = epi != EPI_MSE || make_map(&mt, dt, 2, ex.target, M, N, ex.ld_t, 64, BM, promo);
    if (!ok_a || !ok_b || !ok_c || !ok_t) {
        set_error("%s: cuTensorMapEncodeTiled rejected a matrix (M %lld, N %lld, K %lld)", who, (long long)M, (long long)N, (long long)K);
        return KD_ERR_CUDA;
    }
    if (dtype == KD_DTYPE_BF16) return dispatch_gemm<__nv_bfloat16>(epi, a_mn, b_mn, out32, ma, mb, mc, mt, (int)M, (int)N, (int)K, ex, st);
    return dispatch_gemm<__half>(epi, a_mn, b_mn, out32, ma, mb, mc, mt, (int)M, (int)N, (int)K, ex, st);
}

// loss = sum of the epilogue's partial sums / numel, in a fixed order
__global__ void proj_mse_finalize_kernel(const float* __restrict__ partials, int n, double inv_numel, float* __restrict__ loss_out) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += sh[i];
        *loss_out = (float)(tot * inv_numel);
    }
}

int64_t gemm_tiles(int64_t M, int64_t N) { return ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN); }

}  // namespace

// ---- contrastive_loss on the tensor cores (info_nce.cu drives these; fp16 unit vectors us, ut: [M, D]) ----------------------
int nce_stat_tiles(int64_t M) { return (int)((M + STAT_COLS - 1) / STAT_COLS); }

// rows' (max, sum 2^((x - max) k2)) partials per 128-column half tile of us . ut^T (never stored) + its diagonal
int nce_gemm_stats(const void* us16, const void* ut16, int64_t M, int64_t D, float k2, float* partials, float* diag, cudaStream_t st) {
    GemmExtra ex{};
    ex.partials = reinterpret_cast<float2*>(partials);
    ex.k2 = k2;
    ex.diag = diag;
    // C is never written by this epilogue: any aligned address makes a valid (unused) tensor map
    return run_gemm("kd_info_nce (similarity statistics)", EPI_NCE_STATS, us16, D, false, ut16, D, false, partials, M, false, M, M, D,
                    KD_DTYPE_F16, ex, st);
}

// G16[M, M] = fp16( (2^(us . ut^T k2 - lse2[row]) - I) * 2^14 ): the product is recomputed, only the gradient matrix is stored
int nce_gemm_grad(const void* us16, const void* ut16, int64_t M, int64_t D, float k2, const float* lse2, void* G16, cudaStream_t st) {
    GemmExtra ex{};
    ex.k2 = k2;
    ex.lse2 = lse2;
    return run_gemm("kd_info_nce (softmax - I)", EPI_NCE_GRAD, us16, D, false, ut16, D, false, G16, M, false, M, M, D, KD_DTYPE_F16, ex, st);
}

// dU[M, D] (fp32) = G16[M, M] . ut16[M, D]   (B = ut^T is stored MN-major)
int nce_gemm_du(const void* G16, const void* ut16, float* dU, int64_t M, int64_t D, cudaStream_t st) {
    GemmExtra ex{};
    return run_gemm("kd_info_nce (G . ut)", EPI_STORE, G16, M, false, ut16, D, true, dU, D, true, M, D, M, KD_DTYPE_F16, ex, st);
}

float nce_grad_scale() { return kNceGradScale; }

}  // namespace kd

using namespace kd;

extern "C" {

int kd_lm_head_stat_tiles(int64_t V_local) { return V_local <= 0 ? 0 : (int)((V_local + STAT_COLS - 1) / STAT_COLS); }

int kd_lm_head_fwd(const void* hidden, int64_t ld_h, const void* weight, int64_t ld_w, void* logits, int64_t ld_c,
                   float* stat_partials, int64_t n_rows, int64_t V_local, int64_t H, int dtype, float T, void* stream) {
    KD_CHECK_ARG(n_rows >= 0 && V_local >= 0 && H >= 1, "kd_lm_head_fwd: bad shape");
    KD_CHECK_ARG(!stat_partials || T > 0.f, "kd_lm_head_fwd: T must be positive");
    GemmExtra ex{};
    ex.partials = reinterpret_cast<float2*>(stat_partials);
    ex.k2 = stat_partials ? kLog2e / T : 0.f;
    return run_gemm("kd_lm_head_fwd", stat_partials ? EPI_STATS : EPI_STORE, hidden, ld_h, false, weight, ld_w, false, logits, ld_c, false,
                    n_rows, V_local, H, dtype, ex, static_cast<cudaStream_t>(stream));
}

int kd_gemm_ex(const void* A, int64_t ld_a, int a_mn_major, const void* B, int64_t ld_b, int b_mn_major, void* C, int64_t ld_c, int c_is_f32,
               int64_t M, int64_t N, int64_t K, int dtype, int max_ctas, void* stream) {
    KD_CHECK_ARG(M >= 0 && N >= 0 && K >= 1 && max_ctas >= 0, "kd_gemm: bad shape");
    GemmExtra ex{};
    ex.max_ctas = max_ctas;
    return run_gemm("kd_gemm", EPI_STORE, A, ld_a, a_mn_major != 0, B, ld_b, b_mn_major != 0, C, ld_c, c_is_f32 != 0, M, N, K, dtype, ex,
                    static_cast<cudaStream_t>(stream));
}

int kd_gemm(const void* A, int64_t ld_a, int a_mn_major, const void* B, int64_t ld_b, int b_mn_major, void* C, int64_t ld_c, int c_is_f32,
            int64_t M, int64_t N, int64_t K, int dtype, void* stream) {
    return kd_gemm_ex(A, ld_a, a_mn_major, B, ld_b, b_mn_major, C, ld_c, c_is_f32, M, N, K, dtype, 0, stream);
}

int kd_proj_mse_workspace_bytes(int64_t M, int64_t N, size_t* bytes) {
    KD_CHECK_ARG(bytes != nullptr && M >= 0 && N >= 0, "kd_proj_mse_workspace_bytes: bad arguments");
    *bytes = (size_t)gemm_tiles(M, N) * 2 * EPI_HALVES * 4 * sizeof(float) + 256;
    return KD_OK;
}

int kd_proj_mse_fwd_bwd(const void* input, int64_t ld_x, const void* proj, int64_t ld_w, const void* target, int64_t ld_t,
                        void* grad_projected, int64_t ld_g, float* loss_out, void* workspace, size_t workspace_bytes, int64_t M,
                        int64_t N, int64_t K, int dtype, const float* upstream, void* stream) {
    KD_CHECK_ARG(M >= 1 && N >= 1 && K >= 1, "kd_proj_mse_fwd_bwd: bad shape");
    KD_CHECK_ARG(target && loss_out && workspace, "kd_proj_mse_fwd_bwd: null pointer");
    KD_CHECK_ARG(ld_t >= N && ld_t % 8 == 0 && (reinterpret_cast<uintptr_t>(target) & 15u) == 0, "kd_proj_mse_fwd_bwd: target rows must be 16-byte aligned");
    size_t need = 0;
    kd_proj_mse_workspace_bytes(M, N, &need);
    if (workspace_bytes < need) {
        set_error("kd_proj_mse_fwd_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GemmExtra ex{};
    ex.target = target;
    ex.ld_t = ld_t;
    ex.gscale = (float)(2.0 / ((double)M * (double)N));
    ex.upstream = upstream;
    ex.sq_partials = static_cast<float*>(workspace);
    // projected[M, N] = input[M, K] * proj[K, N]: B = proj^T is stored MN-major
    if (int rc = run_gemm("kd_proj_mse_fwd_bwd", EPI_MSE, input, ld_x, false, proj, ld_w, true, grad_projected, ld_g, false, M, N, K, dtype,
                          ex, st))
        return rc;
    const int n_part = (int)(gemm_tiles(M, N) * 2 * EPI_HALVES * 4);
    proj_mse_finalize_kernel<<<1, 1024, 0, st>>>(ex.sq_partials, n_part, 1.0 / ((double)M * (double)N), loss_out);
    KD_CHECK_LAUNCH("proj_mse_finalize_kernel");
    return KD_OK;
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
