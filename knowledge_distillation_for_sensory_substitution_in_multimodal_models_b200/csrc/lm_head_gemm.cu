// lm_head_gemm.cu — the student lm_head feeding the vocabulary-sharded logit KD (BASELINE configs[4]; the logits the
// reference takes from the HF student at P1:228-240 are `hidden @ W_lm_head^T`, 896 x 151 936 for Qwen2-0.5B).
//
//   logits[n_rows, V_local] = hidden[n_rows, H] * weight[V_local, H]^T        bf16/fp16 operands, fp32 accumulation, 16-bit out
//   partials[tile][row]     = (max, sum 2^((x - max) * log2e / T)) of the row's ROUNDED logits inside column tile `tile`
//
// so that kd_vshard_stats does not read the student logits again (its teacher half is all that is left to stream).
//
// Tensor-core kernel for sm_100a, hand-written:
//   * persistent, one CTA per SM walking 128 x 256 output tiles in groups of 16 row tiles (a wave of CTAs touches ~16
//     hidden tiles and ~9 weight tiles: both stay in L2);
//   * operands arrive by 2-D TMA (SWIZZLE_128B, K blocks of 64) into a 4-stage ring, one full/empty mbarrier pair per stage;
//   * warp 0 lane 0 issues the TMA loads, warp 1 lane 0 issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = 256,
//     K = 16), tcgen05.commit releases the stage / publishes the accumulator;
//   * tensor memory holds TWO 128 x 256 fp32 accumulators (all 512 columns): warps 2-5 drain tile i (tcgen05.ld, thread = row,
//     online max / sum-exp, conversion, XOR-swizzled shared-memory staging, TMA store) while the MMAs of tile i + 1 run;
//   * ragged shapes: TMA zero-fills out-of-range rows / K columns on load and clips on store; the statistics mask columns
//     >= V_local themselves.
// Measured (profiles/r2_tcgen05_gemm_probe.txt, the standalone form tools/tcgen05_gemm_probe.cu): 16 384 x 151 936 x 896 in
// 3.26 ms = 1.37 PFLOP/s, the library GEMM 3.23 ms on the same box class; 16 384 x 8 192 x 3 584: 0.597 ms vs 0.612 ms.
#include <cuda.h>
#include <math.h>

#include "kd_common.cuh"

namespace kd {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SLAB = 64;                           // output columns per TMA store (one 128-byte swizzle atom)
constexpr int C_SLAB_BYTES = BM * SLAB * 2;
constexpr int GEMM_THREADS = 192;
constexpr int GROUP_M = 16;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 2 * C_SLAB_BYTES + 1024;

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
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

// what the teacher side of kd_vshard_stats needs from one 256-column tile of a row
struct TeacherTile {
    float m, s, a1, a2;
    int i1, i2, pad0, pad1;
};

// STATS: 0 = product only, 1 = + student (max, sum-exp) per tile, 2 = + the teacher's (max, sum-exp, best, runner-up) per tile
template <typename T, int STATS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
lm_head_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapT, int M, int N, int K, float k2,
                    float2* __restrict__ partials, int col0, TeacherTile* __restrict__ tparts) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 5];
    __shared__ uint32_t tmem_base_sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + STAGES * 8;
    const uint32_t bar_acc_full = bar_full + 2 * STAGES * 8, bar_acc_empty = bar_acc_full + 2 * 8;
    const uint32_t bar_teacher = bar_acc_empty + 2 * 8;
    const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN, tiles = tiles_m * tiles_n;
    const int kblocks = (K + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_acc_full + a * 8, 1); mbar_init(bar_acc_empty + a * 8, 4); }
        mbar_init(bar_teacher, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sm)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sm;

    if (warp == 0) {
        if (lane == 0) {      // ---------------- TMA producer ----------------
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int tile_m, tile_n;
                tile_coords(t, tiles_m, tiles_n, tile_m, tile_n);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bar_empty + s * 8, ph ^ 1u);
                    mbar_expect_tx(bar_full + s * 8, STAGE_BYTES);
                    const uint32_t dstA = smem_u32(smem + s * STAGE_BYTES), dstB = dstA + A_BYTES;
                    tma_load_2d(dstA, &mapA, kb * BK, tile_m * BM, bar_full + s * 8);
                    tma_load_2d(dstB, &mapB, kb * BK, tile_n * BN, bar_full + s * 8);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {      // ---------------- MMA issuer ----------------
            // D fp32 (bit 4), A/B format (bits 7-9 / 10-12), N >> 3 at bit 17, M >> 4 at bit 24; both operands K-major
            const uint32_t idesc = (1u << 4) | (Pack2<T>::kFormat << 7) | (Pack2<T>::kFormat << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            uint32_t it = 0, i = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
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
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                            "l"(da), "l"(db), "r"(idesc), "r"(acc)
                            : "memory");
                    }
                    // the stage returns to the producer once the MMAs that read it have retired
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_empty + s * 8)
                                 : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_acc_full + as * 8)
                             : "memory");
            }
        }
    } else {                  // ---------------- epilogue: warps 2..5 <-> TMEM lane quarters (warp % 4) ----------------
        const int q = warp & 3;
        const int r_in = q * 32 + lane;
        // STATS == 2 trades the second output staging buffer for a teacher slab (same 128 x 64 box, same swizzle), which the
        // elected thread fetches by TMA one slab ahead
        constexpr int CBUFS = (STATS == 2) ? 1 : 2;
        unsigned char* cstage = smem + STAGES * STAGE_BYTES;
        unsigned char* tbuf = cstage + C_SLAB_BYTES;
        uint32_t i = 0, slab = 0, tphase = 0;
        if constexpr (STATS == 2) {
            if (tid == 64 && (int)blockIdx.x < tiles) {
                int tm0, tn0;
                tile_coords(blockIdx.x, tiles_m, tiles_n, tm0, tn0);
                mbar_expect_tx(bar_teacher, C_SLAB_BYTES);
                tma_load_2d(smem_u32(tbuf), &mapT, tn0 * BN, tm0 * BM, bar_teacher);
            }
        }
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
            int tile_m, tile_n;
            tile_coords(t, tiles_m, tiles_n, tile_m, tile_n);
            const uint32_t as = i & 1, aph = (i >> 1) & 1;
            mbar_wait(bar_acc_full + as * 8, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            MaxSum ms, mt;
            Top2 top;
            ms.init();
            mt.init();
            top.init();
            const int row_g = tile_m * BM + r_in;
#pragma unroll 1
            for (int c = 0; c < BN; c += SLAB, ++slab) {
                uint32_t r[SLAB];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + as * BN + (uint32_t)c;
                tmem_ld32(taddr, r);
                tmem_ld32(taddr + 32, r + 32);
                uint4 traw[SLAB / 8];
                if constexpr (STATS == 2) {      // this row's 64 teacher cells from the TMA-staged slab (swizzled like the output stage)
                    mbar_wait(bar_teacher, tphase);
                    tphase ^= 1u;
#pragma unroll
                    for (int j = 0; j < SLAB / 8; ++j)
                        traw[j] = *reinterpret_cast<const uint4*>(tbuf + r_in * 128 + ((j ^ (r_in & 7)) << 4));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                unsigned char* buf = cstage + (CBUFS == 2 ? (slab & 1) : 0) * C_SLAB_BYTES;
                // the TMA store that last read this staging buffer must be done with it
                if (tid == 64) {
                    if constexpr (CBUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if constexpr (STATS == 2) {      // every thread holds its teacher cells: fetch the next slab (this tile's, or the next tile's first)
                    if (tid == 64) {
                        int nm = tile_m, nn = tile_n, nc = c + SLAB;
                        bool more = true;
                        if (nc == BN) {
                            nc = 0;
                            more = t + (int)gridDim.x < tiles;
                            if (more) tile_coords(t + gridDim.x, tiles_m, tiles_n, nm, nn);
                        }
                        if (more) {
                            mbar_expect_tx(bar_teacher, C_SLAB_BYTES);
                            tma_load_2d(smem_u32(tbuf), &mapT, nn * BN + nc, nm * BM, bar_teacher);
                        }
                    }
                }
                float vmax = -INFINITY;
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
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tid == 64) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&mapC), "r"(col_base),
                                 "r"(tile_m * BM), "r"(smem_u32(buf))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if constexpr (STATS != 0) {      // online (max, sum-exp) of this row over the slab's in-range columns
                    const int valid = min(SLAB, N - col_base);      // > 0: tiles start inside the matrix; N % 8 == 0
                    if (valid == SLAB) {
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
                    const float mk = ms.m * k2;
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int x = 0; x < SLAB; x += 2) {
                        s0 += ex2(fmaf(__uint_as_float(r[x]), k2, -mk));
                        s1 += ex2(fmaf(__uint_as_float(r[x + 1]), k2, -mk));
                    }
                    ms.s += s0 + s1;
                }
                if constexpr (STATS == 2) {      // the teacher's side of kd_vshard_stats for the same cells
                    const int valid_t = (row_g < M) ? min(SLAB, N - col_base) : 0;
#pragma unroll
                    for (int j = 0; j < SLAB / 8; ++j) {
                        if (8 * j < valid_t) {
                            float x[8];
                            Elem<T>::unpack(traw[j], x);
                            float vm = x[0];
#pragma unroll
                            for (int e = 1; e < 8; ++e) vm = fmaxf(vm, x[e]);
                            if (vm > mt.m) {
                                mt.s = (mt.m == -INFINITY) ? 0.f : mt.s * ex2((mt.m - vm) * k2);
                                mt.m = vm;
                            }
                            const float mk = mt.m * k2;
#pragma unroll
                            for (int e = 0; e < 8; ++e) mt.s += ex2(fmaf(x[e], k2, -mk));
                            if (vm >= top.a2) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) top.push(x[e], col0 + col_base + 8 * j + e);
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_acc_empty + as * 8) : "memory");
            if constexpr (STATS != 0) {
                if (row_g < M) partials[(size_t)tile_n * M + row_g] = make_float2(ms.m, ms.s);
            }
            if constexpr (STATS == 2) {
                if (row_g < M) {
                    uint4* o = reinterpret_cast<uint4*>(tparts + (size_t)tile_n * M + row_g);
                    o[0] = make_uint4(__float_as_uint(mt.m), __float_as_uint(mt.s), __float_as_uint(top.a1), __float_as_uint(top.a2));
                    o[1] = make_uint4((uint32_t)top.i1, (uint32_t)top.i2, 0u, 0u);
                }
            }
        }
        if (tid == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
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

// teacher + student: the whole 12-word kd_vshard_stats block of each row from the GEMM's per-tile partials
template <typename T>
__global__ void lm_head_stats_merge_kd_kernel(const float2* __restrict__ partials, const TeacherTile* __restrict__ tparts, int n_tiles,
                                              int64_t n_rows, float k2, const T* __restrict__ teacher, int64_t ld_t,
                                              const int64_t* __restrict__ labels, int col0, int V_local, float* __restrict__ stats,
                                              int stat_words) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    MaxSum ms, mt;
    Top2 top;
    ms.init();
    mt.init();
    top.init();
    for (int t = 0; t < n_tiles; ++t) {
        const float2 p = partials[(size_t)t * n_rows + r];
        ms.merge(p.x, p.y, k2);
        const uint4* tp = reinterpret_cast<const uint4*>(tparts + (size_t)t * n_rows + r);
        const uint4 a = tp[0], b = tp[1];
        mt.merge(__uint_as_float(a.x), __uint_as_float(a.y), k2);
        Top2 o;
        o.a1 = __uint_as_float(a.z); o.a2 = __uint_as_float(a.w); o.i1 = (int)b.x; o.i2 = (int)b.y;
        top.merge(o);
    }
    float x_label = -INFINITY;
    if (labels) {
        const int64_t lab = labels[r] - col0;
        if (lab >= 0 && lab < V_local) x_label = Elem<T>::to_float(teacher[r * ld_t + lab]);
    }
    float* o = stats + r * stat_words;
    o[0] = mt.m; o[1] = mt.s;
    o[2] = top.a1; o[3] = __int_as_float(top.i1); o[4] = top.a2; o[5] = __int_as_float(top.i2);
    o[6] = x_label;
    o[7] = ms.m; o[8] = ms.s;
    o[9] = o[10] = o[11] = 0.f;
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
int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& tmap, int M, int N, int K,
                float k2, float2* partials, int col0, TeacherTile* tparts, cudaStream_t st) {
    auto kern = lm_head_gemm_kernel<T, STATS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);      // per device, cheap
    if (e != cudaSuccess) return cuda_fail(e, "lm_head_gemm_kernel: shared-memory opt-in");
    const int64_t tiles = (int64_t)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = (int)(tiles < gemm_sm_count() ? tiles : gemm_sm_count());
    kern<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(a, b, c, tmap, M, N, K, k2, partials, col0, tparts);
    KD_CHECK_LAUNCH("lm_head_gemm_kernel");
    return KD_OK;
}

}  // namespace

}  // namespace kd

using namespace kd;

extern "C" {

int kd_lm_head_stat_tiles(int64_t V_local) { return V_local <= 0 ? 0 : (int)((V_local + BN - 1) / BN); }

static int lm_head_fwd_impl(const char* who, const void* hidden, int64_t ld_h, const void* weight, int64_t ld_w, void* logits,
                            int64_t ld_c, float* stat_partials, const void* teacher, int64_t ld_t, int64_t col0,
                            float* teacher_partials, int64_t n_rows, int64_t V_local, int64_t H, int dtype, float T, void* stream) {
    KD_CHECK_ARG(n_rows >= 0 && V_local >= 0 && H >= 1, "%s: bad shape", who);
    if (dtype != KD_DTYPE_BF16 && dtype != KD_DTYPE_F16) {
        set_error("%s: bf16 / fp16 operands only (tensor-core path); got dtype %d", who, dtype);
        return KD_ERR_UNSUPPORTED;
    }
    if (n_rows == 0 || V_local == 0) return KD_OK;
    KD_CHECK_ARG(hidden && weight && logits, "%s: null pointer", who);
    KD_CHECK_ARG(!stat_partials || T > 0.f, "%s: T must be positive", who);
    KD_CHECK_ARG(ld_h >= H && ld_w >= H && ld_c >= V_local, "%s: row pitch smaller than the row", who);
    KD_CHECK_ARG(!teacher_partials || (teacher && stat_partials && ld_t >= V_local && col0 >= 0),
                 "%s: the teacher statistics need the teacher columns and the student partials", who);
    bool aligned = ld_h % 8 == 0 && ld_w % 8 == 0 && ld_c % 8 == 0 && V_local % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(hidden) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(logits)) & 15u) == 0;
    if (teacher_partials) aligned = aligned && ld_t % 8 == 0 && (reinterpret_cast<uintptr_t>(teacher) & 15u) == 0;
    if (!aligned || n_rows >= (int64_t(1) << 31) || V_local >= (int64_t(1) << 31) || H >= (int64_t(1) << 31) ||
        col0 + V_local >= (int64_t(1) << 31)) {
        set_error("%s: TMA needs 16-byte aligned bases, row pitches and V_local that are multiples of 8 elements", who);
        return KD_ERR_UNSUPPORTED;
    }
    if (!encode_tiled()) {
        set_error("%s: the driver does not export cuTensorMapEncodeTiled", who);
        return KD_ERR_CUDA;
    }
    const CUtensorMapDataType dt = dtype == KD_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap ma, mb, mc, mt;
    if (!make_map(&ma, dt, hidden, n_rows, H, ld_h, BM, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
        !make_map(&mb, dt, weight, V_local, H, ld_w, BN, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
        !make_map(&mc, dt, logits, n_rows, V_local, ld_c, BM, CU_TENSOR_MAP_L2_PROMOTION_NONE) ||
        !make_map(&mt, dt, teacher_partials ? teacher : logits, n_rows, V_local, teacher_partials ? ld_t : ld_c, BM,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) {
        set_error("%s: cuTensorMapEncodeTiled rejected a matrix (n_rows %lld, V_local %lld, H %lld)", who, (long long)n_rows,
                  (long long)V_local, (long long)H);
        return KD_ERR_CUDA;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float k2 = stat_partials ? kLog2e / T : 0.f;
    float2* parts = reinterpret_cast<float2*>(stat_partials);
    TeacherTile* tparts = reinterpret_cast<TeacherTile*>(teacher_partials);
    const int M = (int)n_rows, N = (int)V_local, K = (int)H, c0 = (int)col0;
#define KD_HEAD_LAUNCH(TT)                                                                                      \
    (tparts ? launch_gemm<TT, 2>(ma, mb, mc, mt, M, N, K, k2, parts, c0, tparts, st)                             \
            : parts ? launch_gemm<TT, 1>(ma, mb, mc, mt, M, N, K, k2, parts, 0, nullptr, st)                    \
                    : launch_gemm<TT, 0>(ma, mb, mc, mt, M, N, K, k2, parts, 0, nullptr, st))
    return dtype == KD_DTYPE_BF16 ? KD_HEAD_LAUNCH(__nv_bfloat16) : KD_HEAD_LAUNCH(__half);
#undef KD_HEAD_LAUNCH
}

int kd_lm_head_fwd(const void* hidden, int64_t ld_h, const void* weight, int64_t ld_w, void* logits, int64_t ld_c,
                   float* stat_partials, int64_t n_rows, int64_t V_local, int64_t H, int dtype, float T, void* stream) {
    return lm_head_fwd_impl("kd_lm_head_fwd", hidden, ld_h, weight, ld_w, logits, ld_c, stat_partials, nullptr, 0, 0, nullptr, n_rows,
                            V_local, H, dtype, T, stream);
}

int kd_lm_head_fwd_kd(const void* hidden, int64_t ld_h, const void* weight, int64_t ld_w, void* logits, int64_t ld_c,
                      float* stat_partials, const void* teacher, int64_t ld_t, int64_t col0, float* teacher_partials, int64_t n_rows,
                      int64_t V_local, int64_t H, int dtype, float T, void* stream) {
    KD_CHECK_ARG(stat_partials && teacher_partials, "kd_lm_head_fwd_kd: both partial blocks are required");
    return lm_head_fwd_impl("kd_lm_head_fwd_kd", hidden, ld_h, weight, ld_w, logits, ld_c, stat_partials, teacher, ld_t, col0,
                            teacher_partials, n_rows, V_local, H, dtype, T, stream);
}

int kd_lm_head_stats_merge_kd(const float* stat_partials, const float* teacher_partials, int64_t n_tiles, const void* teacher,
                              int64_t ld_t, const int64_t* labels, int64_t col0, int64_t V_local, float* stats, int64_t n_rows,
                              int dtype, float T, void* stream) {
    KD_CHECK_ARG(n_rows >= 0 && n_tiles >= 1 && n_tiles < (1 << 30) && T > 0.f && V_local >= 1 && col0 >= 0,
                 "kd_lm_head_stats_merge_kd: bad arguments");
    KD_CHECK_ARG(dtype == KD_DTYPE_BF16 || dtype == KD_DTYPE_F16, "kd_lm_head_stats_merge_kd: bf16 / fp16 only");
    if (n_rows == 0) return KD_OK;
    KD_CHECK_ARG(stat_partials && teacher_partials && teacher && stats, "kd_lm_head_stats_merge_kd: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((n_rows + 127) / 128);
    const float2* sp = reinterpret_cast<const float2*>(stat_partials);
    const TeacherTile* tp = reinterpret_cast<const TeacherTile*>(teacher_partials);
    if (dtype == KD_DTYPE_BF16)
        lm_head_stats_merge_kd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(sp, tp, (int)n_tiles, n_rows, kLog2e / T,
            static_cast<const __nv_bfloat16*>(teacher), ld_t, labels, (int)col0, (int)V_local, stats, KD_VSHARD_STAT_WORDS);
    else
        lm_head_stats_merge_kd_kernel<__half><<<grid, 128, 0, st>>>(sp, tp, (int)n_tiles, n_rows, kLog2e / T,
            static_cast<const __half*>(teacher), ld_t, labels, (int)col0, (int)V_local, stats, KD_VSHARD_STAT_WORDS);
    KD_CHECK_LAUNCH("lm_head_stats_merge_kd_kernel");
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
