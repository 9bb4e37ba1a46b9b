// logit_kd_vshard.cu — vocabulary-sharded execution of the logit-KD family (BASELINE configs[4]).
//
// Every rank holds ALL n_rows token rows but only columns [col0, col0 + V_local) of the teacher and student logits
// (what a vocabulary-parallel lm_head produces).  The per-row quantities that couple the shards cross ranks twice:
//
//   1. kd_vshard_stats     local (max, sum-exp) of both rows, local best / runner-up teacher cells (global column
//                          indices), the teacher logit at the label if this shard owns it        -> 12 floats per row
//      [all-gather of the stats blocks: G x n_rows x 48 B]
//   2. kd_vshard_combine   merges the G blocks per row: lse2_t, lse2_s, the token record {label, v1, k, v2}
//                          (P1:166-185) and the row's own xlogy(v,v) constant (per-row mode)
//      [as-written mode: kd_loca_build_table on the records — every rank holds all rows, so no exchange]
//   3. kd_vshard_accum     local part of the row sum and of Q_r = sum_v q*[sp >= eps]              -> 2 floats per row
//      [all-reduce(SUM) of Q_r: n_rows floats; the row sums are summed into the scalar loss]
//   4. kd_vshard_emit      gradient of the local columns from (lse2_t, lse2_s, Q_r)
//
// Same per-cell math as the single-GPU kernels (PointTerm, logit_kd.cuh).  These are plain streaming kernels: the mode
// exists for logits that are produced sharded; its exchanges are O(n_rows) and latency-bound.
#include <math.h>

#include "logit_kd.cuh"

// tuning knobs (tools/build_variant.sh --src logit_kd_vshard.cu): vector pairs per lane fetched before any is consumed in the accumulate /
// emit sweeps, and loads per lane and batch in the statistics walk (8 against 4: 1.046 -> 1.009 ms for both tensors of a configs[4] shard)
#ifndef KD_VS_ACCUM_BATCH
#define KD_VS_ACCUM_BATCH 2      // accumulate sweep: 2 against 4 = 0.931 -> 0.838 ms (fewer registers, more resident warps)
#endif
#ifndef KD_VS_EMIT_BATCH
#define KD_VS_EMIT_BATCH 5       // emit sweep: 2 / 3 / 4 / 5 / 6 = 1.315 / 1.310 / 1.213 / 1.150 / 1.496 ms (floor 1.140)
#endif
#ifndef KD_VS_STATS_U
#define KD_VS_STATS_U 8
#endif

namespace kd {

namespace {

constexpr int kVsThreads = 256;
constexpr int kVsWarps = kVsThreads / 32;
constexpr int kStatWords = 12;      // m_t, S_t, a1, i1, a2, i2, x_label, m_s, S_s, pad x3

template <typename T>
__device__ __forceinline__ bool vec_ok(const T* row, int n) {        // 16-byte accesses over the whole local row
    return (reinterpret_cast<uintptr_t>(row) & 15u) == 0 && n % Elem<T>::kPack == 0;
}

// `lane` of `stride` cooperating threads (a warp or the whole block) walks the row
template <typename T>
__device__ __forceinline__ void shard_row_stats(const T* __restrict__ row, int n, float k2, MaxSum& ms, Top2* top, int col0, int lane,
                                                int stride) {
    constexpr int P = Elem<T>::kPack;
    ms.init();
    if (top) top->init();
    auto one = [&](float x, int c) {
        if (x > ms.m) {
            ms.s = (ms.m == -INFINITY) ? 0.f : ms.s * ex2((ms.m - x) * k2);
            ms.m = x;
        }
        ms.s += ex2((x - ms.m) * k2);
        if (top) top->push(x, col0 + c);
    };
    if (vec_ok(row, n)) {
        const uint4* rv = reinterpret_cast<const uint4*>(row);
        auto vec = [&](const uint4& raw, int v) {
            float x[P];
            Elem<T>::unpack(raw, x);
            float vm = x[0];
#pragma unroll
            for (int i = 1; i < P; ++i) vm = fmaxf(vm, x[i]);
            if (vm > ms.m) {                      // one rescale per vector instead of one per cell
                ms.s = (ms.m == -INFINITY) ? 0.f : ms.s * ex2((ms.m - vm) * k2);
                ms.m = vm;
            }
            const float mk = ms.m * k2;
#pragma unroll
            for (int i = 0; i < P; ++i) ms.s += ex2(fmaf(x[i], k2, -mk));
            if (top && vm >= top->a2) {
#pragma unroll
                for (int i = 0; i < P; ++i) top->push(x[i], col0 + v * P + i);
            }
        };
        const int nv = n / P;
        int v = lane;
        for (; v + 3 * stride < nv; v += 4 * stride) {      // four loads in flight per thread
            const uint4 a0 = ldg_stream(rv + v), a1 = ldg_stream(rv + v + stride), a2 = ldg_stream(rv + v + 2 * stride),
                        a3 = ldg_stream(rv + v + 3 * stride);
            vec(a0, v);
            vec(a1, v + stride);
            vec(a2, v + 2 * stride);
            vec(a3, v + 3 * stride);
        }
        for (; v < nv; v += stride) vec(ldg_stream(rv + v), v);
    } else {
        for (int c = lane; c < n; c += stride) one(Elem<T>::to_float(row[c]), c);
    }
}

// 16-bit logits, 16-byte aligned rows: the scheme of phase A's row_stats_packed (logit_kd.cu) for a lane of a warp-per-row / block-
// per-row walk — packed 16-bit maxima, fp32x2 arithmetic, a LAZY rescale (only when a cell exceeds the running reference by more
// than `slack`, so 2^((x - ref) k2) stays finite), eight loads in flight per lane, and runner-up candidates behind a threshold
// seeded from the first batch (the 2nd largest lane maximum of the warp: two distinct cells reach it, so it bounds the row's
// runner-up from below).  The generic walk above costs ~29 issue slots per cell (measured: issue-bound at 88 %); this one ~6.
template <typename T, bool WANT_TOP2>
__device__ __forceinline__ void shard_row_stats_packed(const T* __restrict__ row, int n, float k2, MaxSum& ms, Top2& top, int col0,
                                                       int lane, int stride) {
    ms.init();
    if (WANT_TOP2) top.init();
    const int nv = n / 8;
    const uint4* rv = reinterpret_cast<const uint4*>(row);
    const u64 k2p = pk(k2, k2);
    const float slack = 64.f / k2;
    float ref = -INFINITY, lane_max = -INFINITY, thr_top = -INFINITY;
    u64 mkp = pk(0.f, 0.f), acc0 = pk(0.f, 0.f), acc1 = pk(0.f, 0.f);
    auto vec_max = [&](const uint4& raw) -> float {
        const uint32_t m2 = Pair<T>::vmax(Pair<T>::vmax(raw.x, raw.y), Pair<T>::vmax(raw.z, raw.w));
        float m_lo, m_hi;
        upk(Pair<T>::widen(m2), m_lo, m_hi);
        return fmaxf(m_lo, m_hi);
    };
    // The hot path is straight-line code: a vector that may hold a best / runner-up cell only sets its bit in `cand`; the
    // (rare, divergent) per-cell insertion runs once per batch from ONE copy of its code, re-reading the vector (an L1 / L2 hit).
    // Inlined into every unrolled copy of the sweep it pushed the loop out of the instruction cache (ncu: "no instruction" 15.7
    // stall cycles per issue, the kernel 1.7x SLOWER than the generic walk).
    auto consume = [&](const uint4& raw, unsigned bit, unsigned& cand) {
        const float vm = vec_max(raw);
        lane_max = fmaxf(lane_max, vm);
        if (vm > ref + slack || ref == -INFINITY) {     // rare: first vector, or a cell far above the reference
            const float sc = (ref == -INFINITY) ? 0.f : ex2((ref - vm) * k2);
            acc0 = mul2(acc0, pk(sc, sc));
            acc1 = mul2(acc1, pk(sc, sc));
            ref = vm;
            mkp = pk(-vm * k2, -vm * k2);
        }
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float e0, e1;
            upk(fma2(Pair<T>::widen(w[i]), k2p, mkp), e0, e1);
            if (i & 1) acc1 = add2(acc1, pk(ex2(e0), ex2(e1)));
            else acc0 = add2(acc0, pk(ex2(e0), ex2(e1)));
        }
        if (WANT_TOP2 && vm >= thr_top) cand |= bit;
    };
    constexpr int U = KD_VS_STATS_U;
    const int wl = lane & 31;
    const int v_first = lane - wl;                       // the warp's first vector of a batch: warp-uniform trip count
    bool seeded = !WANT_TOP2;
    for (int vb = v_first; vb < nv; vb += U * stride) {
        const int v = vb + wl;
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            raw[u] = make_uint4(Pair<T>::kNegInf2, Pair<T>::kNegInf2, Pair<T>::kNegInf2, Pair<T>::kNegInf2);
            if (v + u * stride < nv) raw[u] = ldg_stream(rv + v + u * stride);
        }
        if (!seeded) {      // (warp-uniform) 2nd largest lane maximum of the first batch
            float bm = vec_max(raw[0]);
#pragma unroll
            for (int u = 1; u < U; ++u) bm = fmaxf(bm, vec_max(raw[u]));
            const float wm = warp_max(bm);
            const unsigned holders = __ballot_sync(0xffffffffu, bm == wm);
            thr_top = warp_max((wl == __ffs(holders) - 1) ? -INFINITY : bm);
            seeded = true;
        }
        unsigned cand = 0u;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (v + u * stride < nv) consume(raw[u], 1u << u, cand);
        if (WANT_TOP2) {
            while (cand) {
                const int u = __ffs(cand) - 1;
                cand &= cand - 1;
                const int vv = v + u * stride;
                const uint4 r = __ldg(rv + vv);
                const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float x0, x1;
                    upk(Pair<T>::widen(w[i]), x0, x1);
                    top.push(x0, col0 + vv * 8 + 2 * i);
                    top.push(x1, col0 + vv * 8 + 2 * i + 1);
                }
                thr_top = fmaxf(thr_top, top.a2);
            }
        }
    }
    float a0, a1;
    upk(add2(acc0, acc1), a0, a1);
    ms.m = lane_max;                                      // the partial relative to the lane's true maximum (<= ref + slack)
    ms.s = (lane_max == -INFINITY) ? 0.f : (a0 + a1) * ex2((ref - lane_max) * k2);
}

// dispatch: the packed walk for aligned 16-bit rows, the generic one otherwise
template <typename T, bool WANT_TOP2>
__device__ __forceinline__ void shard_row_stats_any(const T* __restrict__ row, int n, float k2, MaxSum& ms, Top2& top, int col0, int lane,
                                                    int stride) {
    if constexpr (sizeof(T) == 2) {
        if (vec_ok(row, n)) {        // (uniform over the threads that share the row)
            shard_row_stats_packed<T, WANT_TOP2>(row, n, k2, ms, top, col0, lane, stride);
            return;
        }
    }
    shard_row_stats<T>(row, n, k2, ms, WANT_TOP2 ? &top : nullptr, col0, lane, stride);
}

__device__ __forceinline__ MaxSum block_merge_ms(MaxSum v, float k2, float* sm) {
    warp_merge(v, k2);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { sm[w] = v.m; sm[kVsWarps + w] = v.s; }
    __syncthreads();
    MaxSum r;
    r.init();
    if (l < kVsWarps) { r.m = sm[l]; r.s = sm[kVsWarps + l]; }
    warp_merge(r, k2);
    return r;
}

__device__ __forceinline__ Top2 block_merge_top(Top2 t, float* smf, int* smi) {
    warp_merge(t);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { smf[w] = t.a1; smf[kVsWarps + w] = t.a2; smi[w] = t.i1; smi[kVsWarps + w] = t.i2; }
    __syncthreads();
    Top2 r;
    r.init();
    if (l < kVsWarps) { r.a1 = smf[l]; r.a2 = smf[kVsWarps + l]; r.i1 = smi[l]; r.i2 = smi[kVsWarps + l]; }
    warp_merge(r);
    return r;
}

// ---- 1. local statistics ------------------------------------------------------------------------------
// G threads per row: 32 (a warp per row, shuffle-only reductions: many short rows) or the whole block (few long rows)
template <typename T, int G>
__global__ void __launch_bounds__(kVsThreads)
vshard_stats_kernel(const T* __restrict__ teacher, int64_t ld_t, const T* __restrict__ student, int64_t ld_s,
                    const int64_t* __restrict__ labels, int col0, int64_t n_rows, int V_local, float k2, float* __restrict__ stats) {
    __shared__ float smf[2 * kVsWarps];
    __shared__ int smi[2 * kVsWarps];
    constexpr int RPB = kVsThreads / G;
    const int sub = threadIdx.x / G, gl = threadIdx.x % G;
    for (int64_t r0 = (int64_t)blockIdx.x * RPB; r0 < n_rows; r0 += (int64_t)gridDim.x * RPB) {
        const int64_t r = r0 + sub;
        if (r >= n_rows) continue;      // G == 32 only (warp-uniform); RPB == 1 otherwise
        MaxSum mt, msd;
        Top2 top;
        msd.init();
        shard_row_stats_any<T, true>(teacher + r * ld_t, V_local, k2, mt, top, col0, gl, G);
        if (student) {
            Top2 unused;
            shard_row_stats_any<T, false>(student + r * ld_s, V_local, k2, msd, unused, col0, gl, G);
        }
        if constexpr (G == 32) {
            warp_merge(mt, k2);
            if (student) warp_merge(msd, k2);
            warp_merge(top);
        } else {
            mt = block_merge_ms(mt, k2, smf);
            if (student) msd = block_merge_ms(msd, k2, smf);
            top = block_merge_top(top, smf, smi);
        }
        if (gl == 0) {
            float* o = stats + r * kStatWords;
            float x_label = -INFINITY;
            if (labels) {
                const int64_t lab = labels[r] - col0;
                if (lab >= 0 && lab < V_local) x_label = Elem<T>::to_float(teacher[r * ld_t + lab]);
            }
            o[0] = mt.m; o[1] = mt.s;
            o[2] = top.a1; o[3] = __int_as_float(top.i1); o[4] = top.a2; o[5] = __int_as_float(top.i2);
            o[6] = x_label;
            if (student) { o[7] = msd.m; o[8] = msd.s; }      // else: kd_lm_head_stats_merge fills them from the GEMM's partials
            o[9] = o[10] = o[11] = 0.f;
        }
        if constexpr (G != 32) __syncthreads();
    }
}

// ---- 2. combine the G blocks of a row -----------------------------------------------------------------
__global__ void vshard_combine_kernel(const float* __restrict__ stats_all, int G, int64_t n_rows, const int64_t* __restrict__ labels,
                                      int V_total, float k2, float alpha, int variant, kd_token_record* __restrict__ rec,
                                      float2* __restrict__ lse2, float* __restrict__ rowconst, uint32_t* status) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    MaxSum mt, msd;
    mt.init();
    msd.init();
    Top2 top;
    top.init();
    float x_label = -INFINITY;
    for (int g = 0; g < G; ++g) {
        const float* o = stats_all + ((int64_t)g * n_rows + r) * kStatWords;
        mt.merge(o[0], o[1], k2);
        msd.merge(o[7], o[8], k2);
        Top2 t;
        t.a1 = o[2]; t.i1 = __float_as_int(o[3]); t.a2 = o[4]; t.i2 = __float_as_int(o[5]);
        top.merge(t);
        x_label = fmaxf(x_label, o[6]);
    }
    const float l2t = fmaf(mt.m, k2, log2f(mt.s)), l2s = fmaf(msd.m, k2, log2f(msd.s));
    lse2[r] = make_float2(l2t, l2s);
    const float NaN = __int_as_float(0x7fc00000);
    kd_token_record out;
    out.label = -1; out.v1 = NaN; out.k = top.i2; out.v2 = NaN;
    float rc = 0.f;
    const bool loca = (variant == KD_LOCA_REFERENCE || variant == KD_LOCA_PER_ROW);
    if (loca) {
        const int64_t lab = labels[r];
        if (lab < 0 || lab >= V_total) {
            atomicOr(status, KD_STATUS_BAD_LABEL);
            rc = NaN;
        } else {
            float v1, v2;
            loca_scalars(x_label, top.a2, l2t, k2, alpha, v1, v2);
            out.label = (int)lab; out.v1 = v1; out.k = (V_total > 1) ? top.i2 : 0; out.v2 = v2;
            if (variant == KD_LOCA_PER_ROW) {        // xlogy(v,v) of the row's own override cells, log2 units
                if (out.k != out.label && v1 != 0.f) rc += v1 * log2f(v1);
                if (v2 != 0.f) rc += v2 * log2f(v2);
            }
        }
    }
    rec[r] = out;
    rowconst[r] = rc;
}

// the override of local column c: the table slice (as-written mode) or the row's own cells shifted to local coordinates
template <int VARIANT>
struct ShardOverride {
    const float* table_local;
    int lab_local, k_local;
    float v1, v2;
    __device__ __forceinline__ float of(int c) const {
        const float NaN = __int_as_float(0x7fc00000);
        if constexpr (VARIANT == KD_LOCA_REFERENCE) return __ldg(table_local + c);
        if constexpr (VARIANT == KD_LOCA_PER_ROW) return (c == k_local) ? v2 : ((c == lab_local) ? v1 : NaN);
        return NaN;
    }
    // the overrides of the P cells from local column c (c % P == 0): two 128-bit loads of the table slice when it is 16-byte
    // aligned (one scalar load per CELL made the LSU, not HBM, the bound of the accumulate / emit sweeps)
    template <int P>
    __device__ __forceinline__ void of_vec(int c, float (&o)[P]) const {
        if constexpr (VARIANT == KD_LOCA_REFERENCE) {
            if ((reinterpret_cast<uintptr_t>(table_local) & 15u) == 0) {
#pragma unroll
                for (int i = 0; i < P; i += 4) {
                    const float4 t4 = __ldg(reinterpret_cast<const float4*>(table_local + c + i));
                    o[i] = t4.x; o[i + 1] = t4.y; o[i + 2] = t4.z; o[i + 3] = t4.w;
                }
                return;
            }
        }
#pragma unroll
        for (int i = 0; i < P; ++i) o[i] = of(c + i);
    }
};
// vector pairs fetched per lane before any is consumed (the streaming store is a volatile asm with a memory clobber: loads do not
// move above it, so an unrolled load-compute-store loop keeps ONE pair in flight)
template <typename T> struct VsBatch {
    static constexpr int kAccum = sizeof(T) == 2 ? KD_VS_ACCUM_BATCH : (KD_VS_ACCUM_BATCH + 1) / 2;
    static constexpr int kEmit = sizeof(T) == 2 ? KD_VS_EMIT_BATCH : (KD_VS_EMIT_BATCH + 1) / 2;
};

template <int VARIANT>
__device__ __forceinline__ ShardOverride<VARIANT> make_override(const float* table, const kd_token_record* rec, int64_t r, int col0) {
    ShardOverride<VARIANT> o;
    o.table_local = table ? table + col0 : nullptr;
    o.lab_local = o.k_local = -1;
    o.v1 = o.v2 = __int_as_float(0x7fc00000);
    if constexpr (VARIANT == KD_LOCA_PER_ROW) {
        const kd_token_record t = rec[r];
        if (t.label >= 0) {       // a bad label leaves the row without overrides (its loss is NaN through rowconst)
            o.lab_local = t.label - col0;
            o.k_local = t.k - col0;
            o.v1 = t.v1;
            o.v2 = t.v2;
        }
    }
    return o;
}

// ---- 3. local part of the row sum and of Q_r ----------------------------------------------------------
template <int VARIANT, typename T, int G>
__global__ void __launch_bounds__(kVsThreads)
vshard_accum_kernel(const T* __restrict__ teacher, int64_t ld_t, const T* __restrict__ student, int64_t ld_s,
                    const float2* __restrict__ lse2, const kd_token_record* __restrict__ rec, const float* __restrict__ table,
                    int col0, int64_t n_rows, int V_local, float k2, float thr2, float2* __restrict__ partial) {
    __shared__ float sm[2 * kVsWarps];
    constexpr bool LOCA = (VARIANT == KD_LOCA_REFERENCE || VARIANT == KD_LOCA_PER_ROW);
    constexpr int RPB = kVsThreads / G;
    const float NaN = __int_as_float(0x7fc00000);
    const int sub = threadIdx.x / G, gl = threadIdx.x % G;
    for (int64_t r0 = (int64_t)blockIdx.x * RPB; r0 < n_rows; r0 += (int64_t)gridDim.x * RPB) {
        const int64_t r = r0 + sub;
        if (r >= n_rows) continue;      // G == 32 only (warp-uniform)
        const float2 l2 = lse2[r];
        const ShardOverride<VARIANT> ov = make_override<VARIANT>(table, rec, r, col0);
        const T* trow = teacher + r * ld_t;
        const T* srow = student + r * ld_s;
        float row = 0.f, Q = 0.f;
        if (vec_ok(trow, V_local) && vec_ok(srow, V_local)) {
            constexpr int P = Elem<T>::kPack;
            const uint4* tv = reinterpret_cast<const uint4*>(trow);
            const uint4* sv = reinterpret_cast<const uint4*>(srow);
            constexpr int B = VsBatch<T>::kAccum;
            const int nv = V_local / P;
            for (int v0 = gl; v0 < nv; v0 += G * B) {
                uint4 rt[B], rs[B];
                float ovv[B][P];
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int v = v0 + u * G;
                    if (v < nv) {
                        rt[u] = ldg_stream(tv + v);
                        rs[u] = ldg_stream(sv + v);
                        if constexpr (LOCA) ov.template of_vec<P>(v * P, ovv[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int v = v0 + u * G;
                    if (v < nv) {
                        float xt[P], xs[P];
                        Elem<T>::unpack(rt[u], xt);
                        Elem<T>::unpack(rs[u], xs);
#pragma unroll
                        for (int i = 0; i < P; ++i)
                            PointTerm<VARIANT>::accum(xt[i], xs[i], LOCA ? ovv[u][i] : NaN, k2, l2.x, l2.y, thr2, row, Q);
                    }
                }
            }
        } else {
            for (int c = gl; c < V_local; c += G)
                PointTerm<VARIANT>::accum(Elem<T>::to_float(trow[c]), Elem<T>::to_float(srow[c]), LOCA ? ov.of(c) : NaN, k2, l2.x, l2.y,
                                          thr2, row, Q);
        }
        row = warp_sum(row);
        Q = warp_sum(Q);
        if constexpr (G == 32) {
            if (gl == 0) partial[r] = make_float2(row, Q);
        } else {
            const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
            __syncthreads();
            if (l == 0) { sm[w] = row; sm[kVsWarps + w] = Q; }
            __syncthreads();
            if (w == 0) {
                float a = (l < kVsWarps) ? sm[l] : 0.f, b = (l < kVsWarps) ? sm[kVsWarps + l] : 0.f;
                a = warp_sum(a);
                b = warp_sum(b);
                if (l == 0) partial[r] = make_float2(a, b);
            }
        }
    }
}

// ---- 4. gradient of the local columns -----------------------------------------------------------------
template <int VARIANT, typename T>
__global__ void __launch_bounds__(kVsThreads)
vshard_emit_kernel(const T* __restrict__ teacher, int64_t ld_t, const T* __restrict__ student, int64_t ld_s,
                   const float2* __restrict__ lse2, const float* __restrict__ Q_rows, const kd_token_record* __restrict__ rec,
                   const float* __restrict__ table, int col0, T* __restrict__ grad, int64_t ld_g, int64_t n_rows, int V_local,
                   float k2, float thr2, float gscale, const float* __restrict__ upstream) {
    constexpr bool LOCA = (VARIANT == KD_LOCA_REFERENCE || VARIANT == KD_LOCA_PER_ROW);
    const float NaN = __int_as_float(0x7fc00000);
    const float g = gscale * (upstream ? *upstream : 1.f);
    for (int64_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
        const float2 l2 = lse2[r];
        const float gQ = g * Q_rows[r];
        const ShardOverride<VARIANT> ov = make_override<VARIANT>(table, rec, r, col0);
        const T* trow = teacher + r * ld_t;
        const T* srow = student + r * ld_s;
        T* grow = grad + r * ld_g;
        if (vec_ok(trow, V_local) && vec_ok(srow, V_local) && vec_ok(grow, V_local)) {
            constexpr int P = Elem<T>::kPack;
            const uint4* tv = reinterpret_cast<const uint4*>(trow);
            const uint4* sv = reinterpret_cast<const uint4*>(srow);
            uint4* gv = reinterpret_cast<uint4*>(grow);
            constexpr int B = VsBatch<T>::kEmit;
            const int nv = V_local / P, stride = gridDim.x * kVsThreads;
            for (int v0 = blockIdx.x * kVsThreads + threadIdx.x; v0 < nv; v0 += stride * B) {
                uint4 rt[B], rs[B];
                float ovv[B][P];
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int v = v0 + u * stride;
                    if (v < nv) {
                        rt[u] = ldg_stream(tv + v);
                        rs[u] = ldg_stream(sv + v);
                        if constexpr (LOCA) ov.template of_vec<P>(v * P, ovv[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int v = v0 + u * stride;
                    if (v < nv) {
                        float xt[P], xs[P], o[P];
                        Elem<T>::unpack(rt[u], xt);
                        Elem<T>::unpack(rs[u], xs);
#pragma unroll
                        for (int i = 0; i < P; ++i)
                            o[i] = PointTerm<VARIANT>::grad(xt[i], xs[i], LOCA ? ovv[u][i] : NaN, k2, l2.x, l2.y, thr2, g, gQ);
                        stg_stream(gv + v, Elem<T>::pack(o));
                    }
                }
            }
        } else {
            for (int c = blockIdx.x * kVsThreads + threadIdx.x; c < V_local; c += gridDim.x * kVsThreads)
                grow[c] = Elem<T>::from_float(PointTerm<VARIANT>::grad(Elem<T>::to_float(trow[c]), Elem<T>::to_float(srow[c]),
                                                                       LOCA ? ov.of(c) : NaN, k2, l2.x, l2.y, thr2, g, gQ));
        }
    }
}

// warp-per-row form of the gradient kernel (many short rows): the row's scalars are loaded once per ~V_local cells and
// every lane keeps four teacher + four student loads in flight
template <int VARIANT, typename T>
__global__ void __launch_bounds__(kVsThreads)
vshard_emit_rows_kernel(const T* __restrict__ teacher, int64_t ld_t, const T* __restrict__ student, int64_t ld_s,
                        const float2* __restrict__ lse2, const float* __restrict__ Q_rows, const kd_token_record* __restrict__ rec,
                        const float* __restrict__ table, int col0, T* __restrict__ grad, int64_t ld_g, int64_t n_rows, int V_local,
                        float k2, float thr2, float gscale, const float* __restrict__ upstream) {
    constexpr bool LOCA = (VARIANT == KD_LOCA_REFERENCE || VARIANT == KD_LOCA_PER_ROW);
    const float NaN = __int_as_float(0x7fc00000);
    const float g = gscale * (upstream ? *upstream : 1.f);
    const int sub = threadIdx.x >> 5, gl = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * kVsWarps + sub; r < n_rows; r += (int64_t)gridDim.x * kVsWarps) {
        const float2 l2 = lse2[r];
        const float gQ = g * Q_rows[r];
        const ShardOverride<VARIANT> ov = make_override<VARIANT>(table, rec, r, col0);
        const T* trow = teacher + r * ld_t;
        const T* srow = student + r * ld_s;
        T* grow = grad + r * ld_g;
        if (vec_ok(trow, V_local) && vec_ok(srow, V_local) && vec_ok(grow, V_local)) {
            constexpr int P = Elem<T>::kPack;
            const uint4* tv = reinterpret_cast<const uint4*>(trow);
            const uint4* sv = reinterpret_cast<const uint4*>(srow);
            uint4* gv = reinterpret_cast<uint4*>(grow);
            constexpr int B = VsBatch<T>::kEmit;
            const int nv = V_local / P;
            for (int v0 = gl; v0 < nv; v0 += 32 * B) {
                uint4 rt[B], rs[B];
                float ovv[B][P];
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int v = v0 + u * 32;
                    if (v < nv) {
                        rt[u] = ldg_stream(tv + v);
                        rs[u] = ldg_stream(sv + v);
                        if constexpr (LOCA) ov.template of_vec<P>(v * P, ovv[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int v = v0 + u * 32;
                    if (v < nv) {
                        float xt[P], xs[P], o[P];
                        Elem<T>::unpack(rt[u], xt);
                        Elem<T>::unpack(rs[u], xs);
#pragma unroll
                        for (int i = 0; i < P; ++i)
                            o[i] = PointTerm<VARIANT>::grad(xt[i], xs[i], LOCA ? ovv[u][i] : NaN, k2, l2.x, l2.y, thr2, g, gQ);
                        stg_stream(gv + v, Elem<T>::pack(o));
                    }
                }
            }
        } else {
            for (int c = gl; c < V_local; c += 32)
                grow[c] = Elem<T>::from_float(PointTerm<VARIANT>::grad(Elem<T>::to_float(trow[c]), Elem<T>::to_float(srow[c]),
                                                                       LOCA ? ov.of(c) : NaN, k2, l2.x, l2.y, thr2, g, gQ));
        }
    }
}

// a warp per row once every resident warp has a row of its own; the whole block per row below that
bool vs_warp_rows(int64_t n_rows);

int vs_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

bool vs_warp_rows(int64_t n_rows) { return n_rows >= (int64_t)vs_sm_count() * 64; }

int vs_check(int64_t n_rows, int64_t V_local, int dtype, float T) {
    KD_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31), "n_rows out of range (%lld)", (long long)n_rows);
    KD_CHECK_ARG(V_local >= 1 && V_local < (1ll << 30), "V_local out of range (%lld)", (long long)V_local);
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32, "unknown dtype %d", dtype);
    KD_CHECK_ARG(T > 0.f && isfinite(T), "temperature must be positive and finite (got %g)", (double)T);
    return KD_OK;
}

template <int VARIANT, typename T>
int launch_accum(const void* t, int64_t ld_t, const void* s, int64_t ld_s, const float2* lse2, const kd_token_record* rec,
                 const float* table, int col0, int64_t n_rows, int V_local, float k2, float thr2, float2* partial, cudaStream_t st) {
    const int64_t cap = (int64_t)vs_sm_count() * 8;
    if (vs_warp_rows(n_rows)) {
        const int64_t blocks = (n_rows + kVsWarps - 1) / kVsWarps;
        vshard_accum_kernel<VARIANT, T, 32><<<(int)(blocks < cap ? blocks : cap), kVsThreads, 0, st>>>(
            static_cast<const T*>(t), ld_t, static_cast<const T*>(s), ld_s, lse2, rec, table, col0, n_rows, V_local, k2, thr2, partial);
    } else {
        vshard_accum_kernel<VARIANT, T, kVsThreads><<<(int)(n_rows < cap ? n_rows : cap), kVsThreads, 0, st>>>(
            static_cast<const T*>(t), ld_t, static_cast<const T*>(s), ld_s, lse2, rec, table, col0, n_rows, V_local, k2, thr2, partial);
    }
    KD_CHECK_LAUNCH("vshard_accum_kernel");
    return KD_OK;
}

template <int VARIANT, typename T>
int launch_emit(const void* t, int64_t ld_t, const void* s, int64_t ld_s, const float2* lse2, const float* Q, const kd_token_record* rec,
                const float* table, int col0, void* grad, int64_t ld_g, int64_t n_rows, int V_local, float k2, float thr2, float gscale,
                const float* upstream, cudaStream_t st) {
    if (vs_warp_rows(n_rows)) {
        const int64_t cap = (int64_t)vs_sm_count() * 8, blocks = (n_rows + kVsWarps - 1) / kVsWarps;
        vshard_emit_rows_kernel<VARIANT, T><<<(int)(blocks < cap ? blocks : cap), kVsThreads, 0, st>>>(
            static_cast<const T*>(t), ld_t, static_cast<const T*>(s), ld_s, lse2, Q, rec, table, col0, static_cast<T*>(grad), ld_g,
            n_rows, V_local, k2, thr2, gscale, upstream);
        KD_CHECK_LAUNCH("vshard_emit_rows_kernel");
        return KD_OK;
    }
    int64_t gx = (V_local + kVsThreads - 1) / kVsThreads;
    gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
    int64_t gy = (int64_t)vs_sm_count() * 16 / gx;
    gy = gy < 1 ? 1 : (gy > n_rows ? n_rows : gy);
    vshard_emit_kernel<VARIANT, T><<<dim3((unsigned)gx, (unsigned)gy), kVsThreads, 0, st>>>(
        static_cast<const T*>(t), ld_t, static_cast<const T*>(s), ld_s, lse2, Q, rec, table, col0, static_cast<T*>(grad), ld_g, n_rows,
        V_local, k2, thr2, gscale, upstream);
    KD_CHECK_LAUNCH("vshard_emit_kernel");
    return KD_OK;
}

#define KD_VS_DISPATCH(FN, ...)                                                                                  \
    do {                                                                                                         \
        switch (variant) {                                                                                       \
            case KD_LOCA_REFERENCE: return FN<KD_LOCA_REFERENCE, TT>(__VA_ARGS__);                               \
            case KD_LOCA_PER_ROW: return FN<KD_LOCA_PER_ROW, TT>(__VA_ARGS__);                                   \
            case KD_KL_PLAIN: return FN<KD_KL_PLAIN, TT>(__VA_ARGS__);                                           \
            default: return FN<KD_KL_LOG_TARGET, TT>(__VA_ARGS__);                                               \
        }                                                                                                        \
    } while (0)

template <typename TT>
int accum_t(int variant, const void* t, int64_t ld_t, const void* s, int64_t ld_s, const float2* lse2, const kd_token_record* rec,
            const float* table, int col0, int64_t n_rows, int V_local, float k2, float thr2, float2* partial, cudaStream_t st) {
    KD_VS_DISPATCH(launch_accum, t, ld_t, s, ld_s, lse2, rec, table, col0, n_rows, V_local, k2, thr2, partial, st);
}
template <typename TT>
int emit_t(int variant, const void* t, int64_t ld_t, const void* s, int64_t ld_s, const float2* lse2, const float* Q,
           const kd_token_record* rec, const float* table, int col0, void* grad, int64_t ld_g, int64_t n_rows, int V_local, float k2,
           float thr2, float gscale, const float* upstream, cudaStream_t st) {
    KD_VS_DISPATCH(launch_emit, t, ld_t, s, ld_s, lse2, Q, rec, table, col0, grad, ld_g, n_rows, V_local, k2, thr2, gscale, upstream, st);
}

}  // namespace

}  // namespace kd

using namespace kd;

extern "C" {

int kd_vshard_stats(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const int64_t* labels, int64_t col0,
                    float* stats, int64_t n_rows, int64_t V_local, int dtype, float T, void* stream) {
    if (int rc = vs_check(n_rows, V_local, dtype, T)) return rc;
    KD_CHECK_ARG(teacher && stats, "kd_vshard_stats: null pointer");
    KD_CHECK_ARG(ld_t >= V_local && (!student || ld_s >= V_local) && col0 >= 0, "kd_vshard_stats: bad shape");
    if (n_rows == 0) return KD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float k2 = kLog2e / T;
    const int64_t cap = (int64_t)vs_sm_count() * 8;
    const bool warp_rows = vs_warp_rows(n_rows);
    const int64_t blocks = warp_rows ? (n_rows + kVsWarps - 1) / kVsWarps : n_rows;
    const int grid = (int)(blocks < cap ? blocks : cap);
#define KD_VS_STATS(TT)                                                                                                           \
    do {                                                                                                                          \
        if (warp_rows)                                                                                                            \
            vshard_stats_kernel<TT, 32><<<grid, kVsThreads, 0, st>>>(static_cast<const TT*>(teacher), ld_t,                        \
                static_cast<const TT*>(student), ld_s, labels, (int)col0, n_rows, (int)V_local, k2, stats);                       \
        else                                                                                                                      \
            vshard_stats_kernel<TT, kVsThreads><<<grid, kVsThreads, 0, st>>>(static_cast<const TT*>(teacher), ld_t,                \
                static_cast<const TT*>(student), ld_s, labels, (int)col0, n_rows, (int)V_local, k2, stats);                       \
    } while (0)
    switch (dtype) {
        case KD_DTYPE_BF16: KD_VS_STATS(__nv_bfloat16); break;
        case KD_DTYPE_F16: KD_VS_STATS(__half); break;
        default: KD_VS_STATS(float);
    }
#undef KD_VS_STATS
    KD_CHECK_LAUNCH("vshard_stats_kernel");
    return KD_OK;
}

int kd_vshard_combine(const float* stats_all, int n_shards, int64_t n_rows, const int64_t* labels, int64_t V_total, float T,
                      float alpha, int variant, kd_token_record* records, float* lse2 /* [n_rows][2] */, float* rowconst,
                      uint32_t* status, void* stream) {
    KD_CHECK_ARG(stats_all && records && lse2 && rowconst && status, "kd_vshard_combine: null pointer");
    KD_CHECK_ARG(n_shards >= 1 && n_rows >= 0 && V_total >= 1 && T > 0.f, "kd_vshard_combine: bad arguments");
    KD_CHECK_ARG(variant >= KD_LOCA_REFERENCE && variant <= KD_KL_LOG_TARGET, "unknown logit-KD variant %d", variant);
    const bool loca = (variant == KD_LOCA_REFERENCE || variant == KD_LOCA_PER_ROW);
    KD_CHECK_ARG(!loca || labels, "kd_vshard_combine: LoCa variants need labels");
    if (n_rows == 0) return KD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    vshard_combine_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(stats_all, n_shards, n_rows, labels, (int)V_total, kLog2e / T,
                                                                           alpha, variant, records, reinterpret_cast<float2*>(lse2),
                                                                           rowconst, status);
    KD_CHECK_LAUNCH("vshard_combine_kernel");
    return KD_OK;
}

int kd_vshard_accum(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* lse2,
                    const kd_token_record* records, const float* table, int64_t col0, float* partial /* [n_rows][2] */,
                    int64_t n_rows, int64_t V_local, int dtype, float T, float eps, int variant, void* stream) {
    if (int rc = vs_check(n_rows, V_local, dtype, T)) return rc;
    KD_CHECK_ARG(teacher && student && lse2 && partial, "kd_vshard_accum: null pointer");
    KD_CHECK_ARG(variant >= KD_LOCA_REFERENCE && variant <= KD_KL_LOG_TARGET, "unknown logit-KD variant %d", variant);
    KD_CHECK_ARG(variant != KD_LOCA_REFERENCE || table, "kd_vshard_accum: the as-written mode needs the override table");
    KD_CHECK_ARG(variant != KD_LOCA_PER_ROW || records, "kd_vshard_accum: the per-row mode needs the token records");
    if (n_rows == 0) return KD_OK;
    const bool loca = (variant == KD_LOCA_REFERENCE || variant == KD_LOCA_PER_ROW);
    const float k2 = kLog2e / T, thr2 = loca ? log2f(eps) : -INFINITY;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float2* l2 = reinterpret_cast<const float2*>(lse2);
    float2* part = reinterpret_cast<float2*>(partial);
    switch (dtype) {
        case KD_DTYPE_BF16: return accum_t<__nv_bfloat16>(variant, teacher, ld_t, student, ld_s, l2, records, table, (int)col0, n_rows, (int)V_local, k2, thr2, part, st);
        case KD_DTYPE_F16: return accum_t<__half>(variant, teacher, ld_t, student, ld_s, l2, records, table, (int)col0, n_rows, (int)V_local, k2, thr2, part, st);
        default: return accum_t<float>(variant, teacher, ld_t, student, ld_s, l2, records, table, (int)col0, n_rows, (int)V_local, k2, thr2, part, st);
    }
}

int kd_vshard_emit(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* lse2, const float* Q_rows,
                   const kd_token_record* records, const float* table, int64_t col0, void* grad_student, int64_t ld_g,
                   int64_t n_rows, int64_t V_local, int dtype, float T, float eps, int variant, double inv_norm,
                   const float* upstream, void* stream) {
    if (int rc = vs_check(n_rows, V_local, dtype, T)) return rc;
    KD_CHECK_ARG(teacher && student && lse2 && Q_rows && grad_student, "kd_vshard_emit: null pointer");
    KD_CHECK_ARG(variant >= KD_LOCA_REFERENCE && variant <= KD_KL_LOG_TARGET, "unknown logit-KD variant %d", variant);
    KD_CHECK_ARG(variant != KD_LOCA_REFERENCE || table, "kd_vshard_emit: the as-written mode needs the override table");
    KD_CHECK_ARG(variant != KD_LOCA_PER_ROW || records, "kd_vshard_emit: the per-row mode needs the token records");
    KD_CHECK_ARG(ld_g >= V_local, "kd_vshard_emit: leading dimension < V_local");
    if (n_rows == 0) return KD_OK;
    const bool loca = (variant == KD_LOCA_REFERENCE || variant == KD_LOCA_PER_ROW);
    const float k2 = kLog2e / T, thr2 = loca ? log2f(eps) : -INFINITY;
    const float gscale = static_cast<float>(T * inv_norm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float2* l2 = reinterpret_cast<const float2*>(lse2);
    switch (dtype) {
        case KD_DTYPE_BF16: return emit_t<__nv_bfloat16>(variant, teacher, ld_t, student, ld_s, l2, Q_rows, records, table, (int)col0, grad_student, ld_g, n_rows, (int)V_local, k2, thr2, gscale, upstream, st);
        case KD_DTYPE_F16: return emit_t<__half>(variant, teacher, ld_t, student, ld_s, l2, Q_rows, records, table, (int)col0, grad_student, ld_g, n_rows, (int)V_local, k2, thr2, gscale, upstream, st);
        default: return emit_t<float>(variant, teacher, ld_t, student, ld_s, l2, Q_rows, records, table, (int)col0, grad_student, ld_g, n_rows, (int)V_local, k2, thr2, gscale, upstream, st);
    }
}

}  // extern "C"
