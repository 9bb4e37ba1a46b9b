// logit_kd.cu — logit-KD kernels, general streaming path, and the C ABI of the logit-KD family.
//
// Replaces compute_loca_loss (P1:141-194 / LG:208-261), the KL term of compute_vision_loss (P1:331-343)
// and of compute_loss (FB:206-219) of the reference.  Layout: row-major [n_rows, V] logits (bf16 / f16 / f32),
// one token row per CTA iteration.  Kernels in this file read their rows straight from global memory in three
// sweeps (stats -> accumulate -> emit); logit_kd_cluster.cu keeps the row pair resident in the distributed
// shared memory of a thread-block cluster so that HBM sees every byte once.
#include <math.h>
#include <stdlib.h>

#include "cluster_prims.cuh"
#include "logit_kd.cuh"

// vector pairs per thread and row chunk of the backward re-emission, 16-bit logits (tools/build_variant.sh --src logit_kd.cu,
// tools/knob_sweep.py at 16 384 x 151 936).  The as-written variant also fetches its slice of the override table per pair and
// is best with 2 (1 -> 3.17 ms, 2 -> 2.36, 3 -> 2.54, 4 -> 2.53, 8 -> 3.43); the table-free variants with 4 (per-row 2.47 against
// 2.57 ms with 2, plain KL 2.38 against 2.43).
#ifndef KD_STREAM_U
// vector pairs fetched per thread before any is consumed in sweeps B / C of the streaming emit kernel (as the helper beside clusters of
// 8: per-row 3.275 ms with 2, 3.268 with 3, 3.324 with 4 — flat, the helper's rows are not bound by loads in flight)
#define KD_STREAM_U 2
#endif
#ifndef KD_REGRAD_BATCH_TABLE
#define KD_REGRAD_BATCH_TABLE 2
#endif
#ifndef KD_REGRAD_BATCH
#define KD_REGRAD_BATCH 4
#endif

namespace kd {

namespace {

constexpr int kStatsThreads = 256;
constexpr int kEmitThreads = 512;
constexpr int kMaxWarps = 32;

template <typename T>
__device__ __forceinline__ bool aligned16(const T* p) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

// ------------------------------------------------------------------------------------------------
// block-wide merges (results valid in every thread)
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ MaxSum block_merge(MaxSum v, float k2, float* sm /* 2*kMaxWarps */) {
    warp_merge(v, k2);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { sm[w] = v.m; sm[kMaxWarps + w] = v.s; }
    __syncthreads();
    MaxSum r;
    r.init();
    if (l < NT / 32) { r.m = sm[l]; r.s = sm[kMaxWarps + l]; }
    warp_merge(r, k2);
    return r;
}

template <int NT>
__device__ __forceinline__ Top2 block_merge(Top2 t, float* smf /* 2*kMaxWarps */, int* smi /* 2*kMaxWarps */) {
    warp_merge(t);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { smf[w] = t.a1; smf[kMaxWarps + w] = t.a2; smi[w] = t.i1; smi[kMaxWarps + w] = t.i2; }
    __syncthreads();
    Top2 r;
    r.init();
    if (l < NT / 32) { r.a1 = smf[l]; r.a2 = smf[kMaxWarps + l]; r.i1 = smi[l]; r.i2 = smi[kMaxWarps + l]; }
    warp_merge(r);
    return r;
}

template <int NT>
__device__ __forceinline__ float2 block_sum2(float x, float y, float* sm /* 2*kMaxWarps */) {
    x = warp_sum(x);
    y = warp_sum(y);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { sm[w] = x; sm[kMaxWarps + w] = y; }
    __syncthreads();
    float a = (l < NT / 32) ? sm[l] : 0.f;
    float b = (l < NT / 32) ? sm[kMaxWarps + l] : 0.f;
    return make_float2(warp_sum(a), warp_sum(b));
}

// ------------------------------------------------------------------------------------------------
// per-thread running statistics over a strided share of one row
// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC, int NT, bool WANT_TOP2>
__device__ __forceinline__ void row_stats(const T* __restrict__ row, int V, float k2, MaxSum& ms, Top2& top) {
    constexpr int P = Elem<T>::kPack;
    ms.init();
    if (WANT_TOP2) top.init();
    float mk2 = -INFINITY;     // ms.m * k2
    float thr_top = -INFINITY; // lower bound of the row's runner-up value (warp-refreshed): smaller cells cannot matter
    int done = 0;
    if (VEC) {
        const int nvec = V / P;
        const uint4* rv = reinterpret_cast<const uint4*>(row);
        auto consume = [&](const uint4& raw, int v) {
            float x[P];
            Elem<T>::unpack(raw, x);
            float vm = x[0];
#pragma unroll
            for (int i = 1; i < P; ++i) vm = fmaxf(vm, x[i]);
            if (vm > ms.m) {
                ms.s *= ex2((ms.m - vm) * k2);
                ms.m = vm;
                mk2 = vm * k2;
            }
#pragma unroll
            for (int i = 0; i < P; ++i) ms.s += ex2(fmaf(x[i], k2, -mk2));
            if (WANT_TOP2 && vm >= thr_top) {
#pragma unroll
                for (int i = 0; i < P; ++i) top.push(x[i], v * P + i);
                thr_top = fmaxf(thr_top, top.a2);
            }
        };
        constexpr int U = 4;      // independent 128-bit loads in flight per thread
        int v = threadIdx.x;
        const int last_lane = 31 - (threadIdx.x & 31);      // warp-uniform trip count: test the warp's last lane
        for (; v + last_lane + (U - 1) * NT < nvec; v += U * NT) {
            uint4 raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) raw[u] = ldg_stream(rv + v + u * NT);
#pragma unroll
            for (int u = 0; u < U; ++u) consume(raw[u], v + u * NT);
            if (WANT_TOP2) thr_top = warp_max(thr_top);    // share the best bound across the warp
        }
        for (; v < nvec; v += NT) consume(ldg_stream(rv + v), v);
        done = nvec * P;
    }
    for (int c = done + threadIdx.x; c < V; c += NT) {
        const float x = Elem<T>::to_float(row[c]);
        if (x > ms.m) {
            ms.s *= ex2((ms.m - x) * k2);
            ms.m = x;
            mk2 = x * k2;
        }
        ms.s += ex2(fmaf(x, k2, -mk2));
        if (WANT_TOP2) top.push(x, c);
    }
}

// order-preserving float <-> int map so that a shared atomicMax can carry a float bound
__device__ __forceinline__ int float_key(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// 16-bit logits: packed min/max and fp32x2 arithmetic, lazy rescale (only when the running reference would let
// 2^((x - ref)*k2) approach overflow).  Runner-up candidates sit behind a threshold that is (i) seeded from the first
// batch (2nd largest of the warp's lane maxima) and (ii) shared across the CTA through `thr_shared` (atomicMax), so the
// divergent per-cell path runs about once per warp and row instead of once per batch.
template <typename T, int NT, bool WANT_TOP2>
__device__ __forceinline__ void row_stats_packed(const T* __restrict__ row, int V, float k2, MaxSum& ms, Top2& top, int* thr_shared) {
    ms.init();
    if (WANT_TOP2) top.init();
    const int nvec = V / 8;
    const uint4* rv = reinterpret_cast<const uint4*>(row);
    const u64 k2p = pk(k2, k2);
    const float slack = 64.f / k2;          // logits may exceed the reference by this much before a rescale
    float ref = -INFINITY, lane_max = -INFINITY, thr_top = -INFINITY;
    u64 mkp = pk(0.f, 0.f), acc0 = pk(0.f, 0.f), acc1 = pk(0.f, 0.f);
    auto vec_max = [&](const uint4& raw) -> float {
        const uint32_t m2 = Pair<T>::vmax(Pair<T>::vmax(raw.x, raw.y), Pair<T>::vmax(raw.z, raw.w));
        float m_lo, m_hi;
        upk(Pair<T>::widen(m2), m_lo, m_hi);
        return fmaxf(m_lo, m_hi);
    };
    auto consume = [&](const uint4& raw, int v) {
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
        if (WANT_TOP2 && vm >= thr_top) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float x0, x1;
                upk(Pair<T>::widen(w[i]), x0, x1);
                top.push(x0, v * 8 + 2 * i);
                top.push(x1, v * 8 + 2 * i + 1);
            }
            if (top.a2 > thr_top) {
                thr_top = top.a2;
                atomicMax(thr_shared, float_key(thr_top));
            }
        }
    };
#ifndef KD_STATS_U
#define KD_STATS_U 8
#endif
    constexpr int U = KD_STATS_U;      // independent 128-bit loads in flight per thread
    int v = threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int last_lane = 31 - lane;      // warp-uniform trip count: test the warp's last lane
    bool seeded = !WANT_TOP2;
    for (; v + last_lane + (U - 1) * NT < nvec; v += U * NT) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) raw[u] = ldg_stream(rv + v + u * NT);
        if (!seeded) {      // first batch: 2nd largest lane maximum of the warp is a valid lower bound of the runner-up
            float bm = vec_max(raw[0]);
#pragma unroll
            for (int u = 1; u < U; ++u) bm = fmaxf(bm, vec_max(raw[u]));
            const float wm = warp_max(bm);
            const unsigned holders = __ballot_sync(0xffffffffu, bm == wm);
            thr_top = warp_max((lane == __ffs(holders) - 1) ? -INFINITY : bm);
            if (lane == 0) atomicMax(thr_shared, float_key(thr_top));
            seeded = true;
        }
        if (WANT_TOP2) thr_top = fmaxf(thr_top, key_float(*reinterpret_cast<volatile int*>(thr_shared)));
#pragma unroll
        for (int u = 0; u < U; ++u) consume(raw[u], v + u * NT);
    }
    for (; v < nvec; v += NT) consume(ldg_stream(rv + v), v);
    float a0, a1;
    upk(add2(acc0, acc1), a0, a1);
    // express the partial relative to the lane's true maximum (<= ref + slack, so the factor is finite)
    ms.m = lane_max;
    ms.s = (lane_max == -INFINITY) ? 0.f : (a0 + a1) * ex2((ref - lane_max) * k2);
    for (int c = nvec * 8 + threadIdx.x; c < V; c += NT) {      // scalar tail (V % 8)
        const float x = Elem<T>::to_float(row[c]);
        const float nm = fmaxf(ms.m, x);
        ms.s = ms.s * ex2((ms.m - nm) * k2) + ex2((x - nm) * k2);
        if (ms.m == -INFINITY) ms.s = ex2((x - nm) * k2);
        ms.m = nm;
        if (WANT_TOP2) top.push(x, c);
    }
}

__device__ __forceinline__ float lse2_of(const MaxSum& ms, float k2) { return fmaf(ms.m, k2, log2f(ms.s)); }

// ------------------------------------------------------------------------------------------------
// Phase A of the LoCa reference mode: teacher row -> token record {label, v1, k, v2} and lse2_t
// ------------------------------------------------------------------------------------------------
// 4 CTAs (32 warps) per SM: 64 registers, no spills.  Measured 0.88 ms (5.65 TB/s, 86 % of the measured HBM peak) against
// 1.00 ms with the compiler's own choice of 76 registers / 3 CTAs — the kernel is latency-bound, occupancy buys the difference.
#ifndef KD_STATS_MIN_CTAS
#define KD_STATS_MIN_CTAS 4
#endif
template <typename T, bool VEC>
__global__ void __launch_bounds__(kStatsThreads, KD_STATS_MIN_CTAS)
teacher_stats_kernel(const T* __restrict__ teacher, int64_t ld_t, const int64_t* __restrict__ labels,
                     kd_token_record* __restrict__ rec, float* __restrict__ lse2_t, uint32_t* status,
                     int64_t n_rows, int V, float k2, float alpha, kd_token_record* const* __restrict__ peers, int n_peers,
                     int64_t row_offset) {
    __shared__ float smf[2 * kMaxWarps];
    __shared__ int smi[2 * kMaxWarps];
    __shared__ int thr_shared[2];        // CTA-wide runner-up bound, one slot per row parity
    if (threadIdx.x < 2) thr_shared[threadIdx.x] = float_key(-INFINITY);
    __syncthreads();
    int parity = 0;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x, parity ^= 1) {
        const T* row = teacher + r * ld_t;
        MaxSum ms;
        Top2 top;
        if (threadIdx.x == 0) thr_shared[parity ^ 1] = float_key(-INFINITY);    // next row's slot (rows are separated by the merges' barriers)
        if constexpr (VEC && sizeof(T) == 2) row_stats_packed<T, kStatsThreads, true>(row, V, k2, ms, top, thr_shared + parity);
        else row_stats<T, VEC, kStatsThreads, true>(row, V, k2, ms, top);
        ms = block_merge<kStatsThreads>(ms, k2, smf);
        top = block_merge<kStatsThreads>(top, smf, smi);
        if (threadIdx.x == 0) {
            const float l2 = lse2_of(ms, k2);
            const int64_t lab = labels[r];
            kd_token_record out;
            if (lab < 0 || lab >= V) {
                atomicOr(status, KD_STATUS_BAD_LABEL);
                out.label = -1;
                out.v1 = __int_as_float(0x7fc00000);
                out.k = top.i2;
                out.v2 = __int_as_float(0x7fc00000);
            } else {
                float v1, v2;
                loca_scalars(Elem<T>::to_float(row[lab]), top.a2, l2, k2, alpha, v1, v2);
                out.label = static_cast<int>(lab);
                out.v1 = v1;
                out.k = (V > 1) ? top.i2 : 0;
                out.v2 = v2;
            }
            if (peers) {        // fused all-gather: the record goes straight into every rank's buffer (peer stores over NVLink)
                // (visibility to the peers: kernel completion in stream order + the system-scope release of the barrier that follows)
                for (int gp = 0; gp < n_peers; ++gp) peers[gp][row_offset + r] = out;
            } else {
                rec[r] = out;
            }
            lse2_t[r] = l2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// override table (P1:184-185): last token in flattened order wins; the k pass runs after the label pass
// ------------------------------------------------------------------------------------------------
__global__ void owner_init_kernel(int* own, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) own[i] = -1;
}

__global__ void owner_scatter_kernel(const kd_token_record* __restrict__ rec, int64_t n, int* own_l, int* own_k, int V) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const kd_token_record t = rec[r];
        if (t.label >= 0 && t.label < V) atomicMax(own_l + t.label, static_cast<int>(r));
        if (t.k >= 0 && t.k < V) atomicMax(own_k + t.k, static_cast<int>(r));
    }
}

constexpr int kTableThreads = 256;

__global__ void __launch_bounds__(kTableThreads)
table_fill_kernel(const kd_token_record* __restrict__ rec, const int* __restrict__ own_l, const int* __restrict__ own_k,
                  int V, float* __restrict__ table, double* __restrict__ partials /* 2 * gridDim.x */) {
    __shared__ double sh[2][kTableThreads / 32];
    double xl = 0.0, sv = 0.0;
    const int v = blockIdx.x * kTableThreads + threadIdx.x;
    if (v < V) {
        float ov = __int_as_float(0x7fc00000);
        const int ok = own_k[v], ol = own_l[v];
        if (ok >= 0) ov = rec[ok].v2;
        else if (ol >= 0) ov = rec[ol].v1;
        table[v] = ov;
        if (ok >= 0 || ol >= 0) {
            const double d = static_cast<double>(ov);
            xl = (d == 0.0) ? 0.0 : d * log(d);
            sv = d;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xl += __shfl_xor_sync(0xffffffffu, xl, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = xl; sh[1][threadIdx.x >> 5] = sv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < kTableThreads / 32; ++i) { a += sh[0][i]; b += sh[1][i]; }
        partials[2 * blockIdx.x] = a;
        partials[2 * blockIdx.x + 1] = b;
    }
}

__global__ void table_consts_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ consts) {
    // one warp, fixed order -> deterministic
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) { a += partials[2 * i]; b += partials[2 * i + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (threadIdx.x == 0) { consts[0] = a; consts[1] = b; }
}

// ------------------------------------------------------------------------------------------------
// general streaming emit kernel: three sweeps over one row pair per CTA iteration
// ------------------------------------------------------------------------------------------------
template <int VARIANT, typename T, bool VEC>
__global__ void __launch_bounds__(kEmitThreads)
row_emit_stream_kernel(const EmitParams p) {
    constexpr int P = Elem<T>::kPack;
    constexpr int NT = kEmitThreads;
    constexpr bool LOCA = (VARIANT == KD_LOCA_REFERENCE || VARIANT == KD_LOCA_PER_ROW);
    __shared__ float smf[2 * kMaxWarps];
    __shared__ int smi[2 * kMaxWarps];

    const float NaN = __int_as_float(0x7fc00000);
    const int V = p.V;
    const float k2 = p.k2, thr2 = p.thr2;
    const float up = p.upstream ? *p.upstream : 1.f;
    const float g = p.gscale * up;
    const bool want_grad = !(p.flags & KD_FLAG_NO_GRAD);
    const int nvec = VEC ? V / P : 0;

    __shared__ int next_row_sm;
    int64_t r = blockIdx.x;
    if (p.row_queue) {          // helper mode: rows are claimed from the queue shared with the cluster kernel
        if (threadIdx.x == 0) {
            // leave the last rows to the (much faster per row) cluster pipelines: no long tail behind a late claim
            const int peek = *reinterpret_cast<volatile int*>(p.row_queue);
            next_row_sm = (peek >= p.n_rows - p.queue_tail_reserve) ? 0x7fffffff : atomicAdd(p.row_queue, 1);
        }
        __syncthreads();
        r = next_row_sm;
    }
    for (; r < p.n_rows;) {
        auto advance_row = [&]() {
            if (p.row_queue) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    const int peek = *reinterpret_cast<volatile int*>(p.row_queue);
                    next_row_sm = (peek >= p.n_rows - p.queue_tail_reserve) ? 0x7fffffff : atomicAdd(p.row_queue, 1);
                }
                __syncthreads();
                r = next_row_sm;
            } else {
                r += gridDim.x;
            }
        };
        const T* trow = static_cast<const T*>(p.teacher) + r * p.ld_t;
        const T* srow = static_cast<const T*>(p.student) + r * p.ld_s;
        T* grow = static_cast<T*>(p.grad) + r * p.ld_g;
        const uint4* tv = reinterpret_cast<const uint4*>(trow);
        const uint4* sv = reinterpret_cast<const uint4*>(srow);

        // ---- sweep A: softmax statistics ------------------------------------------------------
        MaxSum ms_s;
        Top2 top;
        row_stats<T, VEC, NT, false>(srow, V, k2, ms_s, top);
        ms_s = block_merge<NT>(ms_s, k2, smf);
        const float lse2_s = lse2_of(ms_s, k2);
        float lse2_t;
        int ov_label = -1, ov_k = -1;     // per-row override columns
        float v1 = NaN, v2 = NaN;
        if constexpr (VARIANT == KD_LOCA_REFERENCE) {
            lse2_t = p.lse2_t_in[r];
        } else {
            MaxSum ms_t;
            row_stats<T, VEC, NT, VARIANT == KD_LOCA_PER_ROW>(trow, V, k2, ms_t, top);
            ms_t = block_merge<NT>(ms_t, k2, smf);
            lse2_t = lse2_of(ms_t, k2);
            if constexpr (VARIANT == KD_LOCA_PER_ROW) {
                top = block_merge<NT>(top, smf, smi);
                const int64_t lab = p.labels[r];
                if (lab < 0 || lab >= V) {
                    if (threadIdx.x == 0) atomicOr(p.status, KD_STATUS_BAD_LABEL);
                } else {
                    ov_label = static_cast<int>(lab);
                    ov_k = (V > 1) ? top.i2 : 0;
                    loca_scalars(Elem<T>::to_float(trow[lab]), top.a2, lse2_t, k2, p.alpha, v1, v2);
                }
            }
        }

        auto override_of = [&](int c) -> float {
            if constexpr (VARIANT == KD_LOCA_REFERENCE) return __ldg(p.table + c);
            if constexpr (VARIANT == KD_LOCA_PER_ROW) return (c == ov_k) ? v2 : ((c == ov_label) ? v1 : NaN);
            return NaN;
        };

        // ---- sweep B: Q_r and the row sum ---------------------------------------------------------
        float row = 0.f, Q = 0.f;
        if (VEC) {
            auto consume = [&](const uint4& rt, const uint4& rs, int v) {
                float xt[P], xs[P];
                Elem<T>::unpack(rt, xt);
                Elem<T>::unpack(rs, xs);
#pragma unroll
                for (int i = 0; i < P; ++i)
                    PointTerm<VARIANT>::accum(xt[i], xs[i], LOCA ? override_of(v * P + i) : NaN, k2, lse2_t, lse2_s, thr2, row, Q);
            };
            constexpr int U = KD_STREAM_U;
            int v = threadIdx.x;
            for (; v + (U - 1) * NT < nvec; v += U * NT) {
                uint4 rt[U], rs[U];
#pragma unroll
                for (int u = 0; u < U; ++u) { rt[u] = ldg_keep(tv + v + u * NT); rs[u] = ldg_keep(sv + v + u * NT); }
#pragma unroll
                for (int u = 0; u < U; ++u) consume(rt[u], rs[u], v + u * NT);
            }
            for (; v < nvec; v += NT) consume(ldg_keep(tv + v), ldg_keep(sv + v), v);
        }
        for (int c = nvec * P + threadIdx.x; c < V; c += NT)
            PointTerm<VARIANT>::accum(Elem<T>::to_float(trow[c]), Elem<T>::to_float(srow[c]), LOCA ? override_of(c) : NaN,
                                      k2, lse2_t, lse2_s, thr2, row, Q);
        const float2 red = block_sum2<NT>(row, Q, smf);
        row = red.x;
        Q = red.y;
        if constexpr (VARIANT == KD_LOCA_PER_ROW) {   // xlogy(v,v) of this row's own overridden cells, in log2 units
            if (ov_label >= 0) {
                if (ov_k != ov_label && v1 != 0.f) row += v1 * log2f(v1);
                if (v2 != 0.f) row += v2 * log2f(v2);
            } else {
                row = NaN;
            }
        }
        if (threadIdx.x == 0) {
            p.rowaux[r] = make_float4(lse2_t, lse2_s, Q, row);
            if constexpr (VARIANT == KD_LOCA_PER_ROW) {
                if (p.records) {
                    kd_token_record rec;
                    rec.label = ov_label; rec.v1 = v1; rec.k = ov_k; rec.v2 = v2;
                    p.records[r] = rec;
                }
            }
        }
        if (!want_grad) { advance_row(); continue; }

        // ---- sweep C: gradient ----------------------------------------------------------------------
        const float gQ = g * Q;
        if (VEC) {
            uint4* gv = reinterpret_cast<uint4*>(grow);
            auto consume = [&](const uint4& rt, const uint4& rs, int v) {
                float xt[P], xs[P], o[P];
                Elem<T>::unpack(rt, xt);
                Elem<T>::unpack(rs, xs);
#pragma unroll
                for (int i = 0; i < P; ++i)
                    o[i] = PointTerm<VARIANT>::grad(xt[i], xs[i], LOCA ? override_of(v * P + i) : NaN, k2, lse2_t, lse2_s, thr2, g, gQ);
                stg_stream(gv + v, Elem<T>::pack(o));
            };
            constexpr int U = KD_STREAM_U;
            int v = threadIdx.x;
            for (; v + (U - 1) * NT < nvec; v += U * NT) {
                uint4 rt[U], rs[U];
#pragma unroll
                for (int u = 0; u < U; ++u) { rt[u] = ldg_stream(tv + v + u * NT); rs[u] = ldg_stream(sv + v + u * NT); }
#pragma unroll
                for (int u = 0; u < U; ++u) consume(rt[u], rs[u], v + u * NT);
            }
            for (; v < nvec; v += NT) consume(ldg_stream(tv + v), ldg_stream(sv + v), v);
        }
        for (int c = nvec * P + threadIdx.x; c < V; c += NT)
            grow[c] = Elem<T>::from_float(PointTerm<VARIANT>::grad(Elem<T>::to_float(trow[c]), Elem<T>::to_float(srow[c]),
                                                                   LOCA ? override_of(c) : NaN, k2, lse2_t, lse2_s, thr2, g, gQ));
        advance_row();
    }
}

// ------------------------------------------------------------------------------------------------
// backward re-emission: one streaming sweep that recomputes the gradient from saved row statistics
// ------------------------------------------------------------------------------------------------
template <int VARIANT, typename T, bool VEC, bool CE = false>
__global__ void __launch_bounds__(256)
regrad_kernel(const EmitParams p, const float* skip_if_one) {
    if (skip_if_one && *skip_if_one == 1.f) return;
    [[maybe_unused]] float g_ce_all = 0.f;
    if constexpr (CE) g_ce_all = (p.upstream ? *p.upstream : 1.f) * p.ce_weight * *p.ce_inv_count;
    constexpr int P = Elem<T>::kPack;
    constexpr bool LOCA = (VARIANT == KD_LOCA_REFERENCE || VARIANT == KD_LOCA_PER_ROW);
    const float NaN = __int_as_float(0x7fc00000);
    const int V = p.V;
    const float g = p.gscale * (p.upstream ? *p.upstream : 1.f);
    const int nvec = VEC ? V / P : 0;
    for (int64_t r = blockIdx.y; r < p.n_rows; r += gridDim.y) {
        const float4 aux = p.rowaux[r];
        const float lse2_t = aux.x, lse2_s = aux.y, gQ = g * aux.z;
        // fused cross-entropy: + g_ce * (2^(xs log2 e - lse1) - [c == next-token id]) on rows whose label is not ignored
        [[maybe_unused]] float g_ce = 0.f, lse1 = 0.f;
        [[maybe_unused]] int ce_c = -1;
        if constexpr (CE) {
            const int64_t c64 = p.ce_labels[r];
            if (c64 != kCeIgnore) {
                const bool ok = c64 >= 0 && c64 < (int64_t)V;
                g_ce = ok ? g_ce_all : NaN;
                ce_c = ok ? (int)c64 : -1;
                lse1 = p.ceaux[r].x;
            }
        }
        auto ce_term = [&](float xs, int c) -> float {
            if constexpr (CE) return g_ce == 0.f ? 0.f : g_ce * (ex2(fmaf(xs, kLog2e, -lse1)) - (c == ce_c ? 1.f : 0.f));
            return 0.f;
        };
        int ov_label = -1, ov_k = -1;
        float v1 = NaN, v2 = NaN;
        if constexpr (VARIANT == KD_LOCA_PER_ROW) {
            const kd_token_record rec = p.records[r];
            ov_label = rec.label; v1 = rec.v1; ov_k = rec.k; v2 = rec.v2;
        }
        auto override_of = [&](int c) -> float {
            if constexpr (VARIANT == KD_LOCA_REFERENCE) return __ldg(p.table + c);
            if constexpr (VARIANT == KD_LOCA_PER_ROW) return (c == ov_k) ? v2 : ((c == ov_label) ? v1 : NaN);
            return NaN;
        };
        const T* trow = static_cast<const T*>(p.teacher) + r * p.ld_t;
        const T* srow = static_cast<const T*>(p.student) + r * p.ld_s;
        T* grow = static_cast<T*>(p.grad) + r * p.ld_g;
        const int t0 = blockIdx.x * 256 + threadIdx.x, stride = gridDim.x * 256;
        if (VEC) {
            const uint4* tv = reinterpret_cast<const uint4*>(trow);
            const uint4* sv = reinterpret_cast<const uint4*>(srow);
            uint4* gv = reinterpret_cast<uint4*>(grow);
            // B vector pairs (and their slice of the override table, as 128-bit loads) are fetched before any is consumed: the
            // streaming store is a volatile asm with a memory clobber, so a load-compute-store loop keeps ONE pair in flight
            constexpr int B = sizeof(T) == 2 ? (VARIANT == KD_LOCA_REFERENCE ? KD_REGRAD_BATCH_TABLE : KD_REGRAD_BATCH) : 2;
            constexpr bool TABLE = (VARIANT == KD_LOCA_REFERENCE);
            const bool tab_vec = TABLE && (reinterpret_cast<uintptr_t>(p.table) & 15u) == 0;
            for (int v0 = t0; v0 < nvec; v0 += stride * B) {
                uint4 rt[B], rs[B];
                [[maybe_unused]] float ovv[B][P];
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int v = v0 + u * stride;
                    if (v < nvec) {
                        rt[u] = ldg_stream(tv + v);
                        rs[u] = ldg_stream(sv + v);
                        if constexpr (TABLE) {
                            if (tab_vec) {
#pragma unroll
                                for (int i = 0; i < P; i += 4) {
                                    const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.table + v * P + i));
                                    ovv[u][i] = t4.x; ovv[u][i + 1] = t4.y; ovv[u][i + 2] = t4.z; ovv[u][i + 3] = t4.w;
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < P; ++i) ovv[u][i] = override_of(v * P + i);
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int v = v0 + u * stride;
                    if (v < nvec) {
                        float xt[P], xs[P], o[P];
                        Elem<T>::unpack(rt[u], xt);
                        Elem<T>::unpack(rs[u], xs);
#pragma unroll
                        for (int i = 0; i < P; ++i)      // per-row LoCa: every cell as q = tp here, the row's own two cells below
                            o[i] = PointTerm<VARIANT>::grad(xt[i], xs[i], TABLE ? ovv[u][i] : NaN, p.k2, lse2_t, lse2_s, p.thr2, g, gQ) +
                                   ce_term(xs[i], v * P + i);
                        stg_stream(gv + v, Elem<T>::pack(o));
                    }
                }
            }
            if constexpr (VARIANT == KD_LOCA_PER_ROW) {
                // the row's own override cells (label -> v1, then runner-up -> v2, P1:184-185 per row): the thread that stored the
                // vector rewrites the element (same thread, program order)
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const int c = which ? ov_k : ov_label;
                    if (c < 0 || c >= nvec * P || (which == 0 && ov_k == ov_label)) continue;
                    const int v = c / P;
                    if (v >= t0 && (v - t0) % stride == 0)
                        grow[c] = Elem<T>::from_float(PointTerm<VARIANT>::grad(Elem<T>::to_float(trow[c]), Elem<T>::to_float(srow[c]),
                                                                               which ? v2 : v1, p.k2, lse2_t, lse2_s, p.thr2, g, gQ));
                }
            }
        }
        for (int c = nvec * P + t0; c < V; c += stride)
            grow[c] = Elem<T>::from_float(PointTerm<VARIANT>::grad(Elem<T>::to_float(trow[c]), Elem<T>::to_float(srow[c]),
                                                                   LOCA ? override_of(c) : NaN, p.k2, lse2_t, lse2_s, p.thr2, g, gQ) +
                                          ce_term(Elem<T>::to_float(srow[c]), c));
    }
}

// ------------------------------------------------------------------------------------------------
// loss finalisation: deterministic double sum of the per-row sums
//   loss = T^2 * inv_norm * ( unit * sum_r row[r] + n_rows * table_const )
// ------------------------------------------------------------------------------------------------
constexpr int kFinalThreads = 1024;
__global__ void __launch_bounds__(kFinalThreads)
loss_finalize_kernel(const float4* __restrict__ rowaux, int64_t n_rows, double unit, double scale,
                     const double* __restrict__ table_consts, const uint32_t* __restrict__ status, float* __restrict__ loss_out,
                     const float2* __restrict__ ceaux = nullptr, const float* __restrict__ ce_inv_count = nullptr, float ce_weight = 0.f,
                     float* __restrict__ ce_out = nullptr) {
    __shared__ double sh[kFinalThreads / 32];
    __shared__ double sh_ce[kFinalThreads / 32];
    double acc = 0.0, ce = 0.0;
    for (int64_t r = threadIdx.x; r < n_rows; r += kFinalThreads) acc += static_cast<double>(rowaux[r].w);
    if (ceaux) {      // fused cross-entropy: sum of the rows' terms (0 on ignored rows), fixed order
        for (int64_t r = threadIdx.x; r < n_rows; r += kFinalThreads) ce += static_cast<double>(ceaux[r].y);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ce += __shfl_xor_sync(0xffffffffu, ce, o);
        if ((threadIdx.x & 31) == 0) sh_ce[threadIdx.x >> 5] = ce;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < kFinalThreads / 32; ++i) tot += sh[i];
        tot *= unit;
        if (table_consts) tot += static_cast<double>(n_rows) * table_consts[0];
        double total = tot * scale;
        if (ceaux) {          // mean over the valid labels (0 valid labels -> 0 * inf = NaN, like F.cross_entropy)
            double ce_tot = 0.0;
            for (int i = 0; i < kFinalThreads / 32; ++i) ce_tot += sh_ce[i];
            ce_tot *= static_cast<double>(*ce_inv_count);
            if (ce_out) *ce_out = static_cast<float>(ce_tot);
            total += static_cast<double>(ce_weight) * ce_tot;
        }
        float out = static_cast<float>(total);
        if (status && (*status & KD_STATUS_BAD_LABEL)) {
            out = __int_as_float(0x7fc00000);
            if (ceaux && ce_out) *ce_out = out;
        }
        *loss_out = out;
    }
}

__global__ void status_clear_kernel(uint32_t* status) { *status = 0u; }
__global__ void queue_init_kernel(int* queue, int value) { *queue = value; }

// ------------------------------------------------------------------------------------------------
// in-place scaling of a gradient (autograd backward with grad_output != 1)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) scale_rows_kernel(T* g, int64_t ld, int64_t n_rows, int64_t V, const float* factor) {
    const float f = *factor;
    if (f == 1.f) return;
    constexpr int P = Elem<T>::kPack;
    for (int64_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
        T* row = g + r * ld;
        const bool vec = aligned16(row);
        const int64_t nvec = vec ? V / P : 0;
        uint4* rv = reinterpret_cast<uint4*>(row);
        for (int64_t v = blockIdx.x * 256ll + threadIdx.x; v < nvec; v += gridDim.x * 256ll) {
            float x[P];
            Elem<T>::unpack(rv[v], x);
#pragma unroll
            for (int i = 0; i < P; ++i) x[i] *= f;
            rv[v] = Elem<T>::pack(x);
        }
        for (int64_t c = nvec * P + blockIdx.x * 256ll + threadIdx.x; c < V; c += gridDim.x * 256ll)
            row[c] = Elem<T>::from_float(Elem<T>::to_float(row[c]) * f);
    }
}

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Workspace {   // layout of the caller-provided logit-KD workspace
    size_t rowaux, records, lse2_t, table_ws, table_ws_bytes, table, consts, status, queue, ceaux, total;
};

size_t table_workspace_bytes(int64_t V) {   // owner arrays (labels pass, k pass) + per-block partial sums
    const size_t blocks = (V + kTableThreads - 1) / kTableThreads;
    return align_up(sizeof(int) * V, 256) * 2 + align_up(sizeof(double) * 2 * blocks, 256);
}

Workspace workspace_layout(int64_t n_rows, int64_t V) {
    Workspace w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    w.rowaux = take(sizeof(float4) * n_rows);
    w.records = take(sizeof(kd_token_record) * n_rows);
    w.lse2_t = take(sizeof(float) * n_rows);
    w.table_ws_bytes = table_workspace_bytes(V);
    w.table_ws = take(w.table_ws_bytes);
    w.table = take(sizeof(float) * V);
    w.consts = take(sizeof(double) * 2);
    w.status = take(sizeof(uint32_t));
    w.queue = take(sizeof(int));
    w.ceaux = take(sizeof(float2) * n_rows);      // fused cross-entropy: per-row {lse at temperature 1, CE term}
    w.total = off;
    return w;
}

template <typename T>
bool rows_vectorizable(const void* base, int64_t ld) {
    return (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (ld * sizeof(T)) % 16 == 0;
}

template <typename T>
int launch_teacher_stats(const void* teacher, int64_t ld_t, const int64_t* labels, kd_token_record* rec, float* lse2_t,
                         uint32_t* status, int64_t n_rows, int V, float k2, float alpha, cudaStream_t st,
                         kd_token_record* const* peers = nullptr, int n_peers = 0, int64_t row_offset = 0) {
    // (A TMA-fed producer/consumer form of this kernel measured 1.05-1.40 ms against 1.00 ms: phase A is bound by issue slots + the
    // XU pipe at its occupancy, not by load latency; it lives in the history of this file, commit 8d7f217.)
    const int grid = static_cast<int>(n_rows < (int64_t)sm_count() * 8 ? n_rows : (int64_t)sm_count() * 8);
    if (rows_vectorizable<T>(teacher, ld_t))
        teacher_stats_kernel<T, true><<<grid, kStatsThreads, 0, st>>>(static_cast<const T*>(teacher), ld_t, labels, rec,
                                                                      lse2_t, status, n_rows, V, k2, alpha, peers, n_peers, row_offset);
    else
        teacher_stats_kernel<T, false><<<grid, kStatsThreads, 0, st>>>(static_cast<const T*>(teacher), ld_t, labels, rec,
                                                                       lse2_t, status, n_rows, V, k2, alpha, peers, n_peers, row_offset);
    KD_CHECK_LAUNCH("teacher_stats_kernel");
    return KD_OK;
}

template <int VARIANT, typename T>
int launch_emit_stream_t(const EmitParams& p, cudaStream_t st) {
    const bool vec = rows_vectorizable<T>(p.teacher, p.ld_t) && rows_vectorizable<T>(p.student, p.ld_s) &&
                     ((p.flags & KD_FLAG_NO_GRAD) || rows_vectorizable<T>(p.grad, p.ld_g));
    const int64_t cap = p.helper_ctas > 0 ? p.helper_ctas : (int64_t)sm_count() * 2;
    const int grid = static_cast<int>(p.n_rows < cap ? p.n_rows : cap);
    if (vec)
        row_emit_stream_kernel<VARIANT, T, true><<<grid, kEmitThreads, 0, st>>>(p);
    else
        row_emit_stream_kernel<VARIANT, T, false><<<grid, kEmitThreads, 0, st>>>(p);
    KD_CHECK_LAUNCH("row_emit_stream_kernel");
    return KD_OK;
}

template <typename T>
int launch_emit_stream_v(const EmitParams& p, int variant, cudaStream_t st) {
    switch (variant) {
        case KD_LOCA_REFERENCE: return launch_emit_stream_t<KD_LOCA_REFERENCE, T>(p, st);
        case KD_LOCA_PER_ROW: return launch_emit_stream_t<KD_LOCA_PER_ROW, T>(p, st);
        case KD_KL_PLAIN: return launch_emit_stream_t<KD_KL_PLAIN, T>(p, st);
        case KD_KL_LOG_TARGET: return launch_emit_stream_t<KD_KL_LOG_TARGET, T>(p, st);
    }
    set_error("unknown logit-KD variant %d", variant);
    return KD_ERR_INVALID_ARGUMENT;
}

// helper stream + events for the concurrent streaming launch on the SMs the cluster kernel cannot occupy
struct HelperStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    int device = -1;
};
HelperStream* helper_stream() {
    static HelperStream hs[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    HelperStream& h = hs[dev];
    if (h.device != dev) {
        if (cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&h.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        h.device = dev;
    }
    return &h;
}

// Tuning switches, read from the environment ONCE (first use) and never again: KD_SPARE_SMS (0 = no helper launch, 1 = force it,
// 2 = dynamic row queue without the helper), KD_HELPER_RESERVE (rows per cluster pipeline the helper leaves alone).
struct Tunables {
    int spare_sms, spare_forced, helper_reserve;
    Tunables() {
        const char* e = getenv("KD_SPARE_SMS");
        spare_sms = e ? atoi(e) : 1;
        spare_forced = e != nullptr;
        const char* r = getenv("KD_HELPER_RESERVE");
        helper_reserve = r ? atoi(r) : 10;
    }
};
const Tunables& tunables() {
    static const Tunables t;
    return t;
}
int spare_sm_mode() { return tunables().spare_sms; }

int launch_emit_stream_any(const EmitParams& p, int variant, int dtype, cudaStream_t st) {
    switch (dtype) {
        case KD_DTYPE_BF16: return launch_emit_stream_v<__nv_bfloat16>(p, variant, st);
        case KD_DTYPE_F16: return launch_emit_stream_v<__half>(p, variant, st);
        case KD_DTYPE_F32: return launch_emit_stream_v<float>(p, variant, st);
    }
    set_error("unknown dtype %d", dtype);
    return KD_ERR_INVALID_ARGUMENT;
}

// Emit = cluster-resident kernel where the shape allows it.  A cluster-of-8 launch reaches 120 of the 148 SMs, so when
// a row queue is available the general streaming kernel runs concurrently on a helper stream over the remaining SMs and
// both kernels claim rows from the same counter:
//   main stream : queue_init -> [fork] -> cluster kernel -> wait(join)
//   helper      : wait(fork) -> streaming kernel (2 CTAs per spare SM) -> [join]
int launch_emit(const EmitParams& p_in, int variant, int dtype, cudaStream_t st, int* queue) {
    EmitParams p = p_in;
    p.row_queue = nullptr;
    p.queue_tail_reserve = 0;
    p.helper_ctas = 0;
    if (p.ce_labels) {
        // the fused cross-entropy exists in the fused cluster kernel only: no streaming fallback, no helper sharing its rows
        const char* why = "the fused cross-entropy needs the fused kernel's shape: as-written LoCa, alpha in [0, 1], bf16 / fp16 logits, "
                          "V a multiple of 8 that fits the cluster, 16-byte aligned rows";
        if ((p.flags & KD_FLAG_FORCE_STREAMING) || !fused_applies(p, variant, dtype)) { set_error("%s", why); return KD_ERR_UNSUPPORTED; }
        const int rc = launch_emit_cluster(p, variant, dtype, st, nullptr, nullptr, nullptr);
        if (rc == KD_ERR_UNSUPPORTED) set_error("%s", why);
        return rc;
    }
    if (p.flags & KD_FLAG_FORCE_STREAMING) return launch_emit_stream_any(p, variant, dtype, st);

    // Per-row LoCa is the one variant where clusters of 8 (120 SMs) + the streaming helper on the other 28 SMs beat clusters of 9
    // (135 SMs, no helper): its row chain is the longest, so a helper row costs relatively less.  tools/cluster_ab.sh at 16 384 x 151 936:
    // per-row 3.44 -> 3.22 ms (T = 0.8 on peaked logits 4.33 -> 3.95); plain KL / log_target within 1 %; as-written 3.75 -> 3.87 (stays at 9).
    struct Prefer8 {
        explicit Prefer8(bool on) { fused_prefer_cluster8(on); }
        ~Prefer8() { fused_prefer_cluster8(false); }
    } prefer8(variant == KD_LOCA_PER_ROW && queue && spare_sm_mode() != 0 && p.n_rows >= 1024);
    int n_clusters = 0, cta_per = 8, pipes = 2, static_rows = 4;
    const int shape_rc = cluster_launch_shape(p, variant, dtype, &n_clusters, &cta_per, &pipes, &static_rows);
    if (shape_rc == KD_ERR_UNSUPPORTED) return launch_emit_stream_any(p, variant, dtype, st);
    if (shape_rc != KD_OK) return shape_rc;

    const int spare = sm_count() - n_clusters * cta_per;
    HelperStream* hs = nullptr;
    // The helper pays off when a fifth of the machine would otherwise idle (clusters of 8: 28 spare SMs, emit 3.07 -> 3.00 ms); with
    // clusters of 9 only 13 SMs are spare and the extra stream / event traffic costs as much as their rows bring (measured: step
    // 3.79-3.85 ms without against 3.86-3.91 ms with, CUDA-graph replay).  KD_SPARE_SMS=1 forces it on, 0 off, default: automatic.
    const int min_spare = (spare_sm_mode() == 1 && tunables().spare_forced) ? 8 : 20;
    if (queue && spare_sm_mode() && spare >= min_spare && p.n_rows >= (int64_t)n_clusters * pipes * 8) hs = helper_stream();
    if (hs) {
        queue_init_kernel<<<1, 1, 0, st>>>(queue, n_clusters * static_rows);   // the first rows of every cluster are static
        KD_CHECK_LAUNCH("queue_init_kernel");
        if (cudaEventRecord(hs->fork, st) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaEventRecord(fork)");
        p.row_queue = queue;
    }
    int rc = launch_emit_cluster(p, variant, dtype, st, nullptr, nullptr, nullptr);
    if (rc != KD_OK) return rc;
    if (hs && spare_sm_mode() != 2) {      // KD_SPARE_SMS=2: dynamic row queue without the helper kernel (experiment)
        if (cudaStreamWaitEvent(hs->stream, hs->fork, 0) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaStreamWaitEvent(fork)");
        EmitParams h = p;
        // the helper stops claiming near the end: one of its rows takes far longer than a cluster row, and a late claim would leave the
        // cluster SMs idle behind it (KD_HELPER_RESERVE = rows per cluster pipeline that stay with the cluster kernel)
        h.queue_tail_reserve = n_clusters * pipes * tunables().helper_reserve;
        h.helper_ctas = spare * 2;
        rc = launch_emit_stream_any(h, variant, dtype, hs->stream);
        if (rc != KD_OK) return rc;
        if (cudaEventRecord(hs->join, hs->stream) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaEventRecord(join)");
        if (cudaStreamWaitEvent(st, hs->join, 0) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaStreamWaitEvent(join)");
    }
    return KD_OK;
}

template <int VARIANT, typename T>
int launch_regrad_t(const EmitParams& p, const float* skip, cudaStream_t st) {
    const bool vec = rows_vectorizable<T>(p.teacher, p.ld_t) && rows_vectorizable<T>(p.student, p.ld_s) &&
                     rows_vectorizable<T>(p.grad, p.ld_g);
    // a thread takes one batch of kRegradBatch vectors of a row (the kernel fetches them all before consuming any)
    constexpr int kRegradBatch = sizeof(T) == 2 ? (VARIANT == KD_LOCA_REFERENCE ? KD_REGRAD_BATCH_TABLE : KD_REGRAD_BATCH) : 2;
    int64_t gx = (p.V / Elem<T>::kPack + 256 * kRegradBatch - 1) / (256 * kRegradBatch);
    gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
    int64_t gy = (int64_t)sm_count() * 16 / gx;
    gy = gy < 1 ? 1 : (gy > p.n_rows ? p.n_rows : gy);
    dim3 grid((unsigned)gx, (unsigned)gy);
    if constexpr (VARIANT == KD_LOCA_REFERENCE) {
        if (p.ce_labels) {
            if (vec) regrad_kernel<VARIANT, T, true, true><<<grid, 256, 0, st>>>(p, skip);
            else regrad_kernel<VARIANT, T, false, true><<<grid, 256, 0, st>>>(p, skip);
            KD_CHECK_LAUNCH("regrad_kernel");
            return KD_OK;
        }
    }
    if (vec) regrad_kernel<VARIANT, T, true><<<grid, 256, 0, st>>>(p, skip);
    else regrad_kernel<VARIANT, T, false><<<grid, 256, 0, st>>>(p, skip);
    KD_CHECK_LAUNCH("regrad_kernel");
    return KD_OK;
}

template <typename T>
int launch_regrad_v(const EmitParams& p, int variant, const float* skip, cudaStream_t st) {
    switch (variant) {
        case KD_LOCA_REFERENCE: return launch_regrad_t<KD_LOCA_REFERENCE, T>(p, skip, st);
        case KD_LOCA_PER_ROW: return launch_regrad_t<KD_LOCA_PER_ROW, T>(p, skip, st);
        case KD_KL_PLAIN: return launch_regrad_t<KD_KL_PLAIN, T>(p, skip, st);
        case KD_KL_LOG_TARGET: return launch_regrad_t<KD_KL_LOG_TARGET, T>(p, skip, st);
    }
    set_error("unknown logit-KD variant %d", variant);
    return KD_ERR_INVALID_ARGUMENT;
}

int check_common(int64_t n_rows, int64_t V, int dtype, float T) {
    KD_CHECK_ARG(n_rows >= 0, "n_rows must be >= 0 (got %lld)", (long long)n_rows);
    KD_CHECK_ARG(V >= 1 && V < (1ll << 30), "V must be in [1, 2^30) (got %lld)", (long long)V);
    KD_CHECK_ARG(n_rows < (1ll << 31), "n_rows must be < 2^31 (got %lld)", (long long)n_rows);
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32, "unknown dtype %d", dtype);
    KD_CHECK_ARG(T > 0.f && isfinite(T), "temperature must be positive and finite (got %g)", (double)T);
    return KD_OK;
}

}  // namespace

}  // namespace kd

using namespace kd;

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int kd_logit_kd_workspace_bytes(int64_t n_rows, int64_t V, size_t* bytes) {
    KD_CHECK_ARG(bytes != nullptr && n_rows >= 0 && V >= 1, "kd_logit_kd_workspace_bytes: bad arguments");
    *bytes = workspace_layout(n_rows, V).total;
    return KD_OK;
}

int kd_logit_kd_status_offset(int64_t n_rows, int64_t V, size_t* offset) {
    KD_CHECK_ARG(offset != nullptr && n_rows >= 0 && V >= 1, "kd_logit_kd_status_offset: bad arguments");
    *offset = workspace_layout(n_rows, V).status;
    return KD_OK;
}

int kd_logit_kd_rowaux_offset(int64_t n_rows, int64_t V, size_t* offset) {
    KD_CHECK_ARG(offset != nullptr && n_rows >= 0 && V >= 1, "kd_logit_kd_rowaux_offset: bad arguments");
    *offset = workspace_layout(n_rows, V).rowaux;
    return KD_OK;
}

int kd_loca_teacher_stats(const void* teacher, int64_t ld_t, const int64_t* labels, kd_token_record* records,
                          float* lse2_t, uint32_t* status, int64_t n_rows, int64_t V, int dtype, float T, float alpha,
                          void* stream) {
    if (int rc = check_common(n_rows, V, dtype, T)) return rc;
    KD_CHECK_ARG(teacher && labels && records && lse2_t && status, "kd_loca_teacher_stats: null pointer");
    KD_CHECK_ARG(ld_t >= V, "teacher leading dimension %lld < V %lld", (long long)ld_t, (long long)V);
    if (n_rows == 0) return KD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float k2 = kLog2e / T;
    switch (dtype) {
        case KD_DTYPE_BF16: return launch_teacher_stats<__nv_bfloat16>(teacher, ld_t, labels, records, lse2_t, status, n_rows, (int)V, k2, alpha, st);
        case KD_DTYPE_F16: return launch_teacher_stats<__half>(teacher, ld_t, labels, records, lse2_t, status, n_rows, (int)V, k2, alpha, st);
        default: return launch_teacher_stats<float>(teacher, ld_t, labels, records, lse2_t, status, n_rows, (int)V, k2, alpha, st);
    }
}

int kd_loca_teacher_stats_peers(const void* teacher, int64_t ld_t, const int64_t* labels, kd_token_record* const* peer_records,
                                int n_peers, int64_t row_offset, float* lse2_t, uint32_t* status, int64_t n_rows, int64_t V,
                                int dtype, float T, float alpha, void* stream) {
    if (int rc = check_common(n_rows, V, dtype, T)) return rc;
    KD_CHECK_ARG(teacher && labels && peer_records && lse2_t && status, "kd_loca_teacher_stats_peers: null pointer");
    KD_CHECK_ARG(n_peers >= 1 && row_offset >= 0, "kd_loca_teacher_stats_peers: bad peer layout");
    KD_CHECK_ARG(ld_t >= V, "teacher leading dimension %lld < V %lld", (long long)ld_t, (long long)V);
    if (n_rows == 0) return KD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float k2 = kLog2e / T;
    switch (dtype) {
        case KD_DTYPE_BF16: return launch_teacher_stats<__nv_bfloat16>(teacher, ld_t, labels, nullptr, lse2_t, status, n_rows, (int)V, k2, alpha, st, peer_records, n_peers, row_offset);
        case KD_DTYPE_F16: return launch_teacher_stats<__half>(teacher, ld_t, labels, nullptr, lse2_t, status, n_rows, (int)V, k2, alpha, st, peer_records, n_peers, row_offset);
        default: return launch_teacher_stats<float>(teacher, ld_t, labels, nullptr, lse2_t, status, n_rows, (int)V, k2, alpha, st, peer_records, n_peers, row_offset);
    }
}

int kd_loca_table_workspace_bytes(int64_t V, size_t* bytes) {
    KD_CHECK_ARG(bytes != nullptr && V >= 1, "kd_loca_table_workspace_bytes: bad arguments");
    *bytes = table_workspace_bytes(V);
    return KD_OK;
}

int kd_loca_build_table(const kd_token_record* records_all, int64_t n_tokens_all, int64_t V, float* table,
                        double* table_consts, void* workspace, size_t workspace_bytes, void* stream) {
    KD_CHECK_ARG(records_all && table && table_consts && workspace, "kd_loca_build_table: null pointer");
    KD_CHECK_ARG(n_tokens_all >= 0 && n_tokens_all < (1ll << 31), "n_tokens_all out of range");
    KD_CHECK_ARG(V >= 1 && V < (1ll << 30), "V out of range");
    size_t need = 0;
    kd_loca_table_workspace_bytes(V, &need);
    if (workspace_bytes < need) {
        set_error("kd_loca_build_table: workspace %zu < %zu bytes", workspace_bytes, need);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    int* own_l = reinterpret_cast<int*>(ws);
    int* own_k = reinterpret_cast<int*>(ws + align_up(sizeof(int) * V, 256));
    double* partials = reinterpret_cast<double*>(ws + 2 * align_up(sizeof(int) * V, 256));
    const int blocks = static_cast<int>((V + kTableThreads - 1) / kTableThreads);
    // own_l and own_k are adjacent up to padding: initialise the whole span in one launch
    const int span = static_cast<int>((reinterpret_cast<char*>(own_k) - reinterpret_cast<char*>(own_l)) / sizeof(int) + V);
    owner_init_kernel<<<sm_count() * 2, 256, 0, st>>>(own_l, span);
    KD_CHECK_LAUNCH("owner_init_kernel");
    if (n_tokens_all > 0) {
        const int g = static_cast<int>((n_tokens_all + 255) / 256);
        owner_scatter_kernel<<<g < 4096 ? g : 4096, 256, 0, st>>>(records_all, n_tokens_all, own_l, own_k, (int)V);
        KD_CHECK_LAUNCH("owner_scatter_kernel");
    }
    table_fill_kernel<<<blocks, kTableThreads, 0, st>>>(records_all, own_l, own_k, (int)V, table, partials);
    KD_CHECK_LAUNCH("table_fill_kernel");
    table_consts_kernel<<<1, 32, 0, st>>>(partials, blocks, table_consts);
    KD_CHECK_LAUNCH("table_consts_kernel");
    return KD_OK;
}

int kd_loca_emit(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* lse2_t,
                 const float* table, const double* table_consts, void* grad_student, int64_t ld_g, float* loss_out,
                 void* rowaux, int64_t n_rows, int64_t V, int dtype, float T, float eps, double inv_norm,
                 const float* upstream, const uint32_t* status, unsigned flags, void* stream) {
    if (int rc = check_common(n_rows, V, dtype, T)) return rc;
    KD_CHECK_ARG(teacher && student && lse2_t && table && table_consts && loss_out && rowaux, "kd_loca_emit: null pointer");
    KD_CHECK_ARG((flags & KD_FLAG_NO_GRAD) || grad_student, "kd_loca_emit: grad_student is null");
    KD_CHECK_ARG(ld_t >= V && ld_s >= V && ((flags & KD_FLAG_NO_GRAD) || ld_g >= V), "leading dimension < V");
    KD_CHECK_ARG(eps > 0.f, "eps must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EmitParams p{};
    p.teacher = teacher; p.ld_t = ld_t; p.student = student; p.ld_s = ld_s;
    p.lse2_t_in = lse2_t; p.table = table; p.grad = grad_student; p.ld_g = ld_g;
    p.rowaux = static_cast<float4*>(rowaux); p.upstream = upstream; p.n_rows = n_rows; p.V = (int)V;
    p.k2 = kLog2e / T; p.thr2 = log2f(eps); p.gscale = static_cast<float>(T * inv_norm); p.flags = flags;
    if (n_rows > 0) {
        if (int rc = launch_emit(p, KD_LOCA_REFERENCE, dtype, st, reinterpret_cast<int*>(p.rowaux + n_rows))) return rc;
    }
    loss_finalize_kernel<<<1, kFinalThreads, 0, st>>>(p.rowaux, n_rows, (double)kLn2, (double)T * T * inv_norm, table_consts,
                                                      status, loss_out);
    KD_CHECK_LAUNCH("loss_finalize_kernel");
    return KD_OK;
}

// kd_loca_emit + the student's cross-entropy of the LOCAL rows (the token-sharded form of kd_logit_kd_ce_fwd_bwd)
int kd_loca_emit_ce(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* lse2_t,
                    const float* table, const double* table_consts, const int64_t* ce_labels, const float* ce_inv_count,
                    float ce_weight, void* grad_student, int64_t ld_g, float* loss_out, float* ce_loss_out, void* rowaux,
                    void* ceaux, int64_t n_rows, int64_t V, int dtype, float T, float eps, double inv_norm,
                    const float* upstream, uint32_t* status, unsigned flags, void* stream) {
    if (int rc = check_common(n_rows, V, dtype, T)) return rc;
    KD_CHECK_ARG(teacher && student && lse2_t && table && table_consts && loss_out && rowaux, "kd_loca_emit_ce: null pointer");
    KD_CHECK_ARG(ce_labels && ce_inv_count && ceaux && status, "kd_loca_emit_ce: null pointer (ce_labels / ce_inv_count / ceaux / status)");
    KD_CHECK_ARG((flags & KD_FLAG_NO_GRAD) || grad_student, "kd_loca_emit_ce: grad_student is null");
    KD_CHECK_ARG(ld_t >= V && ld_s >= V && ((flags & KD_FLAG_NO_GRAD) || ld_g >= V), "leading dimension < V");
    KD_CHECK_ARG(eps > 0.f, "eps must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EmitParams p{};
    p.teacher = teacher; p.ld_t = ld_t; p.student = student; p.ld_s = ld_s;
    p.lse2_t_in = lse2_t; p.table = table; p.grad = grad_student; p.ld_g = ld_g;
    p.rowaux = static_cast<float4*>(rowaux); p.upstream = upstream; p.n_rows = n_rows; p.V = (int)V;
    p.k2 = kLog2e / T; p.thr2 = log2f(eps); p.gscale = static_cast<float>(T * inv_norm); p.flags = flags;
    p.status = status;
    p.ce_labels = ce_labels; p.ce_inv_count = ce_inv_count; p.ce_weight = ce_weight; p.ceaux = static_cast<float2*>(ceaux);
    if (n_rows > 0) {
        if (int rc = launch_emit(p, KD_LOCA_REFERENCE, dtype, st, nullptr)) return rc;
    }
    loss_finalize_kernel<<<1, kFinalThreads, 0, st>>>(p.rowaux, n_rows, (double)kLn2, (double)T * T * inv_norm, table_consts,
                                                      status, loss_out, p.ceaux, ce_inv_count, ce_weight, ce_loss_out);
    KD_CHECK_LAUNCH("loss_finalize_kernel");
    return KD_OK;
}

static int logit_kd_fwd_bwd_impl(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const int64_t* labels,
                                 void* grad_student, int64_t ld_g, float* loss_out, void* workspace, size_t workspace_bytes,
                                 int64_t n_rows, int64_t V, int dtype, float T, float alpha, float eps, int variant,
                                 double inv_norm, const float* upstream, unsigned flags, void* stream,
                                 const int64_t* ce_labels, const float* ce_inv_count, float ce_weight, float* ce_loss_out) {
    if (int rc = check_common(n_rows, V, dtype, T)) return rc;
    KD_CHECK_ARG(variant >= KD_LOCA_REFERENCE && variant <= KD_KL_LOG_TARGET, "unknown logit-KD variant %d", variant);
    const bool loca = (variant == KD_LOCA_REFERENCE || variant == KD_LOCA_PER_ROW);
    KD_CHECK_ARG(teacher && student && loss_out && workspace, "kd_logit_kd_fwd_bwd: null pointer");
    KD_CHECK_ARG(!loca || labels, "kd_logit_kd_fwd_bwd: LoCa variants need labels");
    KD_CHECK_ARG((flags & KD_FLAG_NO_GRAD) || grad_student, "kd_logit_kd_fwd_bwd: grad_student is null");
    KD_CHECK_ARG(ld_t >= V && ld_s >= V && ((flags & KD_FLAG_NO_GRAD) || ld_g >= V), "leading dimension < V");
    KD_CHECK_ARG(!loca || eps > 0.f, "eps must be positive");
    const Workspace w = workspace_layout(n_rows, V);
    if (workspace_bytes < w.total) {
        set_error("kd_logit_kd_fwd_bwd: workspace %zu < %zu bytes", workspace_bytes, w.total);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    uint32_t* status = reinterpret_cast<uint32_t*>(ws + w.status);
    float4* rowaux = reinterpret_cast<float4*>(ws + w.rowaux);
    double* consts = reinterpret_cast<double*>(ws + w.consts);
    status_clear_kernel<<<1, 1, 0, st>>>(status);
    KD_CHECK_LAUNCH("status_clear_kernel");

    EmitParams p{};
    p.teacher = teacher; p.ld_t = ld_t; p.student = student; p.ld_s = ld_s; p.labels = labels;
    p.grad = grad_student; p.ld_g = ld_g; p.rowaux = rowaux; p.status = status; p.upstream = upstream;
    p.n_rows = n_rows; p.V = (int)V; p.k2 = kLog2e / T; p.alpha = alpha; p.thr2 = loca ? log2f(eps) : -INFINITY;
    p.gscale = static_cast<float>(T * inv_norm); p.flags = flags;
    p.records = reinterpret_cast<kd_token_record*>(ws + w.records);
    float2* ceaux = nullptr;
    if (ce_labels) {
        KD_CHECK_ARG(variant == KD_LOCA_REFERENCE, "the fused cross-entropy rides the as-written LoCa mode (variant 0)");
        KD_CHECK_ARG(ce_inv_count != nullptr, "kd_logit_kd_ce_fwd_bwd: ce_inv_count is null");
        ceaux = reinterpret_cast<float2*>(ws + w.ceaux);
        p.ce_labels = ce_labels; p.ce_inv_count = ce_inv_count; p.ce_weight = ce_weight; p.ceaux = ceaux;
    }

    // the as-written mode needs every row's record before any row's q is known: phase A + table build, then the emit.
    // Per-row LoCa derives its teacher statistics inside the emit kernel's single sweep (6V bytes per token).
    if (variant == KD_LOCA_REFERENCE) {
        kd_token_record* rec = reinterpret_cast<kd_token_record*>(ws + w.records);
        float* lse2_t = reinterpret_cast<float*>(ws + w.lse2_t);
        if (n_rows > 0) {
            if (int rc = kd_loca_teacher_stats(teacher, ld_t, labels, rec, lse2_t, status, n_rows, V, dtype, T, alpha, stream)) return rc;
        }
        p.lse2_t_in = lse2_t;
        float* table = reinterpret_cast<float*>(ws + w.table);
        if (int rc = kd_loca_build_table(rec, n_rows, V, table, consts, ws + w.table_ws, w.table_ws_bytes, stream)) return rc;
        p.table = table;
    }
    if (n_rows > 0) {
        if (int rc = launch_emit(p, variant, dtype, st, reinterpret_cast<int*>(ws + w.queue))) return rc;
    }
    const double unit = (variant == KD_KL_LOG_TARGET) ? 1.0 : (double)kLn2;
    loss_finalize_kernel<<<1, kFinalThreads, 0, st>>>(rowaux, n_rows, unit, (double)T * T * inv_norm,
                                                      variant == KD_LOCA_REFERENCE ? consts : nullptr, status, loss_out,
                                                      ceaux, ce_inv_count, ce_weight, ce_loss_out);
    KD_CHECK_LAUNCH("loss_finalize_kernel");
    return KD_OK;
}

int kd_logit_kd_fwd_bwd(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const int64_t* labels,
                        void* grad_student, int64_t ld_g, float* loss_out, void* workspace, size_t workspace_bytes,
                        int64_t n_rows, int64_t V, int dtype, float T, float alpha, float eps, int variant,
                        double inv_norm, const float* upstream, unsigned flags, void* stream) {
    return logit_kd_fwd_bwd_impl(teacher, ld_t, student, ld_s, labels, grad_student, ld_g, loss_out, workspace, workspace_bytes, n_rows, V,
                                 dtype, T, alpha, eps, variant, inv_norm, upstream, flags, stream, nullptr, nullptr, 0.f, nullptr);
}

int kd_logit_kd_ce_fwd_bwd(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const int64_t* labels,
                           const int64_t* ce_labels, const float* ce_inv_count, float ce_weight, void* grad_student, int64_t ld_g,
                           float* loss_out, float* ce_loss_out, void* workspace, size_t workspace_bytes, int64_t n_rows, int64_t V,
                           int dtype, float T, float alpha, float eps, double inv_norm, const float* upstream, unsigned flags,
                           void* stream) {
    KD_CHECK_ARG(ce_labels && ce_inv_count, "kd_logit_kd_ce_fwd_bwd: null pointer");
    return logit_kd_fwd_bwd_impl(teacher, ld_t, student, ld_s, labels, grad_student, ld_g, loss_out, workspace, workspace_bytes, n_rows, V,
                                 dtype, T, alpha, eps, KD_LOCA_REFERENCE, inv_norm, upstream, flags, stream, ce_labels, ce_inv_count,
                                 ce_weight, ce_loss_out);
}

static int logit_kd_regrad_impl(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, void* grad_student,
                                int64_t ld_g, const void* workspace, size_t workspace_bytes, int64_t n_rows, int64_t V, int dtype,
                                float T, float eps, int variant, double inv_norm, const float* upstream, const float* skip_if_one,
                                void* stream, const int64_t* ce_labels, const float* ce_inv_count, float ce_weight) {
    if (int rc = check_common(n_rows, V, dtype, T)) return rc;
    KD_CHECK_ARG(variant >= KD_LOCA_REFERENCE && variant <= KD_KL_LOG_TARGET, "unknown logit-KD variant %d", variant);
    KD_CHECK_ARG(teacher && student && grad_student && workspace, "kd_logit_kd_regrad: null pointer");
    KD_CHECK_ARG(ld_t >= V && ld_s >= V && ld_g >= V, "leading dimension < V");
    const Workspace w = workspace_layout(n_rows, V);
    if (workspace_bytes < w.total) {
        set_error("kd_logit_kd_regrad: workspace %zu < %zu bytes", workspace_bytes, w.total);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    if (n_rows == 0) return KD_OK;
    const bool loca = (variant == KD_LOCA_REFERENCE || variant == KD_LOCA_PER_ROW);
    char* ws = const_cast<char*>(static_cast<const char*>(workspace));
    EmitParams p{};
    p.teacher = teacher; p.ld_t = ld_t; p.student = student; p.ld_s = ld_s; p.grad = grad_student; p.ld_g = ld_g;
    p.rowaux = reinterpret_cast<float4*>(ws + w.rowaux);
    p.records = reinterpret_cast<kd_token_record*>(ws + w.records);
    p.table = reinterpret_cast<const float*>(ws + w.table);
    p.upstream = upstream; p.n_rows = n_rows; p.V = (int)V; p.k2 = kLog2e / T;
    p.thr2 = loca ? log2f(eps) : -INFINITY; p.gscale = static_cast<float>(T * inv_norm);
    if (ce_labels) {
        p.ce_labels = ce_labels; p.ce_inv_count = ce_inv_count; p.ce_weight = ce_weight;
        p.ceaux = reinterpret_cast<float2*>(ws + w.ceaux);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case KD_DTYPE_BF16: return launch_regrad_v<__nv_bfloat16>(p, variant, skip_if_one, st);
        case KD_DTYPE_F16: return launch_regrad_v<__half>(p, variant, skip_if_one, st);
        default: return launch_regrad_v<float>(p, variant, skip_if_one, st);
    }
}

// re-emission for the phase-split path: the same sweep as kd_logit_kd_regrad, fed by the buffers kd_loca_emit(_ce) left
int kd_loca_regrad(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* table, const void* rowaux,
                   const int64_t* ce_labels, const float* ce_inv_count, float ce_weight, const void* ceaux, void* grad_student,
                   int64_t ld_g, int64_t n_rows, int64_t V, int dtype, float T, float eps, double inv_norm, const float* upstream,
                   const float* skip_if_one, void* stream) {
    if (int rc = check_common(n_rows, V, dtype, T)) return rc;
    KD_CHECK_ARG(teacher && student && table && rowaux && grad_student, "kd_loca_regrad: null pointer");
    KD_CHECK_ARG(!ce_labels || (ce_inv_count && ceaux), "kd_loca_regrad: ce_labels without ce_inv_count / ceaux");
    KD_CHECK_ARG(ld_t >= V && ld_s >= V && ld_g >= V, "leading dimension < V");
    KD_CHECK_ARG(eps > 0.f, "eps must be positive");
    if (n_rows == 0) return KD_OK;
    EmitParams p{};
    p.teacher = teacher; p.ld_t = ld_t; p.student = student; p.ld_s = ld_s; p.grad = grad_student; p.ld_g = ld_g;
    p.rowaux = static_cast<float4*>(const_cast<void*>(rowaux));
    p.table = table; p.upstream = upstream; p.n_rows = n_rows; p.V = (int)V; p.k2 = kLog2e / T;
    p.thr2 = log2f(eps); p.gscale = static_cast<float>(T * inv_norm);
    if (ce_labels) {
        p.ce_labels = ce_labels; p.ce_inv_count = ce_inv_count; p.ce_weight = ce_weight;
        p.ceaux = static_cast<float2*>(const_cast<void*>(ceaux));
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case KD_DTYPE_BF16: return launch_regrad_v<__nv_bfloat16>(p, KD_LOCA_REFERENCE, skip_if_one, st);
        case KD_DTYPE_F16: return launch_regrad_v<__half>(p, KD_LOCA_REFERENCE, skip_if_one, st);
        default: return launch_regrad_v<float>(p, KD_LOCA_REFERENCE, skip_if_one, st);
    }
}

int kd_logit_kd_regrad(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, void* grad_student,
                       int64_t ld_g, const void* workspace, size_t workspace_bytes, int64_t n_rows, int64_t V, int dtype,
                       float T, float eps, int variant, double inv_norm, const float* upstream, const float* skip_if_one,
                       void* stream) {
    return logit_kd_regrad_impl(teacher, ld_t, student, ld_s, grad_student, ld_g, workspace, workspace_bytes, n_rows, V, dtype, T, eps,
                                variant, inv_norm, upstream, skip_if_one, stream, nullptr, nullptr, 0.f);
}

int kd_logit_kd_ce_regrad(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const int64_t* ce_labels,
                          const float* ce_inv_count, float ce_weight, void* grad_student, int64_t ld_g, const void* workspace,
                          size_t workspace_bytes, int64_t n_rows, int64_t V, int dtype, float T, float eps, double inv_norm,
                          const float* upstream, const float* skip_if_one, void* stream) {
    KD_CHECK_ARG(ce_labels && ce_inv_count, "kd_logit_kd_ce_regrad: null pointer");
    return logit_kd_regrad_impl(teacher, ld_t, student, ld_s, grad_student, ld_g, workspace, workspace_bytes, n_rows, V, dtype, T, eps,
                                KD_LOCA_REFERENCE, inv_norm, upstream, skip_if_one, stream, ce_labels, ce_inv_count, ce_weight);
}

int kd_scale_rows(void* grad, int64_t ld_g, int64_t n_rows, int64_t V, int dtype, const float* factor, void* stream) {
    KD_CHECK_ARG(grad && factor, "kd_scale_rows: null pointer");
    KD_CHECK_ARG(n_rows >= 0 && V >= 1 && ld_g >= V, "kd_scale_rows: bad shape");
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32, "unknown dtype %d", dtype);
    if (n_rows == 0) return KD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t gx = (V / 8 + 255) / 256;
    gx = gx < 1 ? 1 : (gx > 1024 ? 1024 : gx);
    int64_t gy = 8192 / gx;
    gy = gy < 1 ? 1 : (gy > n_rows ? n_rows : gy);
    dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
    switch (dtype) {
        case KD_DTYPE_BF16: scale_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<__nv_bfloat16*>(grad), ld_g, n_rows, V, factor); break;
        case KD_DTYPE_F16: scale_rows_kernel<__half><<<grid, 256, 0, st>>>(static_cast<__half*>(grad), ld_g, n_rows, V, factor); break;
        default: scale_rows_kernel<float><<<grid, 256, 0, st>>>(static_cast<float*>(grad), ld_g, n_rows, V, factor); break;
    }
    KD_CHECK_LAUNCH("scale_rows_kernel");
    return KD_OK;
}

}  // extern "C"
