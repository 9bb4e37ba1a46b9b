// logit_kd_cluster.cu — cluster-resident logit-KD emit kernel for sm_100a.
//
// One token row pair (teacher + student, 2 x 2V bytes) is held in the DISTRIBUTED SHARED MEMORY of a
// thread-block cluster of CS CTAs: CTA j keeps columns [j*V/CS, (j+1)*V/CS) of both rows, fetched from HBM
// exactly once with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx).  The three sweeps the math
// needs (softmax statistics -> Q_r / row loss -> gradient) then run out of shared memory, and the gradient is
// the only thing written back: HBM sees 6V bytes per token.
//
//   * every warp owns a private slice of its CTA's columns, its own mbarrier and its own TMA refill, so the
//     only cross-warp traffic is the two per-row reductions;
//   * reductions are pushed peer-to-peer: each warp stores its partial into a slot of EVERY CTA of the cluster
//     (st.shared::cluster) and arrives on that CTA's mbarrier; nobody executes a cluster-wide barrier per row;
//   * a CTA runs NPIPE independent row pipelines (disjoint warps, different rows) so that one pipeline's
//     reduction / TMA wait is covered by the other's arithmetic;
//   * LoCa reference mode keeps its slice of the V-entry override table in shared memory for the whole kernel.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "cluster_prims.cuh"
#include "logit_kd.cuh"

namespace kd {

namespace {

#ifndef KD_SWEEP_UNROLL
#define KD_SWEEP_UNROLL 2
#endif
constexpr int kSweepUnroll = KD_SWEEP_UNROLL;
#ifndef KD_CLUSTER_THREADS
#define KD_CLUSTER_THREADS 512
#endif
constexpr int kThreads = KD_CLUSTER_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr float kPush = -3.0e38f;       // ov' * kPush drives log2 q of an overridden cell to -inf
constexpr float kOvFloor = 1.0e-36f;    // encoded "overridden with ~0" (must stay > 0)

// Optional phase tracing (compile with -DKD_TRACE): lane 0 of every warp stamps %globaltimer at the phase
// boundaries of each row into `trace` ([cta][warp][row_iter][8] u64), read back by tools/trace_emit.py.
#ifdef KD_TRACE
#define KD_STAMP(i) do { if (lane == 0 && cp.trace && row_iter < 64) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
    cp.trace[(((size_t)blockIdx.x * kWarps + warp) * 64 + row_iter) * 8 + (i)] = t_; } } while (0)
#else
#define KD_STAMP(i) do {} while (0)
#endif

struct ClusterParams {
    unsigned long long* trace;
    EmitParams e;
    int n_clusters;
    int vpc;                      // 16-byte vectors per CTA slice
    int nvec;                     // 16-byte vectors per row
};

template <int VARIANT>
struct SlotWords { static constexpr int A = (VARIANT == KD_LOCA_REFERENCE) ? 2 : 8; };

// Shared-memory plan (all dynamic, offsets in bytes).  TSM = teacher slice resident in shared memory too; otherwise
// only the student slice is resident and the teacher slice is re-read from L2 (it is prefetched into L2 one row ahead).
template <int VARIANT, int CS, int NPIPE, bool TSM>
struct SmemPlan {
    static constexpr int WPP = kWarps / NPIPE;             // warps per pipeline
    static constexpr int NSLOT = CS * WPP;                  // partials per reduction
    static constexpr int kRowTensors = TSM ? 2 : 1;
    int table, rows, bars, slotsA, slotsB, bcast, total;
    __host__ __device__ SmemPlan(int vpc) {
        int off = 0;
        table = off;
        if (VARIANT == KD_LOCA_REFERENCE) off += vpc * 32;
        rows = off;
        off += NPIPE * kRowTensors * vpc * 16;
        bars = off;
        off += (kWarps + 2 * NPIPE) * 8;
        slotsA = off;
        off += NPIPE * NSLOT * SlotWords<VARIANT>::A * 4;
        slotsB = off;
        off += NPIPE * (NSLOT + 1) * 8;      // +1: the leader's "next row" word rides on reduction B
        bcast = off;                 // per pipeline: 8 words of reduction-A results + 2 words of reduction-B results (+pad)
        off += NPIPE * 64;
        total = off;
    }
};

// Walk the warp's vectors v = lane, lane+32, ... keeping D loads of the primary stream in flight ahead of use.
template <int D, typename LoadF, typename BodyF>
__device__ __forceinline__ void pipelined_sweep(int lane, int w_n, LoadF load, BodyF body) {
    if constexpr (D == 0) {       // shared-memory source: a plain unrolled loop schedules best
#pragma unroll kSweepUnroll
        for (int v = lane; v < w_n; v += 32) body(v, load(v));
        return;
    }
    uint4 ring[D > 0 ? D : 1];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const int v = lane + 32 * j;
        ring[j] = make_uint4(0, 0, 0, 0);
        if (v < w_n) ring[j] = load(v);
    }
    for (int v0 = lane; v0 < w_n; v0 += 32 * D) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const int v = v0 + 32 * j;
            if (v < w_n) {
                const uint4 cur = ring[j];
                if (v + 32 * D < w_n) ring[j] = load(v + 32 * D);
                body(v, cur);
            }
        }
    }
}

__device__ __forceinline__ uint4 ldg_l2(const uint4* p) {     // teacher slice: L2 resident, never worth an L1 line
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// Non-volatile shared load: the compiler may schedule it early inside a sweep.  `epoch` (the row counter) is a fake
// input so that loads of the same address in different rows are distinct values (the slice is refilled by TMA).
__device__ __forceinline__ uint4 lds128_nv(uint32_t addr, uint32_t epoch) {
    uint4 v;
#ifdef KD_ABLATE_LDS       // timing experiment only (wrong results): take shared-memory reads out of the picture
    v = make_uint4(addr, epoch, addr ^ 0x3f80, 0x3f803f80u);
    return v;
#endif
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];  // epoch %5" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "r"(epoch));
    return v;
}

// --------------------------------------------------------------------------------------------------------
template <int VARIANT, typename T, int CS, int NPIPE, bool TSM, bool DYN>
__global__ void __launch_bounds__(kThreads, 1)
row_emit_cluster_kernel(const ClusterParams cp) {
    constexpr bool LOCA = (VARIANT == KD_LOCA_REFERENCE || VARIANT == KD_LOCA_PER_ROW);
    constexpr bool TEACHER_STATS = (VARIANT != KD_LOCA_REFERENCE);
    constexpr int WPP = kWarps / NPIPE;
    constexpr int NSLOT = CS * WPP;
    constexpr int AW = SlotWords<VARIANT>::A;
    constexpr int DT = TSM ? 0 : 3;        // teacher loads kept in flight per lane (0: plain loop over shared memory)
    extern __shared__ __align__(128) unsigned char smem[];
    const EmitParams& p = cp.e;
    const SmemPlan<VARIANT, CS, NPIPE, TSM> plan(cp.vpc);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pipe = warp / WPP, pw = warp % WPP;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x / CS;
    const float NaN = __int_as_float(0x7fc00000);

    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_table = s_base + plan.table;
    const uint32_t bar_full = s_base + plan.bars + warp * 8;
    const uint32_t bar_redA = s_base + plan.bars + (kWarps + 2 * pipe) * 8;
    const uint32_t bar_redB = bar_redA + 8;
    const uint32_t slotA = s_base + plan.slotsA + pipe * NSLOT * AW * 4;
    const uint32_t slotB = s_base + plan.slotsB + pipe * (NSLOT + 1) * 8;
    constexpr bool dynamic_rows = DYN;       // rows claimed from p.row_queue (shared with the streaming helper kernel)
    const uint32_t redB_bytes = NSLOT * 8 + (dynamic_rows ? 8 : 0);
    volatile float* bcast = reinterpret_cast<volatile float*>(smem + plan.bcast + pipe * 64);
    const int barA = 1 + 2 * pipe, barB = 2 + 2 * pipe;     // named barriers of this pipeline (0 is __syncthreads)

    // this CTA's column slice and this warp's private part of it (in 16-byte vectors)
    const int cta_v0 = min((int)rank * cp.vpc, cp.nvec);
    const int cta_v1 = min(cta_v0 + cp.vpc, cp.nvec);
    const int cta_n = cta_v1 - cta_v0;
    // regions are multiples of 32 vectors: only the last non-empty warp of a slice ever runs a partial iteration
    const int w_chunk = ((cta_n + WPP * 32 - 1) / (WPP * 32)) * 32;
    const int w_lo = min(pw * w_chunk, cta_n);
    const int w_hi = min(w_lo + w_chunk, cta_n);
    const int w_n = w_hi - w_lo;
    const uint32_t s_s = s_base + plan.rows + pipe * SmemPlan<VARIANT, CS, NPIPE, TSM>::kRowTensors * cp.vpc * 16 + w_lo * 16;
    const uint32_t s_t = s_s + cp.vpc * 16;          // teacher part (TSM only)
    const uint32_t s_ov = s_table + w_lo * 16;           // cells 0-3; cells 4-7 live cp.vpc*16 bytes further
    const int g_v0 = cta_v0 + w_lo;                  // first global vector of the warp

    if (tid == 0) {
        for (int w = 0; w < kWarps; ++w) mbar_init(s_base + plan.bars + w * 8, 1);
        for (int i = 0; i < 2 * NPIPE; ++i) mbar_init(s_base + plan.bars + (kWarps + i) * 8, 1);
        fence_mbar_init();
        for (int i = 0; i < NPIPE; ++i) {        // arm the first phase of every reduction barrier
            mbar_expect_tx(s_base + plan.bars + (kWarps + 2 * i) * 8, NSLOT * AW * 4);
            mbar_expect_tx(s_base + plan.bars + (kWarps + 2 * i + 1) * 8, NSLOT * 8 + (DYN ? 8 : 0));
        }
    }
    // LoCa reference mode: encode this CTA's table slice once (0 = not overridden, else max(ov, floor) > 0)
    bool fast_table = true;
    if constexpr (VARIANT == KD_LOCA_REFERENCE) {
        // layout: [vpc][4] floats for cells 0-3 of every vector, then [vpc][4] for cells 4-7, so that both 128-bit
        // reads of a sweep have a lane stride of 16 bytes (a [vpc][8] layout costs a 2-way bank conflict per read)
        float* tbl = reinterpret_cast<float*>(smem + plan.table);
        bool ok = true;
        for (int i = tid; i < cta_n * 8; i += kThreads) {
            const float ov = p.table[cta_v0 * 8 + i];
            float enc = 0.f;
            if (ov == ov) {
                ok = ok && (ov >= 0.f) && (ov <= 3.0e38f);
                enc = fmaxf(ov, kOvFloor);
            }
            const int vec = i >> 3, cell = i & 7;
            tbl[(cell >> 2) * cp.vpc * 4 + vec * 4 + (cell & 3)] = enc;
        }
        // a negative / non-finite override in this slice (alpha >= 1) forces the exact per-cell path for this CTA
        fast_table = __syncthreads_and(ok);
    } else {
        __syncthreads();
    }
    cluster_sync_all();     // barriers initialised cluster-wide before any remote push

    const int row_stride = cp.n_clusters * NPIPE;
    int64_t r = (int64_t)cluster_id * NPIPE + pipe;
    const T* tbase = static_cast<const T*>(p.teacher);
    const T* sbase = static_cast<const T*>(p.student);
    auto issue_load = [&](int64_t row) {
        if (lane == 0 && w_n > 0) {
            const uint32_t bytes = (uint32_t)w_n * 16u;
            fence_proxy_async();
            mbar_expect_tx(bar_full, (TSM ? 2u : 1u) * bytes);
            tma_load_1d(s_s, sbase + row * p.ld_s + (int64_t)g_v0 * 8, bytes, bar_full);
            if constexpr (TSM) tma_load_1d(s_t, tbase + row * p.ld_t + (int64_t)g_v0 * 8, bytes, bar_full);
        }
    };
    auto warm_l2 = [&](int64_t row) {           // bring the warp's slice of `row` into L2 ahead of its use
        if (lane == 0 && w_n > 0 && row < p.n_rows) {
            prefetch_l2(tbase + row * p.ld_t + (int64_t)g_v0 * 8, (uint32_t)w_n * 16u);
            prefetch_l2(sbase + row * p.ld_s + (int64_t)g_v0 * 8, (uint32_t)w_n * 16u);
        }
    };
    if (r < p.n_rows) {
        if constexpr (!TSM) warm_l2(r);
        issue_load(r);
    }

    const float k2 = p.k2, thr2 = p.thr2;
    const float up = p.upstream ? *p.upstream : 1.f;
    const float g = p.gscale * up;
    const bool want_grad = !(p.flags & KD_FLAG_NO_GRAD);
    const float inv_k2 = 1.f / k2;
    const u64 k2p = pk(k2, k2);
    uint32_t ph_full = 0, ph_red = 0;

    int row_iter = -1;
    int64_t next_r = r + row_stride;      // dynamic mode: the first two rows of a pipeline are static, later ones are claimed
    while (r < p.n_rows) {
        ++row_iter;
        // dynamic mode: the leader warp (cluster rank 0, first warp of the pipeline) claims the pipeline's next row now;
        // the index travels to every CTA with reduction B, long before the refill at the end of this row needs it
        // (the claim is for the row AFTER the next one: the next row is already known, so its L2 warm-up starts now)
        if (!dynamic_rows) next_r = r + row_stride;
        int claimed = 0;
        if (dynamic_rows && rank == 0 && pw == 0 && lane == 0) claimed = atomicAdd(p.row_queue, 1);
        KD_STAMP(0);
        // per-row scalars fetched early so their latency hides behind sweep A
        float lse2_t = 0.f;
        int lab = -1;
        float x_label = 0.f;
        if constexpr (VARIANT == KD_LOCA_REFERENCE) lse2_t = p.lse2_t_in[r];
        if constexpr (VARIANT == KD_LOCA_PER_ROW) {
            const int64_t l64 = p.labels[r];
            if (l64 >= 0 && l64 < p.V) {
                lab = (int)l64;
                x_label = Elem<T>::to_float(tbase[r * p.ld_t + lab]);
            } else if (tid == 0 && rank == 0) {
                atomicOr(p.status, KD_STATUS_BAD_LABEL);
            }
        }
        warm_l2(next_r);
        const uint32_t epoch = (uint32_t)r;
        const uint4* tvec = reinterpret_cast<const uint4*>(tbase + r * p.ld_t) + g_v0;     // this warp's teacher vectors
        auto load_t = [&](int v) -> uint4 {
            if constexpr (TSM) return lds128_nv(s_t + v * 16, epoch);
            else return ldg_l2(tvec + v);
        };
        auto load_s = [&](int v) -> uint4 { return lds128_nv(s_s + v * 16, epoch); };
        if (w_n > 0) mbar_wait_cta(bar_full, ph_full);
        ph_full ^= 1;
        KD_STAMP(1);

        // ================= sweep A: softmax statistics over the warp's slice ==========================
        // pass 1 (packed 16-bit min/max only) fixes the warp's reference maximum, pass 2 sums 2^((x - max)*k2)
        MaxSum ms_s, ms_t;
        Top2 top;
        ms_s.init();
        ms_t.init();
        top.init();
        float lane_min, lane_max;
        {
            uint32_t smin2 = Pair<T>::kPosInf2, smax2 = Pair<T>::kNegInf2;
#pragma unroll 4
            for (int v = lane; v < w_n; v += 32) {
                const uint4 rs = load_s(v);
                smax2 = Pair<T>::vmax(smax2, Pair<T>::vmax(Pair<T>::vmax(rs.x, rs.y), Pair<T>::vmax(rs.z, rs.w)));
                if constexpr (LOCA)
                    smin2 = Pair<T>::vmin(smin2, Pair<T>::vmin(Pair<T>::vmin(rs.x, rs.y), Pair<T>::vmin(rs.z, rs.w)));
            }
            float lo, hi;
            upk(Pair<T>::widen(smax2), lo, hi);
            lane_max = fmaxf(lo, hi);
            upk(Pair<T>::widen(smin2), lo, hi);
            lane_min = fminf(lo, hi);
        }
        ms_s.m = warp_max(lane_max);
        {
            const float mk = -ms_s.m * k2;
            const u64 mkp = pk(mk, mk);
            u64 acc0 = pk(0.f, 0.f), acc1 = pk(0.f, 0.f);
#pragma unroll 2
            for (int v = lane; v < w_n; v += 32) {
                const uint4 rs = load_s(v);
                const uint32_t w[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float e0, e1;
                    upk(fma2(Pair<T>::widen(w[i]), k2p, mkp), e0, e1);
                    if (i & 1) acc1 = add2(acc1, pk(ex2(e0), ex2(e1)));
                    else acc0 = add2(acc0, pk(ex2(e0), ex2(e1)));
                }
            }
            float a0, a1;
            upk(add2(acc0, acc1), a0, a1);
            ms_s.s = warp_sum(a0 + a1);
            if (ms_s.m == -INFINITY) ms_s.s = 0.f;      // warp without cells
        }
        if constexpr (TEACHER_STATS) {
            // teacher slice, same two passes; the runner-up candidates are found in pass 2 behind a threshold that pass 1
            // makes tight: the 2nd largest of the 32 lane maxima is a lower bound of the row's runner-up value
            uint32_t tmax2 = Pair<T>::kNegInf2;
#pragma unroll 4
            for (int v = lane; v < w_n; v += 32) {
                const uint4 rt = load_t(v);
                tmax2 = Pair<T>::vmax(tmax2, Pair<T>::vmax(Pair<T>::vmax(rt.x, rt.y), Pair<T>::vmax(rt.z, rt.w)));
            }
            float lo, hi;
            upk(Pair<T>::widen(tmax2), lo, hi);
            const float lane_tmax = fmaxf(lo, hi);
            ms_t.m = warp_max(lane_tmax);
            float thr_top = -INFINITY;
            if constexpr (VARIANT == KD_LOCA_PER_ROW) {
                // second largest lane maximum (the lane holding the maximum is masked out once)
                const unsigned holders = __ballot_sync(0xffffffffu, lane_tmax == ms_t.m);
                const bool is_first_holder = (holders != 0u) && (lane == __ffs(holders) - 1);
                thr_top = warp_max(is_first_holder ? -INFINITY : lane_tmax);
            }
            const float mk = -ms_t.m * k2;
            const u64 mkp = pk(mk, mk);
            u64 acc0 = pk(0.f, 0.f), acc1 = pk(0.f, 0.f);
#pragma unroll 2
            for (int v = lane; v < w_n; v += 32) {
                const uint4 rt = load_t(v);
                const uint32_t w[4] = {rt.x, rt.y, rt.z, rt.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float e0, e1;
                    upk(fma2(Pair<T>::widen(w[i]), k2p, mkp), e0, e1);
                    if (i & 1) acc1 = add2(acc1, pk(ex2(e0), ex2(e1)));
                    else acc0 = add2(acc0, pk(ex2(e0), ex2(e1)));
                }
                if constexpr (VARIANT == KD_LOCA_PER_ROW) {
                    const uint32_t m2 = Pair<T>::vmax(Pair<T>::vmax(rt.x, rt.y), Pair<T>::vmax(rt.z, rt.w));
                    float m_lo, m_hi;
                    upk(Pair<T>::widen(m2), m_lo, m_hi);
                    if (fmaxf(m_lo, m_hi) >= thr_top) {          // rare: a cell that may be best or runner-up
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float x0, x1;
                            upk(Pair<T>::widen(w[i]), x0, x1);
                            top.push(x0, (g_v0 + v) * 8 + 2 * i);
                            top.push(x1, (g_v0 + v) * 8 + 2 * i + 1);
                        }
                    }
                }
            }
            float a0, a1;
            upk(add2(acc0, acc1), a0, a1);
            ms_t.s = warp_sum(a0 + a1);
            if (ms_t.m == -INFINITY) ms_t.s = 0.f;
            if constexpr (VARIANT == KD_LOCA_PER_ROW) warp_merge(top);
        }

        KD_STAMP(2);
        // ---- push partial A to every CTA of the cluster, then wait for all NSLOT partials --------------
        if (lane < CS) {
            const uint32_t dst = map_to_cta(slotA + (rank * WPP + pw) * AW * 4, lane);
            const uint32_t rbar = map_to_cta(bar_redA, lane);
            if constexpr (AW == 2) {
                st_async_v2(dst, __float_as_uint(ms_s.m), __float_as_uint(ms_s.s), rbar);
            } else {
                st_async_v4(dst, __float_as_uint(ms_s.m), __float_as_uint(ms_s.s), __float_as_uint(ms_t.m), __float_as_uint(ms_t.s), rbar);
                st_async_v4(dst + 16, __float_as_uint(top.a1), (uint32_t)top.i1, __float_as_uint(top.a2), (uint32_t)top.i2, rbar);
            }
        }
        // one warp per pipeline folds the NSLOT partials and broadcasts the row statistics through shared memory; the
        // other warps park on a named barrier instead of spinning on the mbarrier
        if (pw == 0) {
            mbar_wait_cta(bar_redA, ph_red);
            if (lane == 0) mbar_expect_tx(bar_redA, NSLOT * AW * 4);      // arm the next row's phase
            const float* slots = reinterpret_cast<const float*>(smem + (slotA - s_base));
            float m_s = -INFINITY, m_t = -INFINITY;
            for (int i = lane; i < NSLOT; i += 32) {
                m_s = fmaxf(m_s, slots[i * AW]);
                if constexpr (AW == 8) m_t = fmaxf(m_t, slots[i * AW + 2]);
            }
            m_s = warp_max(m_s);
            if constexpr (AW == 8) m_t = warp_max(m_t);
            float s_sum = 0.f, t_sum = 0.f;
            Top2 all_top;
            all_top.init();
            for (int i = lane; i < NSLOT; i += 32) {
                const float* sl = slots + i * AW;
                if (sl[0] != -INFINITY) s_sum = fmaf(sl[1], ex2((sl[0] - m_s) * k2), s_sum);
                if constexpr (AW == 8) {
                    if (sl[2] != -INFINITY) t_sum = fmaf(sl[3], ex2((sl[2] - m_t) * k2), t_sum);
                    Top2 o;
                    o.a1 = sl[4]; o.i1 = __float_as_int(sl[5]); o.a2 = sl[6]; o.i2 = __float_as_int(sl[7]);
                    all_top.merge(o);
                }
            }
            s_sum = warp_sum(s_sum);
            if constexpr (TEACHER_STATS) t_sum = warp_sum(t_sum);
            if constexpr (VARIANT == KD_LOCA_PER_ROW) warp_merge(all_top);
            if (lane == 0) {
                bcast[0] = fmaf(m_s, k2, __log2f(s_sum));
                if constexpr (TEACHER_STATS) bcast[1] = fmaf(m_t, k2, __log2f(t_sum));
                if constexpr (VARIANT == KD_LOCA_PER_ROW) {
                    bcast[2] = all_top.a2;
                    bcast[3] = __int_as_float(all_top.i2);
                }
            }
        }
        named_bar_sync(barA, WPP * 32);
        KD_STAMP(3);
        const float lse2_s = bcast[0];
        if constexpr (TEACHER_STATS) lse2_t = bcast[1];
        if constexpr (VARIANT == KD_LOCA_PER_ROW) {
            top.a2 = bcast[2];
            top.i2 = __float_as_int(bcast[3]);
        }
        int ov_k = -1;
        float v1 = NaN, v2 = NaN;
        if constexpr (VARIANT == KD_LOCA_PER_ROW) {
            if (lab >= 0) {
                ov_k = (p.V > 1) ? top.i2 : 0;
                loca_scalars(x_label, top.a2, lse2_t, k2, p.alpha, v1, v2);
            }
        }
        // clamp relevance for this warp: sp >= eps  <=>  xs >= theta.  All cells above (mode 1) or all below (mode 2)
        // the threshold -> the mask is a constant and the compare / select work disappears from sweeps B and C.
        int mask_mode = 0;
        if constexpr (LOCA) {
            const float theta = (lse2_s + thr2) * inv_k2;
            const float margin = 1e-3f * (1.f + fabsf(theta));
            if (__all_sync(0xffffffffu, lane_min >= theta + margin)) mask_mode = 1;
            else if (__all_sync(0xffffffffu, lane_max < theta - margin)) mask_mode = 2;
        } else {
            mask_mode = 1;
        }
        auto cell_override = [&](int c) -> float {       // exact per-cell override (slow paths only)
            if constexpr (VARIANT == KD_LOCA_REFERENCE) return p.table[c];
            if constexpr (VARIANT == KD_LOCA_PER_ROW) return (c == ov_k) ? v2 : ((c == lab) ? v1 : NaN);
            return NaN;
        };
        auto vector_is_special = [&](int gvec) -> bool {  // vector holding one of this row's own override cells
            if constexpr (VARIANT == KD_LOCA_PER_ROW)
                return ((unsigned)(lab - gvec * 8) < 8u) || ((unsigned)(ov_k - gvec * 8) < 8u);
            return false;
        };
        const bool always_exact = (VARIANT == KD_KL_LOG_TARGET) || (VARIANT == KD_LOCA_REFERENCE && !fast_table);

        // ================= sweep B: Q_r and the row sum ===================================================
        float row = 0.f, Q = 0.f;
        {
            const u64 nls = pk(-lse2_s, -lse2_s), nlt = pk(-lse2_t, -lse2_t);
            const u64 pushp = pk(kPush, kPush);
            u64 row_pos = pk(0.f, 0.f), row_neg = pk(0.f, 0.f), Q2 = pk(0.f, 0.f);
            float row_x = 0.f, Q_x = 0.f;       // exact-path accumulators
            auto sweepB = [&](auto mode_tag) {
                constexpr int MODE = decltype(mode_tag)::value;
                pipelined_sweep<DT>(lane, w_n, load_t, [&](int v, const uint4& rt) {
                    const uint32_t wt[4] = {rt.x, rt.y, rt.z, rt.w};
                    if (always_exact || vector_is_special(g_v0 + v)) {
                        const uint4 rs = load_s(v);
                        const uint32_t wsv[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float t0, t1, s0, s1;
                            upk(Pair<T>::widen(wt[i]), t0, t1);
                            upk(Pair<T>::widen(wsv[i]), s0, s1);
                            const int c = (g_v0 + v) * 8 + 2 * i;
                            PointTerm<VARIANT>::accum(t0, s0, LOCA ? cell_override(c) : NaN, k2, lse2_t, lse2_s, thr2, row_x, Q_x);
                            PointTerm<VARIANT>::accum(t1, s1, LOCA ? cell_override(c + 1) : NaN, k2, lse2_t, lse2_s, thr2, row_x, Q_x);
                        }
                        return;
                    }
                    uint4 rs = make_uint4(0, 0, 0, 0);
                    if constexpr (MODE != 2) rs = load_s(v);
                    const uint32_t wsv[4] = {rs.x, rs.y, rs.z, rs.w};
                    uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
                    if constexpr (VARIANT == KD_LOCA_REFERENCE) { o0 = lds128_nv(s_ov + v * 16, 0u); o1 = lds128_nv(s_ov + v * 16 + cp.vpc * 16, 0u); }
                    const u64 ovp[4] = {pk(__uint_as_float(o0.x), __uint_as_float(o0.y)), pk(__uint_as_float(o0.z), __uint_as_float(o0.w)),
                                        pk(__uint_as_float(o1.x), __uint_as_float(o1.y)), pk(__uint_as_float(o1.z), __uint_as_float(o1.w))};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const u64 b2 = fma2(Pair<T>::widen(wt[i]), k2p, nlt);
                        float e_lo, e_hi;
                        if constexpr (VARIANT == KD_LOCA_REFERENCE) upk(fma2(ovp[i], pushp, b2), e_lo, e_hi);
                        else upk(b2, e_lo, e_hi);
                        const u64 tp2 = pk(ex2(e_lo), ex2(e_hi));
                        const u64 q2 = (VARIANT == KD_LOCA_REFERENCE) ? add2(tp2, ovp[i]) : tp2;
                        row_pos = fma2(tp2, b2, row_pos);
                        if constexpr (MODE == 2) {            // every cell clamped: log sp == log eps, nothing kept
                            row_neg = add2(row_neg, q2);      // scaled by thr2 after the loop
                        } else {
                            const u64 a2 = fma2(Pair<T>::widen(wsv[i]), k2p, nls);
                            if constexpr (MODE == 1) {        // no cell clamped
                                row_neg = fma2(q2, a2, row_neg);
                                Q2 = add2(Q2, q2);
                            } else {
                                float a_lo, a_hi;
                                upk(a2, a_lo, a_hi);
                                const u64 ac2 = pk(fmaxf(a_lo, thr2), fmaxf(a_hi, thr2));
                                const u64 mf2 = pk(set_ge(a_lo, thr2), set_ge(a_hi, thr2));
                                row_neg = fma2(q2, ac2, row_neg);
                                Q2 = fma2(q2, mf2, Q2);
                            }
                        }
                    }
                });
                if constexpr (MODE == 2) row_neg = mul2(row_neg, pk(thr2, thr2));
            };
            if (mask_mode == 1) sweepB(std::integral_constant<int, 1>{});
            else if (mask_mode == 2) sweepB(std::integral_constant<int, 2>{});
            else sweepB(std::integral_constant<int, 0>{});
            float x0, x1, y0, y1;
            upk(row_pos, x0, x1);
            upk(row_neg, y0, y1);
            row = (x0 + x1) - (y0 + y1) + row_x;
            upk(Q2, x0, x1);
            Q = x0 + x1 + Q_x;
        }
        row = warp_sum(row);
        Q = warp_sum(Q);
        KD_STAMP(4);
        if (lane < CS) {
            st_async_v2(map_to_cta(slotB + (rank * WPP + pw) * 8, lane), __float_as_uint(row), __float_as_uint(Q),
                        map_to_cta(bar_redB, lane));
            if (dynamic_rows && rank == 0 && pw == 0) {
                claimed = __shfl_sync((1u << CS) - 1u, claimed, 0);
                st_async_v2(map_to_cta(slotB + NSLOT * 8, lane), (uint32_t)claimed, 0u, map_to_cta(bar_redB, lane));
            }
        }
        if (pw == 0) {
            mbar_wait_cta(bar_redB, ph_red);
            if (lane == 0) mbar_expect_tx(bar_redB, redB_bytes);
            float rr = 0.f, qq = 0.f;
            for (int i = lane; i < NSLOT; i += 32) {
                const float2 sl = reinterpret_cast<const float2*>(smem + (slotB - s_base))[i];
                rr += sl.x;
                qq += sl.y;
            }
            row = warp_sum(rr);
            Q = warp_sum(qq);
            if (lane == 0) {
                bcast[8] = Q;
                if (dynamic_rows) bcast[9] = __int_as_float(reinterpret_cast<const int*>(smem + (slotB - s_base))[NSLOT * 2]);
                if (rank == 0) {
                    if constexpr (VARIANT == KD_LOCA_PER_ROW) {
                        if (lab >= 0) {
                            if (ov_k != lab && v1 != 0.f) row += v1 * log2f(v1);
                            if (v2 != 0.f) row += v2 * log2f(v2);
                        } else {
                            row = NaN;
                        }
                    }
                    p.rowaux[r] = make_float4(lse2_t, lse2_s, Q, row);
                    if constexpr (VARIANT == KD_LOCA_PER_ROW) {
                        if (p.records) {
                            kd_token_record rec;
                            rec.label = lab; rec.v1 = v1; rec.k = ov_k; rec.v2 = v2;
                            p.records[r] = rec;
                        }
                    }
                }
            }
        }
        named_bar_sync(barB, WPP * 32);
        KD_STAMP(5);
        Q = bcast[8];
        int64_t after_next = 0;
        if (dynamic_rows) after_next = __float_as_int(bcast[9]);
        ph_red ^= 1;

        // ================= sweep C: gradient, written straight to HBM ===================================
        if (want_grad) {
            const float gQ = g * Q;
            const bool fast_sign = (g > 0.f) && (gQ > 0.f) && (gQ < INFINITY);
            // fast path folds g*Q and g into the exponents: 2^(a + log2 gQ) - mask * (2^(b + log2 g) + g*ov)
            const float lgQ = fast_sign ? __log2f(gQ) : 0.f, lg = fast_sign ? __log2f(g) : 0.f;
            const u64 nls = pk(lgQ - lse2_s, lgQ - lse2_s), nlt = pk(lg - lse2_t, lg - lse2_t);
            const float thrC = thr2 + lgQ;          // a >= thr2  <=>  a + lgQ >= thr2 + lgQ
            const u64 gp = pk(g, g), pushp = pk(kPush, kPush), neg1 = pk(-1.f, -1.f);
            T* grow = static_cast<T*>(p.grad) + r * p.ld_g + (int64_t)g_v0 * 8;
            const bool exact_row = !fast_sign || always_exact;
            auto store_out = [&](int v, const uint32_t (&out)[4]) {
                stg_stream(grow + (int64_t)v * 8, make_uint4(out[0], out[1], out[2], out[3]));
            };
            if (mask_mode == 2 && !exact_row && VARIANT != KD_LOCA_PER_ROW) {
                // every cell clamped: grad = g*Q*sp, the teacher is not needed at all
#pragma unroll 2
                for (int v = lane; v < w_n; v += 32) {
                    const uint4 rs = load_s(v);
                    const uint32_t wsv[4] = {rs.x, rs.y, rs.z, rs.w};
                    uint32_t out[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float a_lo, a_hi;
                        upk(fma2(Pair<T>::widen(wsv[i]), k2p, nls), a_lo, a_hi);
                        out[i] = Pair<T>::narrow(ex2(a_lo), ex2(a_hi));
                    }
                    store_out(v, out);
                }
            } else {
                auto sweepC = [&](auto mode_tag) {
                    constexpr int MODE = decltype(mode_tag)::value;
                    pipelined_sweep<DT>(lane, w_n, load_t, [&](int v, const uint4& rt) {
                        const uint4 rs = load_s(v);
                        const uint32_t wsv[4] = {rs.x, rs.y, rs.z, rs.w};
                        const uint32_t wt[4] = {rt.x, rt.y, rt.z, rt.w};
                        uint32_t out[4];
                        if (exact_row || vector_is_special(g_v0 + v)) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float t0, t1, s0, s1;
                                upk(Pair<T>::widen(wt[i]), t0, t1);
                                upk(Pair<T>::widen(wsv[i]), s0, s1);
                                const int c = (g_v0 + v) * 8 + 2 * i;
                                const float g0 = PointTerm<VARIANT>::grad(t0, s0, LOCA ? cell_override(c) : NaN, k2, lse2_t, lse2_s, thr2, g, gQ);
                                const float g1 = PointTerm<VARIANT>::grad(t1, s1, LOCA ? cell_override(c + 1) : NaN, k2, lse2_t, lse2_s, thr2, g, gQ);
                                out[i] = Pair<T>::narrow(g0, g1);
                            }
                        } else {
                            uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
                            if constexpr (VARIANT == KD_LOCA_REFERENCE) { o0 = lds128_nv(s_ov + v * 16, 0u); o1 = lds128_nv(s_ov + v * 16 + cp.vpc * 16, 0u); }
                            const u64 ovp[4] = {pk(__uint_as_float(o0.x), __uint_as_float(o0.y)), pk(__uint_as_float(o0.z), __uint_as_float(o0.w)),
                                                pk(__uint_as_float(o1.x), __uint_as_float(o1.y)), pk(__uint_as_float(o1.z), __uint_as_float(o1.w))};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float a_lo, a_hi, e_lo, e_hi;
                                upk(fma2(Pair<T>::widen(wsv[i]), k2p, nls), a_lo, a_hi);          // log2(g*Q*sp)
                                const u64 b2 = fma2(Pair<T>::widen(wt[i]), k2p, nlt);             // log2(g*tp)
                                if constexpr (VARIANT == KD_LOCA_REFERENCE) upk(fma2(ovp[i], pushp, b2), e_lo, e_hi);
                                else upk(b2, e_lo, e_hi);
                                const u64 sp2 = pk(ex2(a_lo), ex2(a_hi));
                                u64 qg2 = pk(ex2(e_lo), ex2(e_hi));
                                if constexpr (VARIANT == KD_LOCA_REFERENCE) qg2 = fma2(ovp[i], gp, qg2);
                                float n_lo, n_hi;
                                if constexpr (MODE == 1) {
                                    upk(fma2(qg2, neg1, sp2), n_lo, n_hi);
                                } else {
                                    float q_lo, q_hi, s_lo, s_hi;
                                    upk(qg2, q_lo, q_hi);
                                    upk(sp2, s_lo, s_hi);
                                    n_lo = fmaf(-q_lo, set_ge(a_lo, thrC), s_lo);
                                    n_hi = fmaf(-q_hi, set_ge(a_hi, thrC), s_hi);
                                }
                                out[i] = Pair<T>::narrow(n_lo, n_hi);
                            }
                        }
                        store_out(v, out);
                    });
                };
                if (mask_mode == 1) sweepC(std::integral_constant<int, 1>{});
                else sweepC(std::integral_constant<int, 0>{});
            }
        }
        KD_STAMP(6);
        // ---- refill the warp's private slice with its next row (all lanes are done reading it) -----------
        __syncwarp();
        if (next_r < p.n_rows) issue_load(next_r);
        r = next_r;
        if (dynamic_rows) next_r = after_next;
    }
    cluster_sync_all();     // nobody leaves while peers may still push into its shared memory
}

// --------------------------------------------------------------------------------------------------------
template <int VARIANT, typename T, int CS, int NPIPE, bool TSM, bool DYN>
int launch_variant_d(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    auto kfn = row_emit_cluster_kernel<VARIANT, T, CS, NPIPE, TSM, DYN>;
    const int nvec = e.V / 8;
    const int vpc = (nvec + CS - 1) / CS;
    const SmemPlan<VARIANT, CS, NPIPE, TSM> plan(vpc);
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (plan.total > max_smem) return KD_ERR_UNSUPPORTED;
    static int configured_dev = -1;
    static int max_clusters = 0;
    cudaError_t err;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = plan.total;
    cfg.stream = st;
    if (configured_dev != dev) {
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(cluster kernel smem)");
        cfg.gridDim = dim3(CS * 64);
        err = cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &cfg);
        if (err != cudaSuccess || max_clusters < 1) {
            cudaGetLastError();
            max_clusters = 0;
            return KD_ERR_UNSUPPORTED;
        }
        configured_dev = dev;
    }
    if (capacity) *capacity = max_clusters;
    if (query_only) return KD_OK;
    if (max_clusters < 1) return KD_ERR_UNSUPPORTED;
    const int64_t want = (e.n_rows + NPIPE - 1) / NPIPE;
    ClusterParams cp;
    cp.trace = nullptr;
#ifdef KD_TRACE
    {
        static unsigned long long* tbuf = nullptr;
        if (!tbuf) cudaMalloc(&tbuf, (size_t)max_clusters * CS * kWarps * 64 * 8 * sizeof(unsigned long long));
        cp.trace = tbuf;
        if (const char* pth = getenv("KD_TRACE_DUMP")) {       // previous launch's stamps -> file (debug only)
            static int calls = 0;
            if (++calls == 3) {
                cudaDeviceSynchronize();
                const size_t n = (size_t)max_clusters * CS * kWarps * 64 * 8;
                unsigned long long* h = (unsigned long long*)malloc(n * 8);
                cudaMemcpy(h, tbuf, n * 8, cudaMemcpyDeviceToHost);
                FILE* f = fopen(pth, "wb");
                fwrite(h, 8, n, f);
                fclose(f);
                free(h);
            }
        }
    }
#endif
    cp.e = e;
    cp.n_clusters = (int)(want < max_clusters ? want : max_clusters);
    cp.vpc = vpc;
    cp.nvec = nvec;
    cfg.gridDim = dim3(cp.n_clusters * CS);
    if (launched) *launched = cp.n_clusters;
    if (first_free) *first_free = 2 * cp.n_clusters * NPIPE;
    err = cudaLaunchKernelEx(&cfg, kfn, cp);
    if (err != cudaSuccess) return cuda_fail(err, "row_emit_cluster_kernel launch");
    return KD_OK;
}

template <int VARIANT, typename T, int CS, int NPIPE, bool TSM>
int launch_variant(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched = nullptr, int* first_free = nullptr) {
    if (e.row_queue) return launch_variant_d<VARIANT, T, CS, NPIPE, TSM, true>(e, st, query_only, capacity, launched, first_free);
    return launch_variant_d<VARIANT, T, CS, NPIPE, TSM, false>(e, st, query_only, capacity, launched, first_free);
}

// configuration: 0 = teacher + student slices resident, 2 row pipelines; 1 = student resident, teacher via L2, 4 pipelines
int cluster_config() {
    static int cfg = -1;
    if (cfg < 0) {
        const char* e = getenv("KD_CLUSTER_CONFIG");
        cfg = e ? atoi(e) : 0;
    }
    return cfg;
}

template <int VARIANT, typename T>
int launch_cfg(const EmitParams& e, cudaStream_t st, int* launched, int* first_free) {
    if (cluster_config() == 0) return launch_variant<VARIANT, T, 8, 2, true>(e, st, false, nullptr, launched, first_free);
    return launch_variant<VARIANT, T, 8, 4, false>(e, st, false, nullptr, launched, first_free);
}

template <typename T>
int launch_dtype(const EmitParams& e, int variant, cudaStream_t st, int* launched, int* first_free) {
    switch (variant) {
        case KD_LOCA_REFERENCE: return launch_cfg<KD_LOCA_REFERENCE, T>(e, st, launched, first_free);
        case KD_LOCA_PER_ROW: return launch_cfg<KD_LOCA_PER_ROW, T>(e, st, launched, first_free);
        case KD_KL_PLAIN: return launch_cfg<KD_KL_PLAIN, T>(e, st, launched, first_free);
        case KD_KL_LOG_TARGET: return launch_cfg<KD_KL_LOG_TARGET, T>(e, st, launched, first_free);
    }
    return KD_ERR_UNSUPPORTED;
}

bool aligned_rows(const void* base, int64_t ld) { return (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (ld % 8) == 0; }

}  // namespace

int cluster_capacity(int cluster_size, int, size_t) {
    int cap = 0;
    EmitParams e{};
    e.V = 151936;
    if (cluster_size == 8) launch_variant<KD_LOCA_REFERENCE, __nv_bfloat16, 8, 4, false>(e, nullptr, true, &cap);
    return cap;
}

static bool cluster_eligible(const EmitParams& p, int dtype) {
    if (dtype == KD_DTYPE_F32) return false;                                   // 16-bit logits only
    if (p.V % 8 != 0 || p.V < 8 * 8 * 32) return false;                        // whole 16-byte vectors, enough columns per warp
    if (!aligned_rows(p.teacher, p.ld_t) || !aligned_rows(p.student, p.ld_s)) return false;
    if (!(p.flags & KD_FLAG_NO_GRAD) && !aligned_rows(p.grad, p.ld_g)) return false;
    return p.n_rows >= 1;
}

int cluster_launch_shape(const EmitParams& p, int variant, int dtype, int* n_clusters, int* ctas_per_cluster, int* pipelines,
                         int* static_rows) {
    if (!cluster_eligible(p, dtype)) return KD_ERR_UNSUPPORTED;
    int cap = 0, rc;
    int in_flight = 2;
    if (fused_applies(p, variant, dtype) && launch_emit_fused(p, variant, dtype, nullptr, true, &cap, nullptr, nullptr, &in_flight) == KD_OK && cap >= 1) {
        const int64_t want2 = (p.n_rows + in_flight - 1) / in_flight;
        *n_clusters = (int)(want2 < cap ? want2 : cap);
        *ctas_per_cluster = fused_cluster_size();
        *pipelines = in_flight;
        *static_rows = in_flight == 1 ? 3 : 4;       // pipelined kernel: 3 rows; two pipelines: 2 rows each
        return KD_OK;
    }
    const int pipes = cluster_config() == 0 ? 2 : 4;
#define KD_QUERY(VAR, TT) (cluster_config() == 0 ? launch_variant<VAR, TT, 8, 2, true>(p, nullptr, true, &cap) \
                                                 : launch_variant<VAR, TT, 8, 4, false>(p, nullptr, true, &cap))
#define KD_QUERY_T(TT) (variant == KD_LOCA_REFERENCE ? KD_QUERY(KD_LOCA_REFERENCE, TT) : variant == KD_LOCA_PER_ROW ? KD_QUERY(KD_LOCA_PER_ROW, TT) \
                        : variant == KD_KL_PLAIN ? KD_QUERY(KD_KL_PLAIN, TT) : KD_QUERY(KD_KL_LOG_TARGET, TT))
    rc = (dtype == KD_DTYPE_BF16) ? KD_QUERY_T(__nv_bfloat16) : KD_QUERY_T(__half);
#undef KD_QUERY_T
#undef KD_QUERY
    if (rc != KD_OK) return rc;
    if (cap < 1) return KD_ERR_UNSUPPORTED;
    const int64_t want = (p.n_rows + pipes - 1) / pipes;
    *n_clusters = (int)(want < cap ? want : cap);
    *ctas_per_cluster = 8;
    *pipelines = pipes;
    *static_rows = 2 * pipes;
    return KD_OK;
}

int launch_emit_cluster(const EmitParams& p, int variant, int dtype, cudaStream_t stream, int* clusters_launched,
                        int* ctas_per_cluster, int* first_free_row) {
    if (ctas_per_cluster) *ctas_per_cluster = 8;
    if (!cluster_eligible(p, dtype)) return KD_ERR_UNSUPPORTED;
    if (fused_applies(p, variant, dtype)) {
        const int rc = launch_emit_fused(p, variant, dtype, stream, false, nullptr, clusters_launched, first_free_row, nullptr);
        if (rc != KD_ERR_UNSUPPORTED) return rc;
    }
    if (p.ce_labels) return KD_ERR_UNSUPPORTED;      // the fused cross-entropy lives in the fused kernel only
    if (dtype == KD_DTYPE_BF16) return launch_dtype<__nv_bfloat16>(p, variant, stream, clusters_launched, first_free_row);
    return launch_dtype<__half>(p, variant, stream, clusters_launched, first_free_row);
}

}  // namespace kd
