// logit_kd_fused.cu — two-sweep, one-reduction variant of the cluster-resident logit-KD emit kernel (sm_100a).
//
// Same residency scheme as logit_kd_cluster.cu (a cluster of CS CTAs keeps the teacher + student row pair in
// distributed shared memory, fetched once with 1-D TMA bulk copies; per-warp private slices and refills; st.async
// reductions), but the arithmetic is reorganised so that a row costs ~half the instructions and 3 instead of 4
// MUFU.EX2 per cell:
//
//   pass 0   packed 16-bit min/max of the student slice -> the warp's reference maximum m_w
//   pass AB  ONE sweep over both slices: e = 2^(xs*k2 - m_w*k2) summed for the student statistics, q = 2^b (or the
//            override) and the three sums  Q_w = sum q,  R_w = sum t*b,  P_w = sum q*(xs*k2 - m_w*k2).  Because
//            a = log2 sp = (xs*k2 - m_w*k2) + (m_w*k2 - lse2_s) is LINEAR in the still unknown lse2_s, the row sum
//            sum q*(lq - a) = sum_w [ R_w - P_w - (m_w*k2 - lse2_s) * Q_w ] is assembled after the (single) cluster
//            reduction.  q is parked, in fp32, in TENSOR MEMORY: with the 32x32b access shape every thread owns 128
//            private 32-bit words of TMEM, enough for the 80 cells it handles per row; shared memory keeps the raw rows.
//   reduce   one st.async push of (m_w, S_w, Q_w, R_w - P_w, min_w) per warp to every CTA of the cluster
//   [pass B' only for rows where the sp >= eps clamp can bite (min a < thr2): exact masked sums from the raw student
//            slice and the parked q, followed by a second reduction]
//   pass C   gradient: sp from the raw student slice (one ex2), q from tensor memory, 128-bit streaming stores
//
// Variants: KD_LOCA_REFERENCE (teacher lse2_t and the override table come from phase A) and KD_KL_PLAIN (teacher
// statistics are folded into the same single reduction through a per-warp teacher reference maximum).
#include <stdlib.h>

#include <type_traits>

#include "cluster_prims.cuh"
#include "logit_kd.cuh"

namespace kd {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr float kPush = -3.0e38f;       // ov' * kPush drives log2 q of an overridden cell to -inf
constexpr float kOvFloor = 1.0e-36f;    // encoded "overridden with ~0" (must stay > 0)
constexpr int kTmemColsPerWarp = 128;   // 512 columns / 4 warps per lane quarter
constexpr int kMaxIters = kTmemColsPerWarp / 8;

struct FusedParams {
    EmitParams e;
    int n_clusters;
    int vpc;                      // 16-byte vectors per CTA slice
    int nvec;                     // 16-byte vectors per row
};

// Shared-memory plan (dynamic, byte offsets).  A reduction slot is 8 words:
//   [0] m_w  [1] S_w  [2] Q_w  [3] R_w - P_w  [4] min_w  [5] next-row claim (leader slot)  [6] [7] second reduction
//   (KD_KL_PLAIN: [6] teacher reference maximum, no second reduction)
template <int VARIANT, int CS, int NPIPE>
struct FusedPlan {
    static constexpr int WPP = kWarps / NPIPE;
    static constexpr int NSLOT = CS * WPP;
    int table, rows, bars, slots, bcast, tmem_word, total;
    __host__ __device__ FusedPlan(int vpc) {
        int off = 0;
        table = off;
        if (VARIANT == KD_LOCA_REFERENCE) off += vpc * 32;
        rows = off;
        off += NPIPE * 2 * vpc * 16;
        bars = off;
        off += (kWarps + 2 * NPIPE) * 8;
        slots = off;
        off += NPIPE * NSLOT * 32;
        bcast = off;
        off += NPIPE * 32;
        tmem_word = off;
        off += 16;
        total = off;
    }
};

__device__ __forceinline__ uint4 lds128_ep(uint32_t addr, uint32_t epoch) {     // see lds128_nv in logit_kd_cluster.cu
    uint4 v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];  // epoch %5" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "r"(epoch));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int VARIANT, typename T, int CS, int NPIPE, bool DYN>
__global__ void __launch_bounds__(kThreads, 1)
row_emit_fused_kernel(const FusedParams cp) {
    constexpr bool REF = (VARIANT == KD_LOCA_REFERENCE);
    constexpr int WPP = kWarps / NPIPE;
    constexpr int NSLOT = CS * WPP;
    constexpr uint32_t RED1_BYTES = NSLOT * (REF ? 24 : 32);
    constexpr uint32_t RED2_BYTES = NSLOT * 8;
    static_assert(NSLOT % 32 == 0, "the fold assumes whole warps of slots");
    extern __shared__ __align__(128) unsigned char smem[];
    const EmitParams& p = cp.e;
    const FusedPlan<VARIANT, CS, NPIPE> plan(cp.vpc);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pipe = warp / WPP, pw = warp % WPP;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x / CS;
    const float NaN = __int_as_float(0x7fc00000);

    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar_full = s_base + plan.bars + warp * 8;
    const uint32_t bar_red1 = s_base + plan.bars + (kWarps + 2 * pipe) * 8;
    const uint32_t bar_red2 = bar_red1 + 8;
    const uint32_t slots = s_base + plan.slots + pipe * NSLOT * 32;
    const float4* slot_v = reinterpret_cast<const float4*>(smem + plan.slots + pipe * NSLOT * 32);
    volatile float* bcast = reinterpret_cast<volatile float*>(smem + plan.bcast + pipe * 32);
    const int bar1 = 1 + 2 * pipe, bar2 = 2 + 2 * pipe;     // named barriers of this pipeline (0 is __syncthreads)

    // this CTA's column slice and this warp's private part of it (in 16-byte vectors)
    const int cta_v0 = min((int)rank * cp.vpc, cp.nvec);
    const int cta_v1 = min(cta_v0 + cp.vpc, cp.nvec);
    const int cta_n = cta_v1 - cta_v0;
    const int w_chunk = ((cta_n + WPP * 32 - 1) / (WPP * 32)) * 32;
    const int w_lo = min(pw * w_chunk, cta_n);
    const int w_hi = min(w_lo + w_chunk, cta_n);
    const int w_n = w_hi - w_lo;
    const int n_full = w_n >> 5, n_tail = w_n & 31;      // whole warp iterations + lanes of the partial one
    const uint32_t s_s = s_base + plan.rows + pipe * 2 * cp.vpc * 16 + w_lo * 16;
    const uint32_t s_t = s_s + cp.vpc * 16;
    const uint32_t s_ov = s_base + plan.table + w_lo * 16;       // cells 0-3; cells 4-7 live cp.vpc*16 bytes further
    const int g_v0 = cta_v0 + w_lo;

    if (tid == 0) {
        for (int w = 0; w < kWarps + 2 * NPIPE; ++w) mbar_init(s_base + plan.bars + w * 8, 1);
        fence_mbar_init();
        for (int i = 0; i < NPIPE; ++i) {        // arm the first phase of every reduction barrier
            mbar_expect_tx(s_base + plan.bars + (kWarps + 2 * i) * 8, RED1_BYTES);
            mbar_expect_tx(s_base + plan.bars + (kWarps + 2 * i + 1) * 8, RED2_BYTES);
        }
    }
    if (warp == 1) tmem_alloc_512(s_base + plan.tmem_word);
    // LoCa: encode this CTA's table slice once (0 = not overridden, else max(ov, floor) > 0); layout as in logit_kd_cluster.cu
    float poison = 0.f;
    if constexpr (REF) {
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
        tmem_fence_before_sync();
        // a negative / non-finite override (alpha > 1; the reference's loss is NaN there too) poisons the row sums
        if (!__syncthreads_and(ok)) poison = NaN;
    } else {
        tmem_fence_before_sync();
        __syncthreads();
    }
    tmem_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + plan.tmem_word);
    const uint32_t tq = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(kTmemColsPerWarp * (warp >> 2));
    cluster_sync_all();     // barriers initialised cluster-wide before any remote push

    const int row_stride = cp.n_clusters * NPIPE;
    int64_t r = (int64_t)cluster_id * NPIPE + pipe;
    const T* tbase = static_cast<const T*>(p.teacher);
    const T* sbase = static_cast<const T*>(p.student);
    auto issue_load = [&](int64_t row) {
        if (lane == 0 && w_n > 0) {
            const uint32_t bytes = (uint32_t)w_n * 16u;
            fence_proxy_async();
            mbar_expect_tx(bar_full, 2u * bytes);
            tma_load_1d(s_s, sbase + row * p.ld_s + (int64_t)g_v0 * 8, bytes, bar_full);
            tma_load_1d(s_t, tbase + row * p.ld_t + (int64_t)g_v0 * 8, bytes, bar_full);
        }
    };
    auto warm_l2 = [&](int64_t row) {           // bring the warp's slice of `row` into L2 ahead of its use
        if (lane == 0 && w_n > 0 && row < p.n_rows) {
            prefetch_l2(tbase + row * p.ld_t + (int64_t)g_v0 * 8, (uint32_t)w_n * 16u);
            prefetch_l2(sbase + row * p.ld_s + (int64_t)g_v0 * 8, (uint32_t)w_n * 16u);
        }
    };
    if (r < p.n_rows) issue_load(r);

    const float k2 = p.k2, thr2 = p.thr2;
    const float up = p.upstream ? *p.upstream : 1.f;
    const float g = p.gscale * up;
    const bool want_grad = !(p.flags & KD_FLAG_NO_GRAD);
    const u64 k2p = pk(k2, k2);
    const u64 pushp = pk(kPush, kPush);
    uint32_t ph_full = 0, ph1 = 0, ph2 = 0;

    int64_t next_r = r + row_stride;      // dynamic mode: the first two rows of a pipeline are static, later ones are claimed
    while (r < p.n_rows) {
        if (!DYN) next_r = r + row_stride;
        int claimed = 0;
        if (DYN && rank == 0 && pw == 0 && lane == 0) claimed = atomicAdd(p.row_queue, 1);
        float lse2_t = 0.f;
        if constexpr (REF) lse2_t = p.lse2_t_in[r];
        warm_l2(next_r);
        const uint32_t epoch = (uint32_t)r;
        auto load_s = [&](int v) -> uint4 { return lds128_ep(s_s + v * 16, epoch); };
        auto load_t = [&](int v) -> uint4 { return lds128_ep(s_t + v * 16, epoch); };
        if (w_n > 0) mbar_wait_cta(bar_full, ph_full);
        ph_full ^= 1;

        // ================= pass 0: reference maxima (and the student minimum) of the warp's slice ==============
        float m_w, min_w = INFINITY, mt_w = -INFINITY;
        {
            uint32_t smax2 = Pair<T>::kNegInf2, smin2 = Pair<T>::kPosInf2, tmax2 = Pair<T>::kNegInf2;
#pragma unroll 4
            for (int v = lane; v < w_n; v += 32) {
                const uint4 rs = load_s(v);
                smax2 = Pair<T>::vmax(smax2, Pair<T>::vmax(Pair<T>::vmax(rs.x, rs.y), Pair<T>::vmax(rs.z, rs.w)));
                if constexpr (REF) {
                    smin2 = Pair<T>::vmin(smin2, Pair<T>::vmin(Pair<T>::vmin(rs.x, rs.y), Pair<T>::vmin(rs.z, rs.w)));
                } else {
                    const uint4 rt = load_t(v);
                    tmax2 = Pair<T>::vmax(tmax2, Pair<T>::vmax(Pair<T>::vmax(rt.x, rt.y), Pair<T>::vmax(rt.z, rt.w)));
                }
            }
            float lo, hi;
            upk(Pair<T>::widen(smax2), lo, hi);
            m_w = warp_max(fmaxf(lo, hi));
            if constexpr (REF) {
                upk(Pair<T>::widen(smin2), lo, hi);
                min_w = warp_min(fminf(lo, hi));
            } else {
                upk(Pair<T>::widen(tmax2), lo, hi);
                mt_w = warp_max(fmaxf(lo, hi));
            }
        }
        // exponent references (a warp whose cells are all -inf, or that has no cells, uses 0)
        const float mk = __fmul_rn((m_w == -INFINITY) ? 0.f : m_w, k2);
        const float mtk = __fmul_rn((mt_w == -INFINITY) ? 0.f : mt_w, k2);

        // ================= pass AB: statistics + accumulation in one sweep, q parked in tensor memory ==========
        float S_w, Q_w, R_w, P_w;
        {
            const u64 nmk = pk(-mk, -mk);
            const float tref = REF ? lse2_t : mtk;
            const u64 nlt = pk(-tref, -tref);
            u64 accS0 = pk(0.f, 0.f), accS1 = pk(0.f, 0.f), accQ = pk(0.f, 0.f), accR = pk(0.f, 0.f), accP = pk(0.f, 0.f);
            auto ab_body = [&](int j, bool act, auto tail_tag) {
                uint32_t qb[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                if (act) {
                    const int v = lane + 32 * j;
                    const uint4 rs = load_s(v), rt = load_t(v);
                    uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
                    if constexpr (REF) { o0 = lds128_ep(s_ov + v * 16, 0u); o1 = lds128_ep(s_ov + v * 16 + cp.vpc * 16, 0u); }
                    const uint32_t ws[4] = {rs.x, rs.y, rs.z, rs.w}, wt[4] = {rt.x, rt.y, rt.z, rt.w};
                    const u64 ovp[4] = {pk(__uint_as_float(o0.x), __uint_as_float(o0.y)), pk(__uint_as_float(o0.z), __uint_as_float(o0.w)),
                                        pk(__uint_as_float(o1.x), __uint_as_float(o1.y)), pk(__uint_as_float(o1.z), __uint_as_float(o1.w))};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const u64 a2 = fma2(Pair<T>::widen(ws[i]), k2p, nmk);        // xs*k2 - m_w*k2  (<= 0)
                        float a_lo, a_hi;
                        upk(a2, a_lo, a_hi);
                        const u64 e2 = pk(ex2(a_lo), ex2(a_hi));
                        if (i & 1) accS1 = add2(accS1, e2);
                        else accS0 = add2(accS0, e2);
                        const u64 b2 = fma2(Pair<T>::widen(wt[i]), k2p, nlt);        // log2 tp (REF) / relative to mt_w (PLAIN)
                        float e_lo, e_hi;
                        if constexpr (REF) upk(fma2(ovp[i], pushp, b2), e_lo, e_hi);
                        else upk(b2, e_lo, e_hi);
                        const u64 t2 = pk(ex2(e_lo), ex2(e_hi));
                        const u64 q2 = REF ? add2(t2, ovp[i]) : t2;
                        accQ = add2(accQ, q2);
                        accR = fma2(t2, b2, accR);
                        accP = fma2(q2, a2, accP);
                        float q_lo, q_hi;
                        upk(q2, q_lo, q_hi);
                        qb[2 * i] = __float_as_uint(q_lo);
                        qb[2 * i + 1] = __float_as_uint(q_hi);
                    }
                }
                if constexpr (decltype(tail_tag)::value) __syncwarp();     // reconverge before the warp-collective store
                tmem_st8(tq + 8 * j, qb);
            };
#pragma unroll 2
            for (int j = 0; j < n_full; ++j) ab_body(j, true, std::false_type{});
            if (n_tail) ab_body(n_full, lane < n_tail, std::true_type{});
            tmem_wait_st();
            float x0, x1;
            upk(add2(accS0, accS1), x0, x1);
            S_w = warp_sum(x0 + x1);
            upk(accQ, x0, x1);
            Q_w = warp_sum(x0 + x1);
            upk(accR, x0, x1);
            R_w = warp_sum(x0 + x1) + poison;
            upk(accP, x0, x1);
            P_w = warp_sum(x0 + x1);
        }

        // ---- the single reduction: push the warp's partial to every CTA of the cluster ---------------------------
        if (DYN) claimed = __shfl_sync(0xffffffffu, claimed, 0);
        if (lane < CS) {
            const uint32_t dst = map_to_cta(slots + (rank * WPP + pw) * 32, lane);
            const uint32_t rbar = map_to_cta(bar_red1, lane);
            st_async_v4(dst, __float_as_uint(m_w), __float_as_uint(S_w), __float_as_uint(Q_w), __float_as_uint(R_w - P_w), rbar);
            if constexpr (REF) st_async_v2(dst + 16, __float_as_uint(min_w), (uint32_t)claimed, rbar);
            else st_async_v4(dst + 16, 0u, (uint32_t)claimed, __float_as_uint(mt_w), 0u, rbar);
        }
        // one warp per pipeline folds the NSLOT partials and broadcasts the row statistics through shared memory
        if (pw == 0) {
            mbar_wait_cta(bar_red1, ph1);
            if (lane == 0) mbar_expect_tx(bar_red1, RED1_BYTES);      // arm the next row's phase
            constexpr int SPL = NSLOT / 32;                            // slots per lane
            float4 sa[SPL], sb[SPL];
            float m_s = -INFINITY, m_t = -INFINITY, mn = INFINITY;
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                sa[i] = slot_v[2 * (lane + 32 * i)];
                sb[i] = slot_v[2 * (lane + 32 * i) + 1];
                m_s = fmaxf(m_s, sa[i].x);
                if constexpr (REF) mn = fminf(mn, sb[i].x);
                else m_t = fmaxf(m_t, sb[i].z);
            }
            m_s = warp_max(m_s);
            if constexpr (REF) mn = warp_min(mn);
            else m_t = warp_max(m_t);
            float s_sum = 0.f, t_sum = 0.f;
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                if (sa[i].x != -INFINITY) s_sum = fmaf(sa[i].y, ex2((sa[i].x - m_s) * k2), s_sum);
                if constexpr (!REF) {
                    if (sb[i].z != -INFINITY) t_sum = fmaf(sa[i].z, ex2((sb[i].z - m_t) * k2), t_sum);
                }
            }
            s_sum = warp_sum(s_sum);
            const float lse2_s = fmaf(m_s, k2, __log2f(s_sum));
            float Q = 0.f, row = 0.f;
            if constexpr (REF) {
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const float mki = __fmul_rn((sa[i].x == -INFINITY) ? 0.f : sa[i].x, k2);
                    Q += sa[i].z;
                    row += sa[i].w - (mki - lse2_s) * sa[i].z;
                }
            } else {
                t_sum = warp_sum(t_sum);
                lse2_t = fmaf(m_t, k2, __log2f(t_sum));
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const float mki = __fmul_rn((sa[i].x == -INFINITY) ? 0.f : sa[i].x, k2);
                    const float mti = __fmul_rn((sb[i].z == -INFINITY) ? 0.f : sb[i].z, k2);
                    const float sc = ex2(mti - lse2_t);               // this warp's teacher partials are relative to its mt_w
                    Q = fmaf(sc, sa[i].z, Q);
                    row = fmaf(sc, sa[i].w + ((mti - lse2_t) - (mki - lse2_s)) * sa[i].z, row);
                }
            }
            Q = warp_sum(Q);
            row = warp_sum(row);
            int claim_all = 0;
            if (DYN) claim_all = __shfl_sync(0xffffffffu, __float_as_int(sb[0].y), 0);     // slot 0 = (rank 0, warp 0)
            bool all_kept = true;
            if constexpr (REF) {
                const float a_min = fmaf(mn, k2, -lse2_s);
                all_kept = a_min >= thr2 + 1e-3f * (1.f + fabsf(thr2));       // NaN -> masked path
            }
            if (lane == 0) {
                bcast[0] = lse2_s;
                bcast[1] = lse2_t;
                bcast[2] = Q;
                bcast[3] = all_kept ? 1.f : 0.f;
                bcast[4] = __int_as_float(claim_all);
                if (rank == 0 && all_kept) p.rowaux[r] = make_float4(lse2_t, lse2_s, Q, row);
            }
        }
        named_bar_sync(bar1, WPP * 32);
        const float lse2_s = bcast[0];
        if constexpr (!REF) lse2_t = bcast[1];
        float Q = bcast[2];
        const bool all_kept = bcast[3] != 0.f;
        int64_t after_next = 0;
        if (DYN) after_next = __float_as_int(bcast[4]);
        ph1 ^= 1;

        // ================= pass B' (rare): exact masked sums for rows the sp >= eps clamp touches =================
        if constexpr (REF) {
            if (!all_kept) {
                const u64 nls = pk(-lse2_s, -lse2_s);
                u64 accN = pk(0.f, 0.f), accM = pk(0.f, 0.f);
                auto bp_body = [&](int j, bool act, auto tail_tag) {
                    uint32_t qb[8];
                    tmem_ld8(tq + 8 * j, qb);
                    uint4 rs = make_uint4(0, 0, 0, 0);
                    if (act) rs = load_s(lane + 32 * j);
                    tmem_wait_ld();
                    if (act) {
                        const uint32_t ws[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float a_lo, a_hi;
                            upk(fma2(Pair<T>::widen(ws[i]), k2p, nls), a_lo, a_hi);
                            const u64 q2 = pk(__uint_as_float(qb[2 * i]), __uint_as_float(qb[2 * i + 1]));
                            accN = fma2(q2, pk(fmaxf(a_lo, thr2), fmaxf(a_hi, thr2)), accN);
                            accM = fma2(q2, pk(set_ge(a_lo, thr2), set_ge(a_hi, thr2)), accM);
                        }
                    }
                    if constexpr (decltype(tail_tag)::value) __syncwarp();
                };
#pragma unroll 2
                for (int j = 0; j < n_full; ++j) bp_body(j, true, std::false_type{});
                if (n_tail) bp_body(n_full, lane < n_tail, std::true_type{});
                float x0, x1;
                upk(accN, x0, x1);
                const float N_w = warp_sum(x0 + x1);
                upk(accM, x0, x1);
                const float M_w = warp_sum(x0 + x1);
                if (lane < CS)
                    st_async_v2(map_to_cta(slots + (rank * WPP + pw) * 32 + 24, lane), __float_as_uint(R_w - N_w), __float_as_uint(M_w),
                                map_to_cta(bar_red2, lane));
                if (pw == 0) {
                    mbar_wait_cta(bar_red2, ph2);
                    if (lane == 0) mbar_expect_tx(bar_red2, RED2_BYTES);
                    float rr = 0.f, qq = 0.f;
#pragma unroll
                    for (int i = 0; i < NSLOT / 32; ++i) {
                        const float4 sb = slot_v[2 * (lane + 32 * i) + 1];
                        rr += sb.z;
                        qq += sb.w;
                    }
                    rr = warp_sum(rr);
                    qq = warp_sum(qq);
                    if (lane == 0) {
                        bcast[2] = qq;
                        if (rank == 0) p.rowaux[r] = make_float4(lse2_t, lse2_s, qq, rr);
                    }
                }
                named_bar_sync(bar2, WPP * 32);
                Q = bcast[2];
                ph2 ^= 1;
            }
        }

        // ================= pass C: gradient, written straight to HBM ===========================================
        if (want_grad) {
            const float gQ = g * Q;
            const bool fold = (g > 0.f) && (gQ > 0.f) && (gQ < INFINITY);
            // fold: g*Q rides in the exponent, 2^(a + log2 gQ) - g*q;  otherwise gQ * 2^a - g*q
            const float lgQ = fold ? __log2f(gQ) : 0.f;
            const u64 nls = pk(lgQ - lse2_s, lgQ - lse2_s);
            const float thrC = thr2 + lgQ;
            const float gq = REF ? g : g * ex2(mtk - lse2_t);        // PLAIN: the parked values are relative to mt_w
            const u64 negg = pk(-gq, -gq), gQp = pk(gQ, gQ);
            T* grow = static_cast<T*>(p.grad) + r * p.ld_g + (int64_t)g_v0 * 8;
            auto sweepC = [&](auto fold_tag, auto mask_tag) {
                constexpr bool FOLD = decltype(fold_tag)::value;
                constexpr bool MASK = decltype(mask_tag)::value;
                auto c_body = [&](int j, bool act, auto tail_tag) {
                    uint32_t qb[8];
                    tmem_ld8(tq + 8 * j, qb);
                    const int v = lane + 32 * j;
                    uint4 rs = make_uint4(0, 0, 0, 0);
                    if (act) rs = load_s(v);
                    tmem_wait_ld();
                    if (act) {
                        const uint32_t ws[4] = {rs.x, rs.y, rs.z, rs.w};
                        uint32_t out[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float a_lo, a_hi;
                            upk(fma2(Pair<T>::widen(ws[i]), k2p, nls), a_lo, a_hi);
                            const u64 sp2 = pk(ex2(a_lo), ex2(a_hi));
                            u64 q2 = pk(__uint_as_float(qb[2 * i]), __uint_as_float(qb[2 * i + 1]));
                            if constexpr (MASK) q2 = mul2(q2, pk(set_ge(a_lo, thrC), set_ge(a_hi, thrC)));
                            u64 o2;
                            if constexpr (FOLD) o2 = fma2(q2, negg, sp2);
                            else o2 = fma2(sp2, gQp, mul2(q2, negg));
                            float n_lo, n_hi;
                            upk(o2, n_lo, n_hi);
                            out[i] = Pair<T>::narrow(n_lo, n_hi);
                        }
                        stg_stream(grow + (int64_t)v * 8, make_uint4(out[0], out[1], out[2], out[3]));
                    }
                    if constexpr (decltype(tail_tag)::value) __syncwarp();
                };
#pragma unroll 2
                for (int j = 0; j < n_full; ++j) c_body(j, true, std::false_type{});
                if (n_tail) c_body(n_full, lane < n_tail, std::true_type{});
            };
            using Yes = std::true_type;
            using No = std::false_type;
            if (fold) {
                if (all_kept) sweepC(Yes{}, No{});
                else sweepC(Yes{}, Yes{});
            } else {
                if (all_kept) sweepC(No{}, No{});
                else sweepC(No{}, Yes{});
            }
        }
        // ---- refill the warp's private slice with its next row (all lanes are done reading it) -------------------
        __syncwarp();
        if (next_r < p.n_rows) issue_load(next_r);
        r = next_r;
        if (DYN) next_r = after_next;
    }
    tmem_fence_before_sync();
    cluster_sync_all();     // nobody leaves while peers may still push into its shared memory
    tmem_fence_after_sync();
    if (warp == 1) tmem_dealloc_512(tmem_base);
}

// --------------------------------------------------------------------------------------------------------
template <int VARIANT, typename T, bool DYN>
int launch_fused_d(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    constexpr int CS = 8, NPIPE = 2;
    auto kfn = row_emit_fused_kernel<VARIANT, T, CS, NPIPE, DYN>;
    const int nvec = e.V / 8;
    const int vpc = (nvec + CS - 1) / CS;
    const FusedPlan<VARIANT, CS, NPIPE> plan(vpc);
    constexpr int WPP = kWarps / NPIPE;
    const int w_chunk = ((vpc + WPP * 32 - 1) / (WPP * 32)) * 32;
    if (w_chunk / 32 > kMaxIters) return KD_ERR_UNSUPPORTED;        // the thread's q cells must fit its 128 TMEM words
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
    // the whole opt-in budget is requested so that exactly one CTA (which owns all 512 TMEM columns) lives on an SM
    cfg.dynamicSmemBytes = max_smem;
    cfg.stream = st;
    if (configured_dev != dev) {
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(fused kernel smem)");
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
    FusedParams cp;
    cp.e = e;
    cp.n_clusters = (int)(want < max_clusters ? want : max_clusters);
    cp.vpc = vpc;
    cp.nvec = nvec;
    cfg.gridDim = dim3(cp.n_clusters * CS);
    if (launched) *launched = cp.n_clusters;
    if (first_free) *first_free = 2 * cp.n_clusters * NPIPE;
    err = cudaLaunchKernelEx(&cfg, kfn, cp);
    if (err != cudaSuccess) return cuda_fail(err, "row_emit_fused_kernel launch");
    return KD_OK;
}

template <int VARIANT, typename T>
int launch_fused_v(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    if (e.row_queue) return launch_fused_d<VARIANT, T, true>(e, st, query_only, capacity, launched, first_free);
    return launch_fused_d<VARIANT, T, false>(e, st, query_only, capacity, launched, first_free);
}

template <typename T>
int launch_fused_t(const EmitParams& e, int variant, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    if (variant == KD_LOCA_REFERENCE) return launch_fused_v<KD_LOCA_REFERENCE, T>(e, st, query_only, capacity, launched, first_free);
    return launch_fused_v<KD_KL_PLAIN, T>(e, st, query_only, capacity, launched, first_free);
}

int fused_mode() {       // KD_FUSED=0 keeps every shape on the three-sweep cluster kernel (A/B comparisons)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KD_FUSED");
        v = e ? atoi(e) : 1;
    }
    return v;
}

}  // namespace

bool fused_applies(const EmitParams& p, int variant, int dtype) {
    if (!fused_mode()) return false;
    if (variant == KD_LOCA_REFERENCE) {
        if (!(p.alpha >= 0.f && p.alpha <= 1.f)) return false;      // overrides outside [0, 1]: exact per-cell path
    } else if (variant != KD_KL_PLAIN) {
        return false;
    }
    return dtype == KD_DTYPE_BF16 || dtype == KD_DTYPE_F16;
}

int launch_emit_fused(const EmitParams& p, int variant, int dtype, cudaStream_t stream, bool query_only, int* capacity,
                      int* clusters_launched, int* first_free_row) {
    if (dtype == KD_DTYPE_BF16) return launch_fused_t<__nv_bfloat16>(p, variant, stream, query_only, capacity, clusters_launched, first_free_row);
    return launch_fused_t<__half>(p, variant, stream, query_only, capacity, clusters_launched, first_free_row);
}

}  // namespace kd
