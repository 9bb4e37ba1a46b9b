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
#include <stdio.h>
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
// exponentials evaluated on the FMA pipe instead of the XU pipe, per 8-cell vector (0-4 pairs each)
#ifndef KD_POLY_AB
#define KD_POLY_AB 0        // student pairs of pass AB (degree 4)
#endif
#ifndef KD_POLY_C
#define KD_POLY_C 0         // student pairs of pass C (degree 3: only the bf16 gradient depends on it)
#endif
constexpr float kPolyFloor = -120.f;
// two-pipeline kernel: the student exponentials of pass AB are parked next to q as fp16 pairs scaled by 2^15 (words 80..119 of
// the thread's tensor memory), so pass C of an unclamped row needs no MUFU at all
#ifndef KD_ECACHE
#define KD_ECACHE 1
#endif
constexpr int kTmE = 80;
constexpr float kEScaleLog2 = 15.f;    // the polynomial path needs its argument >= -126

// Optional phase tracing (compile with -DKD_TRACE): lane 0 of every warp stamps %globaltimer at the phase boundaries of
// each row into `trace` ([cta][warp][row_iter][8] u64), decoded by tools/trace_emit.py.
#ifdef KD_TRACE
#define KD_STAMP(i) do { if (lane == 0 && cp.trace && row_iter < 64) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
    cp.trace[(((size_t)blockIdx.x * kWarps + warp) * 64 + row_iter) * 8 + (i)] = t_; } } while (0)
#else
#define KD_STAMP(i) do {} while (0)
#endif

struct FusedParams {
    unsigned long long* trace;
    EmitParams e;
    int n_clusters;
    int vpc;                      // 16-byte vectors per CTA slice
    int nvec;                     // 16-byte vectors per row
    int tok_after;                // two-pipeline kernel: hand the pass-AB token over this many iterations early (KD_TOKEN_EARLY)
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
        off += (kWarps + 3 * NPIPE) * 8;         // full[16], red1/red2 per pipeline, one pass-AB token per pipeline
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
// one-instruction warp max / min of floats through an order-preserving integer key (redux.sync); NaN is not ordered
__device__ __forceinline__ int fkey(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
__device__ __forceinline__ float warp_max_redux(float v) { return fkey_inv(__reduce_max_sync(0xffffffffu, fkey(v))); }
__device__ __forceinline__ float warp_min_redux(float v) { return fkey_inv(__reduce_min_sync(0xffffffffu, fkey(v))); }
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int VARIANT, typename T, int CS, int NPIPE, bool DYN>
__global__ void __launch_bounds__(kThreads, 1)
row_emit_fused_kernel(const FusedParams cp) {
    constexpr bool TAB = (VARIANT == KD_LOCA_REFERENCE);            // shared override table (phase A + build_table)
    constexpr bool PER_ROW = (VARIANT == KD_LOCA_PER_ROW);          // the row's own label / runner-up cells only
    constexpr bool LOCA = TAB || PER_ROW;                           // teacher statistics and records come from phase A
    constexpr int WPP = kWarps / NPIPE;
    constexpr int NSLOT = CS * WPP;
    constexpr uint32_t RED1_BYTES = NSLOT * (LOCA ? 24 : 32);
    constexpr uint32_t RED2_BYTES = NSLOT * 8;
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
    // Ping-pong of the two pipelines: left alone they fall into lockstep (measured: both in pass AB, then both idle on their
    // reductions, then both bursting stores).  A pipeline may start pass AB only when the other one has finished its own, so
    // one pipeline's XU-bound sweep overlaps the other's reduction wait, pass C, refill and pass 0.
    const uint32_t tok_mine = s_base + plan.bars + (kWarps + 2 * NPIPE + pipe) * 8;
    const uint32_t tok_other = s_base + plan.bars + (kWarps + 2 * NPIPE + (NPIPE - 1 - pipe)) * 8;
    volatile float* done_other = reinterpret_cast<volatile float*>(smem + plan.bcast + (NPIPE - 1 - pipe) * 32) + 7;
    bool other_done = false;

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
        for (int i = 0; i < NPIPE; ++i) {
            mbar_init(s_base + plan.bars + (kWarps + 2 * NPIPE + i) * 8, WPP);      // one arrival per warp of the other pipeline
            reinterpret_cast<volatile float*>(smem + plan.bcast + i * 32)[7] = 0.f;     // "this pipeline has left its row loop"
        }
        fence_mbar_init();
        for (int i = 0; i < NPIPE; ++i) {        // arm the first phase of every reduction barrier
            mbar_expect_tx(s_base + plan.bars + (kWarps + 2 * i) * 8, RED1_BYTES);
            mbar_expect_tx(s_base + plan.bars + (kWarps + 2 * i + 1) * 8, RED2_BYTES);
        }
    }
    if (warp == 1) tmem_alloc_512(s_base + plan.tmem_word);
    // LoCa: encode this CTA's table slice once (0 = not overridden, else max(ov, floor) > 0); layout as in logit_kd_cluster.cu
    float poison = 0.f;
    if constexpr (TAB) {
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
    // the refill of a warp's private slices is split: the teacher part can go as soon as pass AB is done with it, the
    // student part once no later pass of the row will read it (both complete on the warp's mbarrier, armed by the first)
    auto issue_teacher = [&](int64_t row) {
        if (lane == 0 && w_n > 0) {
            const uint32_t bytes = (uint32_t)w_n * 16u;
            fence_proxy_async();
            mbar_expect_tx(bar_full, 2u * bytes);
            tma_load_1d(s_t, tbase + row * p.ld_t + (int64_t)g_v0 * 8, bytes, bar_full);
        }
    };
    auto issue_student = [&](int64_t row) {
        if (lane == 0 && w_n > 0) {
            const uint32_t bytes = (uint32_t)w_n * 16u;
            fence_proxy_async();
            tma_load_1d(s_s, sbase + row * p.ld_s + (int64_t)g_v0 * 8, bytes, bar_full);
        }
    };
    auto issue_load = [&](int64_t row) {
        issue_teacher(row);
        issue_student(row);
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
    int row_iter = -1;
    while (r < p.n_rows) {
        ++row_iter;
        KD_STAMP(0);
        if (!DYN) next_r = r + row_stride;
        int claimed = 0;
        if (DYN && rank == 0 && pw == 0 && lane == 0) claimed = atomicAdd(p.row_queue, 1);
        float lse2_t = 0.f;
        if constexpr (LOCA) lse2_t = p.lse2_t_in[r];
        kd_token_record rec_r;
        rec_r.label = -1; rec_r.k = -1; rec_r.v1 = NaN; rec_r.v2 = NaN;
        if constexpr (PER_ROW) rec_r = p.records[r];
        warm_l2(next_r);
        const uint32_t epoch = (uint32_t)r;
        auto load_s = [&](int v) -> uint4 { return lds128_ep(s_s + v * 16, epoch); };
        auto load_t = [&](int v) -> uint4 { return lds128_ep(s_t + v * 16, epoch); };
        if (w_n > 0) mbar_wait_cta(bar_full, ph_full);
        ph_full ^= 1;
        KD_STAMP(1);

        // ================= pass 0: reference maxima (and the student minimum) of the warp's slice ==============
        float m_w, min_w = INFINITY, mt_w = -INFINITY;
        {
            uint32_t smax2 = Pair<T>::kNegInf2, smin2 = Pair<T>::kPosInf2, tmax2 = Pair<T>::kNegInf2;
#pragma unroll 4
            for (int v = lane; v < w_n; v += 32) {
                const uint4 rs = load_s(v);
                smax2 = Pair<T>::vmax(smax2, Pair<T>::vmax(Pair<T>::vmax(rs.x, rs.y), Pair<T>::vmax(rs.z, rs.w)));
                if constexpr (LOCA) {
                    smin2 = Pair<T>::vmin(smin2, Pair<T>::vmin(Pair<T>::vmin(rs.x, rs.y), Pair<T>::vmin(rs.z, rs.w)));
                } else {
                    const uint4 rt = load_t(v);
                    tmax2 = Pair<T>::vmax(tmax2, Pair<T>::vmax(Pair<T>::vmax(rt.x, rt.y), Pair<T>::vmax(rt.z, rt.w)));
                }
            }
            // (student max, -student min) [LoCa] or (student max, teacher max) [plain KL] ride one packed butterfly
            smax2 = Pair<T>::vmax(smax2, __byte_perm(smax2, smax2, 0x1032));                 // both halves: the lane's maximum
            uint32_t other;
            if constexpr (LOCA) other = Pair<T>::vmin(smin2, __byte_perm(smin2, smin2, 0x1032)) ^ 0x80008000u;      // -min (exact)
            else other = Pair<T>::vmax(tmax2, __byte_perm(tmax2, tmax2, 0x1032));
            uint32_t both = __byte_perm(smax2, other, 0x7610);                                // low half: student max, high half: the other
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) both = Pair<T>::vmax(both, __shfl_xor_sync(0xffffffffu, both, o));
            float lo, hi;
            upk(Pair<T>::widen(both), lo, hi);
            m_w = lo;
            if constexpr (LOCA) min_w = -hi;
            else mt_w = hi;
        }
        KD_STAMP(2);
        if constexpr (NPIPE == 2) {
            // pipeline 1 runs its n-th pass AB after pipeline 0's n-th, pipeline 0 its n-th after pipeline 1's (n-1)-th
            const int need = (pipe == 1) ? row_iter : row_iter - 1;
            if (need >= 0 && !other_done) {
                while (!mbar_try_wait(tok_mine, (uint32_t)need & 1u)) {
                    if (*done_other != 0.f) { other_done = true; break; }
                }
            }
        }
        // exponent references (a warp whose cells are all -inf, or that has no cells, uses 0)
        const float mk = __fmul_rn((m_w == -INFINITY) ? 0.f : m_w, k2);
        const float mtk = __fmul_rn((mt_w == -INFINITY) ? 0.f : mt_w, k2);

        // ================= pass AB: statistics + accumulation in one sweep, q parked in tensor memory ==========
        float S_w, Q_w, R_w, P_w;
        {
            constexpr bool ECACHE = KD_ECACHE;        // every variant of this kernel parks its student exponentials
            // with the exponent cache the student argument carries +15 (e is parked as fp16(e * 2^15)); S_w and P_w are corrected below
            const float esh = ECACHE ? kEScaleLog2 : 0.f;
            const u64 nmk = pk(esh - mk, esh - mk);
            const float tref = LOCA ? lse2_t : mtk;
            const u64 nlt = pk(-tref, -tref);
            u64 accS0 = pk(0.f, 0.f), accS1 = pk(0.f, 0.f), accQ = pk(0.f, 0.f), accR = pk(0.f, 0.f), accP = pk(0.f, 0.f);
            auto ab_body = [&](int j, bool act, auto tail_tag, auto poly_tag) {
                constexpr bool POLY = decltype(poly_tag)::value;
                uint32_t qb[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                uint32_t eb[4] = {0u, 0u, 0u, 0u};
                if (act) {
                    const int v = lane + 32 * j;
                    const uint4 rs = load_s(v), rt = load_t(v);
                    uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
                    if constexpr (TAB) { o0 = lds128_ep(s_ov + v * 16, 0u); o1 = lds128_ep(s_ov + v * 16 + cp.vpc * 16, 0u); }
                    const uint32_t ws[4] = {rs.x, rs.y, rs.z, rs.w}, wt[4] = {rt.x, rt.y, rt.z, rt.w};
                    const u64 ovp[4] = {pk(__uint_as_float(o0.x), __uint_as_float(o0.y)), pk(__uint_as_float(o0.z), __uint_as_float(o0.w)),
                                        pk(__uint_as_float(o1.x), __uint_as_float(o1.y)), pk(__uint_as_float(o1.z), __uint_as_float(o1.w))};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const u64 a2 = fma2(Pair<T>::widen(ws[i]), k2p, nmk);        // xs*k2 - m_w*k2  (<= 0)
                        const u64 e2 = (POLY && i >= 4 - KD_POLY_AB) ? ex2_poly2<4>(a2) : ex2_mufu2(a2);
                        if constexpr (ECACHE) {
                            float e_lo2, e_hi2;
                            upk(e2, e_lo2, e_hi2);
                            const __half2 h = __floats2half2_rn(e_lo2, e_hi2);
                            eb[i] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        if (i & 1) accS1 = add2(accS1, e2);
                        else accS0 = add2(accS0, e2);
                        const u64 b2 = fma2(Pair<T>::widen(wt[i]), k2p, nlt);        // log2 tp (LOCA) / relative to mt_w (PLAIN)
                        float e_lo, e_hi;
                        if constexpr (TAB) upk(fma2(ovp[i], pushp, b2), e_lo, e_hi);
                        else upk(b2, e_lo, e_hi);
                        const u64 t2 = pk(ex2(e_lo), ex2(e_hi));
                        const u64 q2 = TAB ? add2(t2, ovp[i]) : t2;
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
                if constexpr (ECACHE) tmem_st4(tq + kTmE + 4 * j, eb);
            };
            // the FMA-pipe exponentials need xs*k2 - m_w*k2 >= -126 for every cell of the warp's slice
            const bool poly_ab = LOCA && KD_POLY_AB > 0 && (fmaf(min_w, k2, esh - mk) >= kPolyFloor);
            bool token_given = false;
            if (poly_ab) {
#pragma unroll 2
                for (int j = 0; j < n_full; ++j) ab_body(j, true, std::false_type{}, std::true_type{});
                if (n_tail) ab_body(n_full, lane < n_tail, std::true_type{}, std::true_type{});
            } else {
                // the other pipeline may start its pass AB once this warp has `tok_after` iterations left (0 = strict alternation)
                const int j_split = max(0, n_full - cp.tok_after);
#pragma unroll 2
                for (int j = 0; j < j_split; ++j) ab_body(j, true, std::false_type{}, std::false_type{});
                if (NPIPE == 2 && cp.tok_after > 0) {
                    if (lane == 0) mbar_arrive_cta(tok_other);
                    token_given = true;
                }
#pragma unroll 2
                for (int j = j_split; j < n_full; ++j) ab_body(j, true, std::false_type{}, std::false_type{});
                if (n_tail) ab_body(n_full, lane < n_tail, std::true_type{}, std::false_type{});
            }
            tmem_wait_st();
            if constexpr (!PER_ROW) {
                __syncwarp();
                if (next_r < p.n_rows) issue_teacher(next_r);      // nobody reads the teacher slice after pass AB
            }
            if (NPIPE == 2 && !token_given && lane == 0) mbar_arrive_cta(tok_other);
            float x0, x1;
            upk(add2(accS0, accS1), x0, x1);
            S_w = warp_sum(x0 + x1) * (ECACHE ? 1.f / 32768.f : 1.f);
            float q_l, r_l, p_l;                       // this lane's shares of Q_w, R_w, P_w
            upk(accQ, x0, x1);
            q_l = x0 + x1;
            upk(accR, x0, x1);
            r_l = x0 + x1;
            upk(accP, x0, x1);
            p_l = x0 + x1;
            if constexpr (PER_ROW) {
                // The sweep treated every cell as q = tp.  The row's own override cells (label -> v1, then runner-up -> v2,
                // P1:184-185 applied per row) are patched now: the owning lane corrects its sums and the parked q.
                if (rec_r.label >= 0) {
                    const bool same = (rec_r.k == rec_r.label);
#pragma unroll
                    for (int which = 0; which < 2; ++which) {
                        const int c = which ? rec_r.k : rec_r.label;
                        const float ov = which ? rec_r.v2 : rec_r.v1;
                        const int v = (c >> 3) - g_v0;                        // warp-uniform
                        if ((which == 1 || !same) && c >= 0 && v >= 0 && v < w_n) {
                            const int j = v >> 5, cell = c & 7;
                            uint32_t qb[8];
                            tmem_ld8(tq + 8 * j, qb);
                            tmem_wait_ld();
                            if (lane == (v & 31)) {
                                const float xs = Elem<T>::to_float(*reinterpret_cast<const T*>(smem + (s_s - s_base) + v * 16 + cell * 2));
                                const float xt = Elem<T>::to_float(*reinterpret_cast<const T*>(smem + (s_t - s_base) + v * 16 + cell * 2));
                                const float b = fmaf(xt, k2, -lse2_t), t = ex2(b);
                                const float a = fmaf(xs, k2, esh - mk);
                                q_l += ov - t;
                                r_l -= t * b;
                                p_l = fmaf(ov - t, a, p_l);
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (i == cell) qb[i] = __float_as_uint(ov);
                            }
                            __syncwarp();
                            tmem_st8(tq + 8 * j, qb);
                            tmem_wait_st();
                        }
                    }
                }
            }
            Q_w = warp_sum(q_l);
            R_w = warp_sum(r_l) + poison;
            P_w = warp_sum(p_l) - esh * Q_w;          // sum q*(a + 15) -> sum q*a
        }
        if constexpr (PER_ROW) {
            __syncwarp();
            if (next_r < p.n_rows) issue_teacher(next_r);          // after the override fix-up, which still reads the teacher slice
        }
        KD_STAMP(3);
        // ---- the single reduction: push the warp's partial to every CTA of the cluster ---------------------------
        if (DYN) claimed = __shfl_sync(0xffffffffu, claimed, 0);
        if (lane < CS) {
            const uint32_t dst = map_to_cta(slots + (rank * WPP + pw) * 32, lane);
            const uint32_t rbar = map_to_cta(bar_red1, lane);
            st_async_v4(dst, __float_as_uint(m_w), __float_as_uint(S_w), __float_as_uint(Q_w), __float_as_uint(R_w - P_w), rbar);
            if constexpr (LOCA) st_async_v2(dst + 16, __float_as_uint(min_w), (uint32_t)claimed, rbar);
            else st_async_v4(dst + 16, 0u, (uint32_t)claimed, __float_as_uint(mt_w), 0u, rbar);
        }
        // xlogy(v, v) of the row's own override cells (log2 units); a bad label poisons the row like the reference's gather would raise
        float own_xlogy = 0.f;                  // computed before the reduction wait: it depends on the token record only
        if constexpr (PER_ROW) {
            if (rec_r.label >= 0) {
                if (rec_r.k != rec_r.label && rec_r.v1 != 0.f) own_xlogy += rec_r.v1 * __log2f(rec_r.v1);
                if (rec_r.v2 != 0.f) own_xlogy += rec_r.v2 * __log2f(rec_r.v2);
            } else {
                own_xlogy = NaN;
            }
        }
        // one warp per pipeline folds the NSLOT partials and broadcasts the row statistics through shared memory
        if (pw == 0) {
            mbar_wait_cta(bar_red1, ph1);
            KD_STAMP(7);
            if (lane == 0) mbar_expect_tx(bar_red1, RED1_BYTES);      // arm the next row's phase
            constexpr int SPL = (NSLOT + 31) / 32;                     // slots per lane (the last ones may not exist: neutral values)
            float4 sa[SPL], sb[SPL];
            float m_s = -INFINITY, m_t = -INFINITY, mn = INFINITY;
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const int idx = lane + 32 * i;
                sa[i] = make_float4(-INFINITY, 0.f, 0.f, 0.f);
                sb[i] = make_float4(INFINITY, 0.f, -INFINITY, 0.f);
                if (NSLOT % 32 == 0 || idx < NSLOT) {
                    sa[i] = slot_v[2 * idx];
                    sb[i] = slot_v[2 * idx + 1];
                }
                m_s = fmaxf(m_s, sa[i].x);
                if constexpr (LOCA) mn = fminf(mn, sb[i].x);
                else m_t = fmaxf(m_t, sb[i].z);
            }
            m_s = warp_max_redux(m_s);
            if constexpr (LOCA) mn = warp_min_redux(mn);
            else m_t = warp_max_redux(m_t);
            float s_sum = 0.f, t_sum = 0.f, Q = 0.f, A = 0.f;
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                if (sa[i].x != -INFINITY) s_sum = fmaf(sa[i].y, ex2((sa[i].x - m_s) * k2), s_sum);
                if constexpr (LOCA) {
                    const float mki = __fmul_rn((sa[i].x == -INFINITY) ? 0.f : sa[i].x, k2);
                    Q += sa[i].z;
                    A += sa[i].w - mki * sa[i].z;            // row = A + lse2_s * Q
                } else {
                    if (sb[i].z != -INFINITY) t_sum = fmaf(sa[i].z, ex2((sb[i].z - m_t) * k2), t_sum);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s_sum += __shfl_xor_sync(0xffffffffu, s_sum, o);
                if constexpr (LOCA) {
                    Q += __shfl_xor_sync(0xffffffffu, Q, o);
                    A += __shfl_xor_sync(0xffffffffu, A, o);
                } else {
                    t_sum += __shfl_xor_sync(0xffffffffu, t_sum, o);
                }
            }
            const float lse2_s = fmaf(m_s, k2, __log2f(s_sum));
            float row;
            if constexpr (LOCA) {
                row = fmaf(lse2_s, Q, A);
            } else {
                lse2_t = fmaf(m_t, k2, __log2f(t_sum));
                row = 0.f;
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const float mki = __fmul_rn((sa[i].x == -INFINITY) ? 0.f : sa[i].x, k2);
                    const float mti = __fmul_rn((sb[i].z == -INFINITY) ? 0.f : sb[i].z, k2);
                    const float sc = ex2(mti - lse2_t);               // this warp's teacher partials are relative to its mt_w
                    Q = fmaf(sc, sa[i].z, Q);
                    row = fmaf(sc, sa[i].w + ((mti - lse2_t) - (mki - lse2_s)) * sa[i].z, row);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    Q += __shfl_xor_sync(0xffffffffu, Q, o);
                    row += __shfl_xor_sync(0xffffffffu, row, o);
                }
            }
            int claim_all = 0;
            if (DYN) claim_all = __shfl_sync(0xffffffffu, __float_as_int(sb[0].y), 0);     // slot 0 = (rank 0, warp 0)
            bool all_kept = true;
            if constexpr (LOCA) {
                const float a_min = fmaf(mn, k2, -lse2_s);
                all_kept = a_min >= thr2 + 1e-3f * (1.f + fabsf(thr2));       // NaN -> masked path
            }
            if (lane == 0) {
                bcast[0] = lse2_s;
                bcast[1] = lse2_t;
                bcast[2] = Q;
                bcast[3] = all_kept ? 1.f : 0.f;
                bcast[4] = __int_as_float(claim_all);
                if (rank == 0 && all_kept) p.rowaux[r] = make_float4(lse2_t, lse2_s, Q, row + own_xlogy);
            }
        }
        named_bar_sync(bar1, WPP * 32);
        KD_STAMP(4);
        const float lse2_s = bcast[0];
        if constexpr (!LOCA) lse2_t = bcast[1];
        float Q = bcast[2];
        const bool all_kept = bcast[3] != 0.f;
        int64_t after_next = 0;
        if (DYN) after_next = __float_as_int(bcast[4]);
        ph1 ^= 1;

        // ================= pass B' (rare): exact masked sums for rows the sp >= eps clamp touches =================
        // With the exponent cache the exact mask is applied to the parked q right here (q <- q*[a >= thr2], written back to
        // tensor memory), so the gradient pass of a clamped row is the same cached, MUFU-free sweep as for an unclamped row and
        // the student slice can be re-armed before the second reduction.
        bool student_refilled = false;
        bool q_masked = false;
        if constexpr (LOCA) {
            if (!all_kept) {
                constexpr bool WRITE_BACK = LOCA && KD_ECACHE;
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
                            // max(a, thr2) = a*m + thr2*(1 - m):  sum q*max(a, thr2) = sum (q*m)*a + thr2 * (Q_w - sum q*m)
                            const u64 a2 = pk(a_lo, a_hi);
                            const u64 qm2 = mul2(q2, pk(set_ge(a_lo, thr2), set_ge(a_hi, thr2)));
                            accN = fma2(qm2, a2, accN);
                            accM = add2(accM, qm2);
                            if constexpr (WRITE_BACK) {
                                float m_lo, m_hi;
                                upk(qm2, m_lo, m_hi);
                                qb[2 * i] = __float_as_uint(m_lo);
                                qb[2 * i + 1] = __float_as_uint(m_hi);
                            }
                        }
                    }
                    if constexpr (decltype(tail_tag)::value) __syncwarp();
                    if constexpr (WRITE_BACK) tmem_st8(tq + 8 * j, qb);
                };
#pragma unroll 2
                for (int j = 0; j < n_full; ++j) bp_body(j, true, std::false_type{});
                if (n_tail) bp_body(n_full, lane < n_tail, std::true_type{});
                if constexpr (WRITE_BACK) {
                    tmem_wait_st();
                    q_masked = true;
                    if (want_grad) {            // nothing reads the student slice any more: the cached gradient sweep will run
                        __syncwarp();
                        if (next_r < p.n_rows) issue_student(next_r);
                        student_refilled = true;
                    }
                }
                float x0, x1;
                upk(accM, x0, x1);
                const float M_w = warp_sum(x0 + x1);
                upk(accN, x0, x1);
                const float N_w = warp_sum(x0 + x1) + thr2 * (Q_w - M_w);
                if (lane < CS)
                    st_async_v2(map_to_cta(slots + (rank * WPP + pw) * 32 + 24, lane), __float_as_uint(R_w - N_w), __float_as_uint(M_w),
                                map_to_cta(bar_red2, lane));
                if (pw == 0) {
                    mbar_wait_cta(bar_red2, ph2);
                    if (lane == 0) mbar_expect_tx(bar_red2, RED2_BYTES);
                    float rr = 0.f, qq = 0.f;
#pragma unroll
                    for (int i = 0; i < (NSLOT + 31) / 32; ++i) {
                        const int idx = lane + 32 * i;
                        if (NSLOT % 32 == 0 || idx < NSLOT) {
                            const float4 sb = slot_v[2 * idx + 1];
                            rr += sb.z;
                            qq += sb.w;
                        }
                    }
                    rr = warp_sum(rr);
                    qq = warp_sum(qq);
                    if (lane == 0) {
                        bcast[2] = qq;
                        if (rank == 0) p.rowaux[r] = make_float4(lse2_t, lse2_s, qq, rr + own_xlogy);
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
            const float gq = LOCA ? g : g * ex2(mtk - lse2_t);        // PLAIN: the parked values are relative to mt_w
            const u64 negg = pk(-gq, -gq), gQp = pk(gQ, gQ);
            T* grow = static_cast<T*>(p.grad) + r * p.ld_g + (int64_t)g_v0 * 8;
            auto sweepC = [&](auto fold_tag, auto mask_tag, auto poly_tag) {
                constexpr bool FOLD = decltype(fold_tag)::value;
                constexpr bool MASK = decltype(mask_tag)::value;
                constexpr bool POLY = decltype(poly_tag)::value;
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
                            const u64 a2 = fma2(Pair<T>::widen(ws[i]), k2p, nls);
                            float a_lo, a_hi;
                            upk(a2, a_lo, a_hi);
                            const u64 sp2 = (POLY && i >= 4 - KD_POLY_C) ? ex2_poly2<3>(a2) : ex2_mufu2(a2);
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
            // exponent cache (unclamped rows, and clamped rows whose q was masked by pass B'): no shared memory, no MUFU
            auto sweepC_cached_impl = [&](float ce) {
                const u64 cep = pk(ce, ce);
                auto c_step = [&](int j, auto nv_tag, bool act, auto tail_tag) {
                    constexpr int NV = decltype(nv_tag)::value;       // vectors per tensor-memory wait
                    uint32_t qb[NV][8], eb[NV][4];
#pragma unroll
                    for (int u = 0; u < NV; ++u) {
                        tmem_ld8(tq + 8 * (j + u), qb[u]);
                        tmem_ld4(tq + kTmE + 4 * (j + u), eb[u]);
                    }
                    tmem_wait_ld();
                    if (act) {
#pragma unroll
                        for (int u = 0; u < NV; ++u) {
                            uint32_t out[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 ef = __half22float2(*reinterpret_cast<const __half2*>(&eb[u][i]));
                                const u64 q2 = pk(__uint_as_float(qb[u][2 * i]), __uint_as_float(qb[u][2 * i + 1]));
                                float n_lo, n_hi;
                                upk(fma2(pk(ef.x, ef.y), cep, mul2(q2, negg)), n_lo, n_hi);
                                out[i] = Pair<T>::narrow(n_lo, n_hi);
                            }
                            stg_stream(grow + (int64_t)(lane + 32 * (j + u)) * 8, make_uint4(out[0], out[1], out[2], out[3]));
                        }
                    }
                    if constexpr (decltype(tail_tag)::value) __syncwarp();
                };
                using V4 = std::integral_constant<int, 4>;
                using V2 = std::integral_constant<int, 2>;
                using V1 = std::integral_constant<int, 1>;
                int j = 0;
                for (; j + 3 < n_full; j += 4) c_step(j, V4{}, true, std::false_type{});
                if (j + 1 < n_full) { c_step(j, V2{}, true, std::false_type{}); j += 2; }
                if (j < n_full) c_step(j, V1{}, true, std::false_type{});
                if (n_tail) c_step(n_full, V1{}, lane < n_tail, std::true_type{});
            };
            // g*Q*sp = gQ * e' * 2^(m_w*k2 - 15 - lse2_s): valid for any sign of g and g*Q
            auto sweepC_cached = [&]() { sweepC_cached_impl(gQ * ex2((mk - kEScaleLog2) - lse2_s)); };
            using Yes = std::true_type;
            using No = std::false_type;
            // FMA-pipe exponentials in the common (folded, unmasked) sweep when log2(g*Q*sp) >= -126 for the whole slice
            const bool poly_c = LOCA && KD_POLY_C > 0 && (fmaf(min_w, k2, lgQ - lse2_s) >= kPolyFloor);
            if ((all_kept || q_masked) && KD_ECACHE) {
                if (!student_refilled) {
                    if (next_r < p.n_rows) issue_student(next_r);       // the cached sweep reads tensor memory only
                    student_refilled = true;
                }
                sweepC_cached();
            } else if (fold) {
                if (all_kept) {
                    if (poly_c) sweepC(Yes{}, No{}, Yes{});
                    else sweepC(Yes{}, No{}, No{});
                } else {
                    sweepC(Yes{}, Yes{}, No{});
                }
            } else {
                if (all_kept) sweepC(No{}, No{}, No{});
                else sweepC(No{}, Yes{}, No{});
            }
        }
        KD_STAMP(5);
        // ---- refill the warp's private slice with its next row (all lanes are done reading it) -------------------
        __syncwarp();
        if (!student_refilled && next_r < p.n_rows) issue_student(next_r);
        r = next_r;
        if (DYN) next_r = after_next;
        KD_STAMP(6);
    }
    if (pw == 0 && lane == 0) reinterpret_cast<volatile float*>(smem + plan.bcast + pipe * 32)[7] = 1.f;      // releases a partner waiting for our pass AB
    tmem_fence_before_sync();
    cluster_sync_all();     // nobody leaves while peers may still push into its shared memory
    tmem_fence_after_sync();
    if (warp == 1) tmem_dealloc_512(tmem_base);
}


// ========================================================================================================
// Software-pipelined variant: ONE pipeline of 16 warps per CTA, rows overlapped in time instead of side by side.
//
//   iteration i of a cluster:   [fold(i-1)] -> (B'(i-1), push2) -> pass 0 (i) -> pass AB (i) -> push(i) -> pass C (i-1)
//
// The reduction of row i has the whole of pass C (i-1) and the start of the next iteration to complete, so no warp
// waits for it and all 16 warps of the SM stay busy (the two-pipeline kernel above parks 8 of them at every
// reduction).  Residency: student slices in a ring of 3, teacher slices in a ring of 2 (the teacher is consumed by
// pass AB only); the TMA of row i+1 is issued at the top of iteration i by whichever warp arrives there LAST (every
// warp is then done with the slots being refilled), one bulk copy per tensor for the whole CTA slice.  A thread
// handles at most 5 vectors of a row, so tensor memory holds q double-buffered (2 x 40 words) AND the thread's 40
// override-table entries (each column has exactly one owner thread now), which takes the table out of shared memory.
// Per-row scalars every warp needs (lse2_s, g*Q and its log2, the clamp flag, the next claimed row) are computed once
// by the folding warp — the one with the fewest vectors — and broadcast through shared memory.
constexpr int kMaxIt = 5;
constexpr int kTmQ = 0, kTmTable = 80;     // TMEM word offsets inside the thread's 128 words: q[2][40], table[40]
using Int1 = std::integral_constant<int, 1>;
using Int2 = std::integral_constant<int, 2>;

template <int VARIANT, int CS>
struct PipePlan {
    static constexpr int NSLOT = CS * kWarps;
    int students, teachers, bars, slots, bcast, misc, total;
    __host__ __device__ PipePlan(int vpc) {
        int off = 0;
        students = off;
        off += 3 * vpc * 16;
        teachers = off;
        off += 2 * vpc * 16;
        bars = off;
        off += 10 * 8;                    // full[2], red1[2], red2[2], res1[2], res2[2]
        slots = off;
        off += 2 * NSLOT * 32;            // double-buffered by row parity
        bcast = off;
        off += 2 * 32;
        misc = off;                       // [0] tensor-memory base address, [1] arrival counter of the TMA election
        off += 16;
        total = off;
    }
};

__device__ __forceinline__ void st_async_b32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 :: "r"(cluster_addr), "r"(v), "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ uint32_t swap_halves(uint32_t w) { return __byte_perm(w, w, 0x1032); }

// Sums of four per-lane values in 6 shuffles: afterwards lanes 8g..8g+7 hold the warp total of value g.
__device__ __forceinline__ float warp_sum4_grouped(float a, float b, float c, float d, int lane) {
    const bool hi4 = lane & 16, hi3 = lane & 8;
    float x = hi4 ? c : a, y = hi4 ? d : b;
    const float sx = hi4 ? a : c, sy = hi4 ? b : d;
    x += __shfl_xor_sync(0xffffffffu, sx, 16);
    y += __shfl_xor_sync(0xffffffffu, sy, 16);
    float k = hi3 ? y : x;
    const float sk = hi3 ? x : y;
    k += __shfl_xor_sync(0xffffffffu, sk, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

template <int VARIANT, typename T, int CS, bool DYN>
__global__ void __launch_bounds__(kThreads, 1)
row_emit_pipelined_kernel(const FusedParams cp) {
    constexpr bool REF = (VARIANT == KD_LOCA_REFERENCE);
    constexpr int NSLOT = CS * kWarps;
    constexpr int SPL = NSLOT / 32;                 // slots per lane in the fold
    constexpr uint32_t RED1_BYTES = NSLOT * (REF ? 24 : 32);
    constexpr uint32_t RED2_BYTES = NSLOT * 8;
    constexpr int FOLD_WARP = kWarps - 2;           // a warp with the fewest vectors (the last one also runs the partial step)
    static_assert(CS == 8, "the grouped push below addresses 8 CTAs with 3 lane bits");
    extern __shared__ __align__(128) unsigned char smem[];
    const EmitParams& p = cp.e;
    const PipePlan<VARIANT, CS> plan(cp.vpc);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x / CS;
    const float NaN = __int_as_float(0x7fc00000);

    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar_full = s_base + plan.bars;               // + par * 8
    const uint32_t bar_red1 = s_base + plan.bars + 2 * 8;
    const uint32_t bar_red2 = s_base + plan.bars + 4 * 8;
    const uint32_t bar_res1 = s_base + plan.bars + 6 * 8;
    const uint32_t bar_res2 = s_base + plan.bars + 8 * 8;
    const uint32_t slots = s_base + plan.slots;                 // + par * NSLOT * 32
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem + plan.misc);

    // this CTA's column slice; warp iterations (32 vectors each) are dealt out evenly
    const int cta_v0 = min((int)rank * cp.vpc, cp.nvec);
    const int cta_v1 = min(cta_v0 + cp.vpc, cp.nvec);
    const int cta_n = cta_v1 - cta_v0;
    const int it_total = (cta_n + 31) >> 5;
    const int it_base = it_total / kWarps, it_extra = it_total % kWarps;
    const int w_lo = min((warp * it_base + min(warp, it_extra)) * 32, cta_n);
    const int w_n = max(0, min((it_base + (warp < it_extra ? 1 : 0)) * 32, cta_n - w_lo));
    const int n_full = w_n >> 5, n_tail = w_n & 31;
    const int g_v0 = cta_v0 + w_lo;
    const uint32_t slice_bytes = (uint32_t)cp.vpc * 16u;
    const uint32_t s_stu = s_base + plan.students, s_tea = s_base + plan.teachers;

    if (tid == 0) {
        for (int w = 0; w < 10; ++w) mbar_init(s_base + plan.bars + w * 8, 1);
        misc[1] = 0u;
        fence_mbar_init();
        for (int par = 0; par < 2; ++par) {        // arm the first phase of every reduction barrier
            mbar_expect_tx(bar_red1 + par * 8, RED1_BYTES);
            mbar_expect_tx(bar_red2 + par * 8, RED2_BYTES);
        }
    }
    if (warp == 1) tmem_alloc_512(s_base + plan.misc);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(misc);
    const uint32_t tm = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(kTmemColsPerWarp * (warp >> 2));

    // LoCa: the thread's own table entries go to tensor memory, encoded (0 = not overridden, else max(ov, floor) > 0)
    float poison = 0.f;
    if constexpr (REF) {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < kMaxIt; ++j) {
            uint32_t enc[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            const int v = lane + 32 * j;
            if (v < w_n) {
                const float4* src = reinterpret_cast<const float4*>(p.table + (int64_t)(g_v0 + v) * 8);
                const float4 a = src[0], b = src[1];
                const float ov[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (ov[i] == ov[i]) {
                        ok = ok && (ov[i] >= 0.f) && (ov[i] <= 3.0e38f);
                        enc[i] = __float_as_uint(fmaxf(ov[i], kOvFloor));
                    }
                }
            }
            __syncwarp();
            tmem_st8(tm + kTmTable + 8 * j, enc);
        }
        tmem_wait_st();
        // a negative / non-finite override (alpha > 1; the reference's loss is NaN there too) poisons the row sums
        if (!__syncthreads_and(ok)) poison = NaN;
    }
    cluster_sync_all();     // barriers initialised cluster-wide before any remote push

    const T* tbase = static_cast<const T*>(p.teacher);
    const T* sbase = static_cast<const T*>(p.student);
    // one thread loads the whole CTA slice of `row` (student -> ring slot s_dst, teacher -> t_dst)
    auto issue_load = [&](int64_t row, uint32_t s_dst, uint32_t t_dst, uint32_t bar) {
        const uint32_t bytes = (uint32_t)cta_n * 16u;
        fence_proxy_async();
        mbar_expect_tx(bar, 2u * bytes);
        tma_load_1d(s_dst, sbase + row * p.ld_s + (int64_t)cta_v0 * 8, bytes, bar);
        tma_load_1d(t_dst, tbase + row * p.ld_t + (int64_t)cta_v0 * 8, bytes, bar);
    };

    const float k2 = p.k2, thr2 = p.thr2;
    const float up = p.upstream ? *p.upstream : 1.f;
    const float g = p.gscale * up;
    const bool want_grad = !(p.flags & KD_FLAG_NO_GRAD);
    const u64 k2p = pk(k2, k2);
    const u64 pushp = pk(kPush, kPush);
    uint32_t ph2_0 = 0, ph2_1 = 0;

    // rows of this cluster: the first three are static, later ones are claimed from the queue (the claim made in
    // iteration i travels with reduction i and names the row of iteration i+3)
    const int64_t n_rows = p.n_rows;
    int64_t r_cur = cluster_id, r_nxt = r_cur + cp.n_clusters, r_nn = r_nxt + cp.n_clusters;
    int64_t r_prev = -1;
    // ring slots (byte addresses of the CTA slice): student of the current / previous / next row, teacher current / next
    uint32_t st_cur = s_stu, st_prev = s_stu + 2 * slice_bytes, st_nxt = s_stu + slice_bytes;
    uint32_t te_cur = s_tea, te_nxt = s_tea + slice_bytes;
    float lse2_t_prev = 0.f, mtk_prev = 0.f, R_prev = 0.f, min_prev = 0.f;     // carried from pass AB (i) to pass C (i) one iteration later
    if (tid == 0 && cta_n > 0 && r_cur < n_rows) issue_load(r_cur, st_cur, te_cur, bar_full);

    for (int it = 0;; ++it) {
        const bool have_cur = r_cur < n_rows, have_prev = r_prev >= 0;
        if (!have_cur && !have_prev) break;
        const int par = it & 1, ppar = par ^ 1;                 // parity of the current / previous row
        if (have_cur) {
            if (cta_n > 0) mbar_wait_hint(bar_full + par * 8, (it >> 1) & 1, 2000u);
            // the LAST warp to get here knows every warp has finished iteration it-1, i.e. is done reading the student slot of
            // row it-2 and the teacher slot of row it-1: it refills them with row it+1
            unsigned old = 0;
            if (lane == 0) old = atomicAdd(&misc[1], 1u);
            if (lane == 0 && (old & (kWarps - 1)) == kWarps - 1 && cta_n > 0 && r_nxt < n_rows)
                issue_load(r_nxt, st_nxt, te_nxt, bar_full + ppar * 8);
            __syncwarp();
        }
        int claimed = 0;
        if (DYN && have_cur && rank == 0 && warp == FOLD_WARP && lane == 0) claimed = atomicAdd(p.row_queue, 1);

        // ---------------- result of the previous row's reduction (folded by one warp, long complete) ----------------
        float lse2_s_p = 0.f, gQ_p = 0.f, lgQ_p = 0.f;
        bool kept_p = true, fold_p = false;
        volatile float* bc = reinterpret_cast<volatile float*>(smem + plan.bcast + ppar * 32);
        if (have_prev) {
            const uint32_t ph = ((it - 1) >> 1) & 1;
            if (warp == FOLD_WARP) {
                mbar_wait_hint(bar_red1 + ppar * 8, ph, 1000u);
                if (lane == 0) mbar_expect_tx(bar_red1 + ppar * 8, RED1_BYTES);      // arm this parity's next use
                const float4* slot_v = reinterpret_cast<const float4*>(smem + plan.slots + ppar * NSLOT * 32);
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
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {          // independent butterflies, interleaved
                    m_s = fmaxf(m_s, __shfl_xor_sync(0xffffffffu, m_s, o));
                    if constexpr (REF) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    else m_t = fmaxf(m_t, __shfl_xor_sync(0xffffffffu, m_t, o));
                }
                float s_sum = 0.f, t_sum = 0.f, Q = 0.f, A = 0.f;
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    if (sa[i].x != -INFINITY) s_sum = fmaf(sa[i].y, ex2((sa[i].x - m_s) * k2), s_sum);
                    if constexpr (REF) {
                        const float mki = __fmul_rn((sa[i].x == -INFINITY) ? 0.f : sa[i].x, k2);
                        Q += sa[i].z;
                        A += sa[i].w - mki * sa[i].z;            // row = A + lse2_s * Q
                    } else {
                        if (sb[i].z != -INFINITY) t_sum = fmaf(sa[i].z, ex2((sb[i].z - m_t) * k2), t_sum);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s_sum += __shfl_xor_sync(0xffffffffu, s_sum, o);
                    if constexpr (REF) {
                        Q += __shfl_xor_sync(0xffffffffu, Q, o);
                        A += __shfl_xor_sync(0xffffffffu, A, o);
                    } else {
                        t_sum += __shfl_xor_sync(0xffffffffu, t_sum, o);
                    }
                }
                const float lse2_s = fmaf(m_s, k2, __log2f(s_sum));
                float lse2_t = lse2_t_prev, row;
                if constexpr (REF) {
                    row = fmaf(lse2_s, Q, A);
                } else {
                    lse2_t = fmaf(m_t, k2, __log2f(t_sum));
                    row = 0.f;
#pragma unroll
                    for (int i = 0; i < SPL; ++i) {
                        const float mki = __fmul_rn((sa[i].x == -INFINITY) ? 0.f : sa[i].x, k2);
                        const float mti = __fmul_rn((sb[i].z == -INFINITY) ? 0.f : sb[i].z, k2);
                        const float sc = ex2(mti - lse2_t);           // that warp's teacher partials are relative to its mt_w
                        Q = fmaf(sc, sa[i].z, Q);
                        row = fmaf(sc, sa[i].w + ((mti - lse2_t) - (mki - lse2_s)) * sa[i].z, row);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        Q += __shfl_xor_sync(0xffffffffu, Q, o);
                        row += __shfl_xor_sync(0xffffffffu, row, o);
                    }
                }
                // the leader's slot (cluster rank 0, fold warp) carries the claimed row index
                int claim_all = 0;
                if (DYN) claim_all = __shfl_sync(0xffffffffu, __float_as_int(sb[FOLD_WARP / 32].y), FOLD_WARP % 32);
                bool all_kept = true;
                if constexpr (REF) {
                    const float a_min = fmaf(mn, k2, -lse2_s);
                    all_kept = a_min >= thr2 + 1e-3f * (1.f + fabsf(thr2));       // NaN -> masked path
                }
                if (lane == 0) {
                    const float gQ = g * Q;
                    const bool fold = (g > 0.f) && (gQ > 0.f) && (gQ < INFINITY);
                    bc[0] = lse2_s;
                    bc[1] = lse2_t;
                    bc[2] = gQ;
                    bc[3] = __int_as_float((all_kept ? 1 : 0) | (fold ? 2 : 0));
                    bc[4] = __int_as_float(claim_all);
                    bc[5] = fold ? __log2f(gQ) : 0.f;
                    if (rank == 0 && all_kept) p.rowaux[r_prev] = make_float4(lse2_t, lse2_s, Q, row);
                    mbar_arrive_cta(bar_res1 + ppar * 8);
                }
            }
        }
        if (!DYN) r_nn = r_nxt + cp.n_clusters;
        const uint32_t s_sp = st_prev + w_lo * 16;
        auto load_sp = [&](int v) -> uint4 { return lds128_ep(s_sp + v * 16, (uint32_t)it); };

        // ---------------- current row: pass 0, pass AB, push ----------------
        float R_cur = 0.f, mtk_cur = 0.f, lse2_t_cur = 0.f;
        float red = 0.f, m_w = -INFINITY, min_w = INFINITY, mt_w = -INFINITY;
        if (have_cur) {
            if constexpr (REF) lse2_t_cur = p.lse2_t_in[r_cur];
            const uint32_t epoch = (uint32_t)it;
            const uint32_t s_s = st_cur + w_lo * 16, s_t = te_cur + w_lo * 16;
            auto load_s = [&](int v) -> uint4 { return lds128_ep(s_s + v * 16, epoch); };
            auto load_t = [&](int v) -> uint4 { return lds128_ep(s_t + v * 16, epoch); };
            // pass 0: (student max, -student min) [LoCa] or (student max, teacher max) [plain KL] ride one packed butterfly
            {
                uint32_t smax2 = Pair<T>::kNegInf2, smin2 = Pair<T>::kPosInf2, tmax2 = Pair<T>::kNegInf2;
#pragma unroll
                for (int j = 0; j < kMaxIt; ++j) {
                    const int v = lane + 32 * j;
                    if (v < w_n) {
                        const uint4 rs = load_s(v);
                        smax2 = Pair<T>::vmax(smax2, Pair<T>::vmax(Pair<T>::vmax(rs.x, rs.y), Pair<T>::vmax(rs.z, rs.w)));
                        if constexpr (REF) {
                            smin2 = Pair<T>::vmin(smin2, Pair<T>::vmin(Pair<T>::vmin(rs.x, rs.y), Pair<T>::vmin(rs.z, rs.w)));
                        } else {
                            const uint4 rt = load_t(v);
                            tmax2 = Pair<T>::vmax(tmax2, Pair<T>::vmax(Pair<T>::vmax(rt.x, rt.y), Pair<T>::vmax(rt.z, rt.w)));
                        }
                    }
                }
                smax2 = Pair<T>::vmax(smax2, swap_halves(smax2));                 // both halves: the lane's maximum
                uint32_t other;
                if constexpr (REF) other = Pair<T>::vmin(smin2, swap_halves(smin2)) ^ 0x80008000u;      // -min (sign flip is exact)
                else other = Pair<T>::vmax(tmax2, swap_halves(tmax2));
                uint32_t both = __byte_perm(smax2, other, 0x7610);                // low half: student max, high half: the other
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) both = Pair<T>::vmax(both, __shfl_xor_sync(0xffffffffu, both, o));
                float lo, hi;
                upk(Pair<T>::widen(both), lo, hi);
                m_w = lo;
                if constexpr (REF) min_w = -hi;
                else mt_w = hi;
            }
            const float mk = __fmul_rn((m_w == -INFINITY) ? 0.f : m_w, k2);
            mtk_cur = __fmul_rn((mt_w == -INFINITY) ? 0.f : mt_w, k2);

            // `red` after the grouped reduction: lanes 0-7 S_w, 8-15 Q_w, 16-23 R_w - P_w
            {
                const u64 nmk = pk(-mk, -mk);
                const float tref = REF ? lse2_t_cur : mtk_cur;
                const u64 nlt = pk(-tref, -tref);
                u64 accS0 = pk(0.f, 0.f), accS1 = pk(0.f, 0.f), accQ = pk(0.f, 0.f), accR = pk(0.f, 0.f), accP = pk(0.f, 0.f);
                auto ab_step = [&](int j, auto nv_tag, bool act, auto tail_tag, auto poly_tag) {
                    constexpr int NV = decltype(nv_tag)::value;
                    constexpr bool POLY = decltype(poly_tag)::value;
                    uint32_t qb[NV][8], ob[NV][8];
                    uint4 rs[NV], rt[NV];
#pragma unroll
                    for (int u = 0; u < NV; ++u) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { qb[u][i] = 0u; ob[u][i] = 0u; }
                        if constexpr (REF) tmem_ld8(tm + kTmTable + 8 * (j + u), ob[u]);
                    }
#pragma unroll
                    for (int u = 0; u < NV; ++u) {
                        rs[u] = make_uint4(0, 0, 0, 0);
                        rt[u] = rs[u];
                        if (act) {
                            rs[u] = load_s(lane + 32 * (j + u));
                            rt[u] = load_t(lane + 32 * (j + u));
                        }
                    }
                    if constexpr (REF) tmem_wait_ld();
                    if (act) {
#pragma unroll
                        for (int u = 0; u < NV; ++u) {
                            const uint32_t ws[4] = {rs[u].x, rs[u].y, rs[u].z, rs[u].w}, wt[4] = {rt[u].x, rt[u].y, rt[u].z, rt[u].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const u64 ovp = pk(__uint_as_float(ob[u][2 * i]), __uint_as_float(ob[u][2 * i + 1]));
                                const u64 a2 = fma2(Pair<T>::widen(ws[i]), k2p, nmk);        // xs*k2 - m_w*k2  (<= 0)
                                const u64 e2 = (POLY && i >= 4 - KD_POLY_AB) ? ex2_poly2<4>(a2) : ex2_mufu2(a2);
                                if (i & 1) accS1 = add2(accS1, e2);
                                else accS0 = add2(accS0, e2);
                                const u64 b2 = fma2(Pair<T>::widen(wt[i]), k2p, nlt);        // log2 tp (REF) / relative to mt_w (PLAIN)
                                float e_lo, e_hi;
                                if constexpr (REF) upk(fma2(ovp, pushp, b2), e_lo, e_hi);
                                else upk(b2, e_lo, e_hi);
                                const u64 t2 = pk(ex2(e_lo), ex2(e_hi));
                                const u64 q2 = REF ? add2(t2, ovp) : t2;
                                accQ = add2(accQ, q2);
                                accR = fma2(t2, b2, accR);
                                accP = fma2(q2, a2, accP);
                                float q_lo, q_hi;
                                upk(q2, q_lo, q_hi);
                                qb[u][2 * i] = __float_as_uint(q_lo);
                                qb[u][2 * i + 1] = __float_as_uint(q_hi);
                            }
                        }
                    }
                    if constexpr (decltype(tail_tag)::value) __syncwarp();     // reconverge before the warp-collective store
#pragma unroll
                    for (int u = 0; u < NV; ++u) tmem_st8(tm + kTmQ + 40 * par + 8 * (j + u), qb[u]);
                };
                const bool poly_ab = REF && KD_POLY_AB > 0 && (fmaf(min_w, k2, -mk) >= kPolyFloor);
                if (poly_ab) {
                    int j = 0;
                    for (; j + 1 < n_full; j += 2) ab_step(j, Int2{}, true, std::false_type{}, std::true_type{});
                    if (j < n_full) ab_step(j, Int1{}, true, std::false_type{}, std::true_type{});
                    if (n_tail) ab_step(n_full, Int1{}, lane < n_tail, std::true_type{}, std::true_type{});
                } else {
                    int j = 0;
                    for (; j + 1 < n_full; j += 2) ab_step(j, Int2{}, true, std::false_type{}, std::false_type{});
                    if (j < n_full) ab_step(j, Int1{}, true, std::false_type{}, std::false_type{});
                    if (n_tail) ab_step(n_full, Int1{}, lane < n_tail, std::true_type{}, std::false_type{});
                }
                tmem_wait_st();
                float s0, s1, q0, q1, r0, r1, p0, p1;
                upk(add2(accS0, accS1), s0, s1);
                upk(accQ, q0, q1);
                upk(accR, r0, r1);
                upk(accP, p0, p1);
                R_cur = (r0 + r1) + poison;                     // lane-level share, used by pass B' of a masked row
                red = warp_sum4_grouped(s0 + s1, q0 + q1, R_cur - (p0 + p1), 0.f, lane);
            }
        }
        // ---------------- the previous row's reduction result: needed from here on (push gating, clamp flag, pass C) ----------------
        if (have_prev) {
            mbar_wait_hint(bar_res1 + ppar * 8, ((it - 1) >> 1) & 1, 2000u);
            lse2_s_p = bc[0];
            if constexpr (!REF) lse2_t_prev = bc[1];
            gQ_p = bc[2];
            const int fl = __float_as_int(bc[3]);
            kept_p = fl & 1;
            fold_p = fl & 2;
            if (DYN) r_nn = __float_as_int(bc[4]);
            lgQ_p = bc[5];
        }
        if (have_cur) {
            // push to every CTA of the cluster.  Every warp has observed the result of row it-1 by now, which implies that every
            // CTA's fold warp has finished reading the slots of row it-2 (same parity) before pushing its own row it-1 partial.
            {
                const int grp = lane >> 3;                      // lane group g pushes word g+1 (S, Q, R-P); group 3 pushes m_w (word 0)
                const uint32_t dst_cta = lane & 7;
                const uint32_t slot = slots + par * NSLOT * 32 + (rank * kWarps + warp) * 32;
                const uint32_t rbar = map_to_cta(bar_red1 + par * 8, dst_cta);
                const uint32_t word = (grp == 3) ? 0u : (uint32_t)(grp + 1);
                const uint32_t val = (grp == 3) ? __float_as_uint(m_w) : __float_as_uint(red);
                st_async_b32(map_to_cta(slot + word * 4, dst_cta), val, rbar);
                if (DYN) claimed = __shfl_sync(0xffffffffu, claimed, 0);
                if (lane < CS) {
                    if constexpr (REF) st_async_v2(map_to_cta(slot + 16, dst_cta), __float_as_uint(min_w), (uint32_t)claimed, rbar);
                    else st_async_v4(map_to_cta(slot + 16, dst_cta), 0u, (uint32_t)claimed, __float_as_uint(mt_w), 0u, rbar);
                }
            }
        }

        // ---------------- pass B' (previous row, rare): exact masked sums and a second, exposed reduction ----------------
        if constexpr (REF) {
            if (have_prev && !kept_p) {
                const u64 nls = pk(-lse2_s_p, -lse2_s_p);
                u64 accN = pk(0.f, 0.f), accM = pk(0.f, 0.f);
                auto bp_step = [&](int j, auto nv_tag, bool act, auto tail_tag) {
                    constexpr int NV = decltype(nv_tag)::value;
                    uint32_t qb[NV][8];
                    uint4 rs[NV];
#pragma unroll
                    for (int u = 0; u < NV; ++u) tmem_ld8(tm + kTmQ + 40 * ppar + 8 * (j + u), qb[u]);
#pragma unroll
                    for (int u = 0; u < NV; ++u) {
                        rs[u] = make_uint4(0, 0, 0, 0);
                        if (act) rs[u] = load_sp(lane + 32 * (j + u));
                    }
                    tmem_wait_ld();
                    if (act) {
#pragma unroll
                        for (int u = 0; u < NV; ++u) {
                            const uint32_t ws[4] = {rs[u].x, rs[u].y, rs[u].z, rs[u].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float a_lo, a_hi;
                                upk(fma2(Pair<T>::widen(ws[i]), k2p, nls), a_lo, a_hi);
                                const u64 q2 = pk(__uint_as_float(qb[u][2 * i]), __uint_as_float(qb[u][2 * i + 1]));
                                accN = fma2(q2, pk(fmaxf(a_lo, thr2), fmaxf(a_hi, thr2)), accN);
                                accM = fma2(q2, pk(set_ge(a_lo, thr2), set_ge(a_hi, thr2)), accM);
                            }
                        }
                    }
                    if constexpr (decltype(tail_tag)::value) __syncwarp();
                };
                {
                    int j = 0;
                    for (; j + 1 < n_full; j += 2) bp_step(j, Int2{}, true, std::false_type{});
                    if (j < n_full) bp_step(j, Int1{}, true, std::false_type{});
                    if (n_tail) bp_step(n_full, Int1{}, lane < n_tail, std::true_type{});
                }
                float x0, x1, y0, y1;
                upk(accN, x0, x1);
                upk(accM, y0, y1);
                const float N_w = warp_sum(R_prev - (x0 + x1));        // R_prev is this lane's share of sum t*b
                const float M_w = warp_sum(y0 + y1);
                if (lane < CS)
                    st_async_v2(map_to_cta(slots + ppar * NSLOT * 32 + (rank * kWarps + warp) * 32 + 24, lane), __float_as_uint(N_w),
                                __float_as_uint(M_w), map_to_cta(bar_red2 + ppar * 8, lane));
            }
        }

        // ---------------- previous row: second reduction result (masked rows only), then pass C ----------------
        if (have_prev) {
            if constexpr (REF) {
                if (!kept_p) {
                    const uint32_t ph2 = ppar ? ph2_1 : ph2_0;
                    if (warp == FOLD_WARP) {
                        mbar_wait_hint(bar_red2 + ppar * 8, ph2, 1000u);
                        if (lane == 0) mbar_expect_tx(bar_red2 + ppar * 8, RED2_BYTES);
                        const float4* slot_v = reinterpret_cast<const float4*>(smem + plan.slots + ppar * NSLOT * 32);
                        float rr = 0.f, qq = 0.f;
#pragma unroll
                        for (int i = 0; i < SPL; ++i) {
                            const float4 sb = slot_v[2 * (lane + 32 * i) + 1];
                            rr += sb.z;
                            qq += sb.w;
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            rr += __shfl_xor_sync(0xffffffffu, rr, o);
                            qq += __shfl_xor_sync(0xffffffffu, qq, o);
                        }
                        if (lane == 0) {
                            const float gQ = g * qq;
                            const bool fold = (g > 0.f) && (gQ > 0.f) && (gQ < INFINITY);
                            bc[2] = gQ;
                            bc[3] = __int_as_float(fold ? 2 : 0);
                            bc[5] = fold ? __log2f(gQ) : 0.f;
                            if (rank == 0) p.rowaux[r_prev] = make_float4(lse2_t_prev, lse2_s_p, qq, rr);
                            mbar_arrive_cta(bar_res2 + ppar * 8);
                        }
                    }
                    mbar_wait_hint(bar_res2 + ppar * 8, ph2, 2000u);
                    gQ_p = bc[2];
                    fold_p = __float_as_int(bc[3]) & 2;
                    lgQ_p = bc[5];
                    if (ppar) ph2_1 ^= 1;
                    else ph2_0 ^= 1;
                }
            }
            if (want_grad) {
                // fold: g*Q rides in the exponent, 2^(a + log2 gQ) - g*q;  otherwise gQ * 2^a - g*q
                const float nl = lgQ_p - lse2_s_p;
                const u64 nls = pk(nl, nl);
                const float thrC = thr2 + lgQ_p;
                const float gq = REF ? g : g * ex2(mtk_prev - lse2_t_prev);        // PLAIN: the parked values are relative to mt_w
                const u64 negg = pk(-gq, -gq), gQp = pk(gQ_p, gQ_p);
                T* grow = static_cast<T*>(p.grad) + r_prev * p.ld_g + (int64_t)g_v0 * 8;
                auto sweepC = [&](auto fold_tag, auto mask_tag, auto poly_tag) {
                    constexpr bool FOLD = decltype(fold_tag)::value;
                    constexpr bool MASK = decltype(mask_tag)::value;
                    constexpr bool POLY = decltype(poly_tag)::value;
                    auto c_step = [&](int j, auto nv_tag, bool act, auto tail_tag) {
                        constexpr int NV = decltype(nv_tag)::value;
                        uint32_t qb[NV][8];
                        uint4 rs[NV];
#pragma unroll
                        for (int u = 0; u < NV; ++u) tmem_ld8(tm + kTmQ + 40 * ppar + 8 * (j + u), qb[u]);
#pragma unroll
                        for (int u = 0; u < NV; ++u) {
                            rs[u] = make_uint4(0, 0, 0, 0);
                            if (act) rs[u] = load_sp(lane + 32 * (j + u));
                        }
                        tmem_wait_ld();
                        if (act) {
#pragma unroll
                            for (int u = 0; u < NV; ++u) {
                                const uint32_t ws[4] = {rs[u].x, rs[u].y, rs[u].z, rs[u].w};
                                uint32_t out[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const u64 a2 = fma2(Pair<T>::widen(ws[i]), k2p, nls);
                                    float a_lo, a_hi;
                                    upk(a2, a_lo, a_hi);
                                    const u64 sp2 = (POLY && i >= 4 - KD_POLY_C) ? ex2_poly2<3>(a2) : ex2_mufu2(a2);
                                    u64 q2 = pk(__uint_as_float(qb[u][2 * i]), __uint_as_float(qb[u][2 * i + 1]));
                                    if constexpr (MASK) q2 = mul2(q2, pk(set_ge(a_lo, thrC), set_ge(a_hi, thrC)));
                                    u64 o2;
                                    if constexpr (FOLD) o2 = fma2(q2, negg, sp2);
                                    else o2 = fma2(sp2, gQp, mul2(q2, negg));
                                    float n_lo, n_hi;
                                    upk(o2, n_lo, n_hi);
                                    out[i] = Pair<T>::narrow(n_lo, n_hi);
                                }
                                stg_stream(grow + (int64_t)(lane + 32 * (j + u)) * 8, make_uint4(out[0], out[1], out[2], out[3]));
                            }
                        }
                        if constexpr (decltype(tail_tag)::value) __syncwarp();
                    };
                    int j = 0;
                    for (; j + 1 < n_full; j += 2) c_step(j, Int2{}, true, std::false_type{});
                    if (j < n_full) c_step(j, Int1{}, true, std::false_type{});
                    if (n_tail) c_step(n_full, Int1{}, lane < n_tail, std::true_type{});
                };
                using Yes = std::true_type;
                using No = std::false_type;
                const bool poly_c = REF && KD_POLY_C > 0 && (fmaf(min_prev, k2, nl) >= kPolyFloor);
                if (fold_p) {
                    if (kept_p) {
                        if (poly_c) sweepC(Yes{}, No{}, Yes{});
                        else sweepC(Yes{}, No{}, No{});
                    } else {
                        sweepC(Yes{}, Yes{}, No{});
                    }
                } else {
                    if (kept_p) sweepC(No{}, No{}, No{});
                    else sweepC(No{}, Yes{}, No{});
                }
            }
        }
        __syncwarp();
        // advance: the current row becomes the previous one, ring slots rotate
        r_prev = have_cur ? r_cur : -1;
        if constexpr (REF) lse2_t_prev = lse2_t_cur;
        mtk_prev = mtk_cur;
        R_prev = R_cur;
        min_prev = min_w;
        r_cur = r_nxt;
        r_nxt = r_nn;
        {
            const uint32_t t = st_prev;
            st_prev = st_cur;
            st_cur = st_nxt;
            st_nxt = t;
            const uint32_t u = te_cur;
            te_cur = te_nxt;
            te_nxt = u;
        }
    }
    tmem_fence_before_sync();
    cluster_sync_all();     // nobody leaves while peers may still push into its shared memory
    tmem_fence_after_sync();
    if (warp == 1) tmem_dealloc_512(tmem_base);
}

// --------------------------------------------------------------------------------------------------------
template <int VARIANT, typename T, int CS, bool DYN>
int launch_fused_d(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    constexpr int NPIPE = 2;
    auto kfn = row_emit_fused_kernel<VARIANT, T, CS, NPIPE, DYN>;
    const int nvec = e.V / 8;
    const int vpc = (nvec + CS - 1) / CS;
    const FusedPlan<VARIANT, CS, NPIPE> plan(vpc);
    constexpr int WPP = kWarps / NPIPE;
    const int w_chunk = ((vpc + WPP * 32 - 1) / (WPP * 32)) * 32;
    // the thread's q cells (8 words per iteration, from word 0) and fp16 exponent pairs (4 per iteration, from word kTmE)
    // must fit its 128 tensor-memory words without overlapping
    if (w_chunk / 32 > (KD_ECACHE ? kTmE / 8 : kMaxIters)) return KD_ERR_UNSUPPORTED;
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
        if (CS > 8) {       // clusters of 9 are "non-portable": B200 co-schedules 15 of them (135 SMs) against 15 x 8 = 120
            err = cudaFuncSetAttribute(kfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (err != cudaSuccess) { cudaGetLastError(); return KD_ERR_UNSUPPORTED; }
        }
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
    cp.trace = nullptr;
#ifdef KD_TRACE
    {
        static unsigned long long* tbuf = nullptr;
        const size_t n = (size_t)max_clusters * CS * kWarps * 64 * 8;
        if (!tbuf) { cudaMalloc(&tbuf, n * sizeof(unsigned long long)); cudaMemset(tbuf, 0, n * sizeof(unsigned long long)); }
        cp.trace = tbuf;
        if (const char* pth = getenv("KD_TRACE_DUMP")) {       // previous launch's stamps -> file (debug only)
            static int calls = 0;
            if (++calls == 3) {
                cudaDeviceSynchronize();
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
    {
        static int early = -1;
        if (early < 0) { const char* ev = getenv("KD_TOKEN_EARLY"); early = ev ? atoi(ev) : 0; }
        cp.tok_after = early;
    }
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

template <int VARIANT, typename T, int CS>
int launch_fused_c(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    if (e.row_queue) return launch_fused_d<VARIANT, T, CS, true>(e, st, query_only, capacity, launched, first_free);
    return launch_fused_d<VARIANT, T, CS, false>(e, st, query_only, capacity, launched, first_free);
}

// cluster size: whichever of 8 / 9 puts more SMs to work for this shape (KD_CLUSTER_SIZE=8|9 forces one)
int g_last_cluster_size = 8;
template <int VARIANT, typename T>
int launch_fused_v(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    static int forced = -1;
    if (forced < 0) { const char* ev = getenv("KD_CLUSTER_SIZE"); forced = ev ? atoi(ev) : 0; }
    int cap8 = 0, cap9 = 0;
    const int rc8 = (forced == 9) ? KD_ERR_UNSUPPORTED : launch_fused_c<VARIANT, T, 8>(e, st, true, &cap8, nullptr, nullptr);
    const int rc9 = (forced == 8) ? KD_ERR_UNSUPPORTED : launch_fused_c<VARIANT, T, 9>(e, st, true, &cap9, nullptr, nullptr);
    const bool use9 = (rc9 == KD_OK) && (rc8 != KD_OK || cap9 * 9 > cap8 * 8);
    if (!use9 && rc8 != KD_OK) return rc8;
    g_last_cluster_size = use9 ? 9 : 8;
    if (use9) return launch_fused_c<VARIANT, T, 9>(e, st, query_only, capacity, launched, first_free);
    return launch_fused_c<VARIANT, T, 8>(e, st, query_only, capacity, launched, first_free);
}

template <typename T>
int launch_fused_t(const EmitParams& e, int variant, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    if (variant == KD_LOCA_REFERENCE) return launch_fused_v<KD_LOCA_REFERENCE, T>(e, st, query_only, capacity, launched, first_free);
    if (variant == KD_LOCA_PER_ROW) return launch_fused_v<KD_LOCA_PER_ROW, T>(e, st, query_only, capacity, launched, first_free);
    return launch_fused_v<KD_KL_PLAIN, T>(e, st, query_only, capacity, launched, first_free);
}

template <int VARIANT, typename T, bool DYN>
int launch_pipe_d(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    constexpr int CS = 8;
    auto kfn = row_emit_pipelined_kernel<VARIANT, T, CS, DYN>;
    const int nvec = e.V / 8;
    const int vpc = (nvec + CS - 1) / CS;
    const PipePlan<VARIANT, CS> plan(vpc);
    if ((vpc + 31) / 32 > kWarps * kMaxIt) return KD_ERR_UNSUPPORTED;      // a thread's cells must fit its tensor-memory words
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
    cfg.dynamicSmemBytes = max_smem;     // one CTA per SM: it owns all 512 tensor-memory columns
    cfg.stream = st;
    if (configured_dev != dev) {
        err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(pipelined kernel smem)");
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
    FusedParams cp;
    cp.trace = nullptr;
    cp.tok_after = 0;
    cp.e = e;
    cp.n_clusters = (int)(e.n_rows < max_clusters ? e.n_rows : max_clusters);
    cp.vpc = vpc;
    cp.nvec = nvec;
    cfg.gridDim = dim3(cp.n_clusters * CS);
    if (launched) *launched = cp.n_clusters;
    if (first_free) *first_free = 3 * cp.n_clusters;
    err = cudaLaunchKernelEx(&cfg, kfn, cp);
    if (err != cudaSuccess) return cuda_fail(err, "row_emit_pipelined_kernel launch");
    return KD_OK;
}

template <typename T>
int launch_pipe_t(const EmitParams& e, int variant, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    if (variant == KD_LOCA_REFERENCE) {
        if (e.row_queue) return launch_pipe_d<KD_LOCA_REFERENCE, T, true>(e, st, query_only, capacity, launched, first_free);
        return launch_pipe_d<KD_LOCA_REFERENCE, T, false>(e, st, query_only, capacity, launched, first_free);
    }
    if (e.row_queue) return launch_pipe_d<KD_KL_PLAIN, T, true>(e, st, query_only, capacity, launched, first_free);
    return launch_pipe_d<KD_KL_PLAIN, T, false>(e, st, query_only, capacity, launched, first_free);
}

// KD_FUSED: 0 = three-sweep cluster kernel everywhere, 1 = two-pipeline fused kernel (default), 2 = software-pipelined kernel
int fused_mode() {
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
    } else if (variant == KD_LOCA_PER_ROW) {
        if (!p.lse2_t_in || !p.records) return false;               // needs phase A (teacher statistics + token records)
    } else if (variant != KD_KL_PLAIN) {
        return false;
    }
    return dtype == KD_DTYPE_BF16 || dtype == KD_DTYPE_F16;
}

int fused_cluster_size() { return g_last_cluster_size; }

int launch_emit_fused(const EmitParams& p, int variant, int dtype, cudaStream_t stream, bool query_only, int* capacity,
                      int* clusters_launched, int* first_free_row, int* rows_in_flight) {
    g_last_cluster_size = 8;
    if (fused_mode() >= 2 && variant != KD_LOCA_PER_ROW) {
        if (rows_in_flight) *rows_in_flight = 1;
        const int rc = (dtype == KD_DTYPE_BF16)
                           ? launch_pipe_t<__nv_bfloat16>(p, variant, stream, query_only, capacity, clusters_launched, first_free_row)
                           : launch_pipe_t<__half>(p, variant, stream, query_only, capacity, clusters_launched, first_free_row);
        if (rc != KD_ERR_UNSUPPORTED) return rc;
    }
    if (rows_in_flight) *rows_in_flight = 2;
    if (dtype == KD_DTYPE_BF16) return launch_fused_t<__nv_bfloat16>(p, variant, stream, query_only, capacity, clusters_launched, first_free_row);
    return launch_fused_t<__half>(p, variant, stream, query_only, capacity, clusters_launched, first_free_row);
}

}  // namespace kd
