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
// Variants: KD_LOCA_REFERENCE (teacher lse2_t and the override table come from phase A), KD_KL_PLAIN (teacher
// statistics are folded into the same single reduction through a per-warp teacher reference maximum) and
// KD_KL_LOG_TARGET (FB:206-219: w = e^{tp} in place of q) — e^{tp} is NOT linear in the unknown teacher log-sum-exp, but its
// power series is: with tp = sigma_w * t' (t' = teacher exponential against the warp's reference maximum, sigma_w known after the
// fold) sum_v e^{tp} = n + sum_k sigma^k T_k / k!, T_k = sum t'^k, and likewise for sum e^{tp} tp and sum e^{tp} a with the cross
// sums U_k = sum t'^k a'.  tp <= 1 and the totals are dominated by the order-0 terms (~V, ~V log V), so T_1..T_3 and U_0..U_3
// bound the truncation at 3e-7 of Q_r and 1e-6 of the row sum; the gradient's e^{tp} is exact (one ex2 per cell in pass C) —
// and KD_LOCA_PER_ROW, which is the plain sweep plus, in the SAME reduction, each warp's best / runner-up teacher cell (value,
// column, student logit there), the row's label logits and the student minimum: the fold derives v1, v2 (P1:166-185),
// corrects Q_r and the row sum for the row's own two override cells analytically and the two gradient cells are
// re-written after the gradient sweep — the teacher row is read ONCE (6V bytes per token, no phase A).
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
// Prefetching the next row's slices into L2 at the start of a row (cp.async.bulk.prefetch.L2) is measured HARMFUL now that the
// refills are issued early — the TMA refill right after the sweep already hides the latency.  Three A/B pairs on one box: as-written
// 3.645 -> 3.564 ms, per-row 3.74 -> 3.56, clamped rows 4.63 -> 4.50, log_target 3.63 -> 3.53, plain KL unchanged.  Kept as a knob.
#ifndef KD_WARM_L2
#define KD_WARM_L2 0
#endif
#ifndef KD_STREAM_HINTS
#define KD_STREAM_HINTS 0
#endif
#ifndef KD_P0_UNROLL
#define KD_P0_UNROLL 4         // pass 0 (packed maxima)
#endif
#ifndef KD_BP_UNROLL
#define KD_BP_UNROLL 3         // pass B' (clamped rows): 3 against 2 = 4.77 -> 4.62 ms at T = 0.8 on peaked logits (1: 4.71)
#endif
#ifndef KD_C_NV
#define KD_C_NV 4             // vectors per tensor-memory wait in the cached gradient sweep
#endif
#ifndef KD_AB_UNROLL
#define KD_AB_UNROLL 3        // iterations of the statistics + accumulation sweep unrolled together: a 9-CTA cluster gives a warp 9 of them.
#endif                        // Measured, two boxes (tools/build_variant.sh): 3 against 2: as-written 3.72 -> 3.66 / 3.71 -> 3.67 ms, plain KL 2.86 -> 2.78 /
                              // 3.03 -> 2.93, clamped rows 4.80 -> 4.75; 4 and 9 no better; per-row LoCa within noise either way (kept at 2)
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
    int tok_after;                // two-pipeline kernel: hand the pass-AB token over this many iterations early (0 in the product)
};

// Shared-memory plan (dynamic, byte offsets).  A reduction slot is 8 words:
//   [0] m_w  [1] S_w  [2] Q_w  [3] R_w - P_w  [4] min_w  [5] next-row claim (leader slot)  [6] [7] second reduction
//   (KD_KL_PLAIN: [6] teacher reference maximum, no second reduction)
// KD_KL_LOG_TARGET uses 12 of 16 words: [0] m_w [1] S_w [2] T_1 [3] U_1 [4] T_2 [5] claim [6] mt_w [7] T_3 [8] U_0 [9] U_2 [10] U_3
//   [11] number of cells in the warp's slice
// KD_LOCA_PER_ROW uses 16 words: [0-5] as above, [6] teacher maximum of the warp's slice = its best cell's value, [7] that
//   cell's column, [8] [9] runner-up value / column, [10] [11] student logits at those two cells, [12] [13] teacher / student
//   logit at the row's label (meaningful in the slot of the warp that owns the column), [14] [15] second reduction
// With the fused cross-entropy (CE, as-written LoCa only) a slot also carries [8] S1_w = sum 2^((xs - m_w) log2 e) and [9] the
//   student logit at the row's CE label (meaningful in the slot of the warp that owns the column): 16 words.
template <int VARIANT, int CS, int NPIPE, bool CE = false>
struct FusedPlan {
    static constexpr int WPP = kWarps / NPIPE;
    static constexpr int NSLOT = CS * WPP;
    static constexpr int SLOTB = (VARIANT == KD_LOCA_PER_ROW || VARIANT == KD_KL_LOG_TARGET || CE) ? 64 : 32;      // bytes per reduction slot
    static constexpr int BCASTB = 128;                                        // bytes of per-pipeline broadcast words
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
        off += NPIPE * NSLOT * SLOTB;
        bcast = off;
        off += NPIPE * BCASTB;
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
// xlogy(v, v) of a row's own override cells in log2 units (P1:184-185 applied per row: the runner-up cell wins when it is the
// label's); NaN for a label outside [0, V) — the reference's gather raises there (P1:166)
__device__ __forceinline__ float own_xlogy2(float v1, float v2, bool same, int label) {
    float x = 0.f;
    if (!same && v1 != 0.f) x += v1 * __log2f(v1);
    if (v2 != 0.f) x += v2 * __log2f(v2);
    return label < 0 ? __int_as_float(0x7fc00000) : x;
}

template <int VARIANT, typename T, int CS, int NPIPE, bool DYN, bool CE = false>
__global__ void __launch_bounds__(kThreads, 1)
row_emit_fused_kernel(const FusedParams cp) {
    static_assert(!CE || VARIANT == KD_LOCA_REFERENCE, "the fused cross-entropy rides the as-written LoCa kernel");
    static_assert(VARIANT != KD_KL_LOG_TARGET || KD_ECACHE, "the log_target variant relies on the cached gradient sweep");
    constexpr bool TAB = (VARIANT == KD_LOCA_REFERENCE);            // shared override table (phase A + build_table)
    constexpr bool PER_ROW = (VARIANT == KD_LOCA_PER_ROW);          // the row's own label / runner-up cells only
    constexpr bool KLLT = (VARIANT == KD_KL_LOG_TARGET);            // w = e^{tp}: power sums of the teacher exponentials
    constexpr bool CLAMP = TAB || PER_ROW;                          // sp >= eps clamp: student minimum, masked pass B'
    // Experiment knob (off): evict-first row loads + cache-streaming gradient stores for the single-launch variants.  Two A/B pairs on
    // one box: no difference (plain KL 2.81-2.84 against 2.80, per-row 3.47 against 3.45 ms); a first reading of -3 % was the
    // compiler laying the kernel out differently, not the hints — differences of a few percent between builds of this kernel are noise
    // of that kind unless they repeat across unrelated edits.
    constexpr bool STREAM_HINTS = (KD_STREAM_HINTS != 0) && !TAB;
    constexpr int kP0Unroll = KD_P0_UNROLL, kBpUnroll = KD_BP_UNROLL;
    constexpr int kAbUnroll = PER_ROW ? 2 : KD_AB_UNROLL;           // (a macro is not expanded inside #pragma unroll; a constant expression is)
    constexpr int WPP = kWarps / NPIPE;
    constexpr int NSLOT = CS * WPP;
    constexpr int SLOTB = FusedPlan<VARIANT, CS, NPIPE, CE>::SLOTB;
    constexpr int SLOTV = SLOTB / 16;                               // float4s per slot
    constexpr int BCASTB = FusedPlan<VARIANT, CS, NPIPE, CE>::BCASTB;
    constexpr uint32_t RED1_BYTES = NSLOT * (TAB ? (CE ? 32 : 24) : PER_ROW ? 56 : KLLT ? 48 : 32);
    constexpr uint32_t RED2_BYTES = NSLOT * 8;
    constexpr uint32_t RED2_OFF = (PER_ROW || CE) ? 56 : 24;        // byte offset of the second reduction's two words (the slot's last two)
    extern __shared__ __align__(128) unsigned char smem[];
    const EmitParams& p = cp.e;
    const FusedPlan<VARIANT, CS, NPIPE, CE> plan(cp.vpc);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pipe = warp / WPP, pw = warp % WPP;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x / CS;
    const float NaN = __int_as_float(0x7fc00000);

    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar_full = s_base + plan.bars + warp * 8;
    const uint32_t bar_red1 = s_base + plan.bars + (kWarps + 2 * pipe) * 8;
    const uint32_t bar_red2 = bar_red1 + 8;
    const uint32_t slots = s_base + plan.slots + pipe * NSLOT * SLOTB;
    const float4* slot_v = reinterpret_cast<const float4*>(smem + plan.slots + pipe * NSLOT * SLOTB);
    volatile float* bcast = reinterpret_cast<volatile float*>(smem + plan.bcast + pipe * BCASTB);
    const int bar1 = 1 + 2 * pipe, bar2 = 2 + 2 * pipe;     // named barriers of this pipeline (0 is __syncthreads)
    const int bar3 = 1 + 2 * NPIPE + pipe;                  // per-row LoCa: hand-over between the pipeline's two fold warps
    // Ping-pong of the two pipelines: left alone they fall into lockstep (measured: both in pass AB, then both idle on their
    // reductions, then both bursting stores).  A pipeline may start pass AB only when the other one has finished its own, so
    // one pipeline's XU-bound sweep overlaps the other's reduction wait, pass C, refill and pass 0.
    const uint32_t tok_mine = s_base + plan.bars + (kWarps + 2 * NPIPE + pipe) * 8;
    const uint32_t tok_other = s_base + plan.bars + (kWarps + 2 * NPIPE + (NPIPE - 1 - pipe)) * 8;
    volatile float* done_other = reinterpret_cast<volatile float*>(smem + plan.bcast + (NPIPE - 1 - pipe) * BCASTB) + 7;
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
            reinterpret_cast<volatile float*>(smem + plan.bcast + i * BCASTB)[7] = 0.f;     // "this pipeline has left its row loop"
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
            if constexpr (STREAM_HINTS) tma_load_1d_evict_first(s_t, tbase + row * p.ld_t + (int64_t)g_v0 * 8, bytes, bar_full);
            else tma_load_1d(s_t, tbase + row * p.ld_t + (int64_t)g_v0 * 8, bytes, bar_full);
        }
    };
    auto issue_student = [&](int64_t row) {
        if (lane == 0 && w_n > 0) {
            const uint32_t bytes = (uint32_t)w_n * 16u;
            fence_proxy_async();
            if constexpr (STREAM_HINTS) tma_load_1d_evict_first(s_s, sbase + row * p.ld_s + (int64_t)g_v0 * 8, bytes, bar_full);
            else tma_load_1d(s_s, sbase + row * p.ld_s + (int64_t)g_v0 * 8, bytes, bar_full);
        }
    };
    auto issue_load = [&](int64_t row) {
        issue_teacher(row);
        issue_student(row);
    };
    [[maybe_unused]] auto warm_l2 = [&](int64_t row) {           // bring the warp's slice of `row` into L2 ahead of its use (KD_WARM_L2)
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
    int64_t label64 = -1;
    if constexpr (PER_ROW) {
        if (r < p.n_rows) label64 = p.labels[r];
    }
    // fused cross-entropy (temperature 1): exponent scale, upstream * ce_weight / n_valid, whether it shares the KD temperature
    [[maybe_unused]] const float k1 = kLog2e;
    [[maybe_unused]] const bool ce_same_t = (k2 == k1);
    // T == 2: 2^((xs - m) log2 e) is the SQUARE of the KD exponential 2^((xs - m) log2 e / 2) — no second ex2 in either sweep
    [[maybe_unused]] const bool ce_square = (k2 + k2 == k1);
    [[maybe_unused]] float g_ce = 0.f;
    [[maybe_unused]] int64_t cel64 = kCeIgnore;
    if constexpr (CE) {
        g_ce = up * p.ce_weight * *p.ce_inv_count;
        if (r < p.n_rows) cel64 = p.ce_labels[r];
    }
    while (r < p.n_rows) {
        ++row_iter;
        KD_STAMP(0);
        if (!DYN) next_r = r + row_stride;
        int claimed = 0;
        if (DYN && rank == 0 && pw == 0 && lane == 0) claimed = atomicAdd(p.row_queue, 1);
        float lse2_t = 0.f;
        if constexpr (TAB) lse2_t = p.lse2_t_in[r];
        int label = -1;                                   // the row's label column (per-row LoCa); -1 = out of range
        if constexpr (PER_ROW) {
            label = (label64 >= 0 && label64 < (int64_t)p.V) ? (int)label64 : -1;
            if (next_r < p.n_rows) label64 = p.labels[next_r];         // the next row's label: its latency hides behind this row
        }
        // fused CE: the row's next-token id; -100 = ignored (prompt / padding: no CE work at all for the row), anything else outside
        // [0, V) is an error like in F.cross_entropy (status bit + NaN)
        [[maybe_unused]] bool ce_on = false, ce_bad = false;
        [[maybe_unused]] int ce_c = -1;
        if constexpr (CE) {
            ce_on = (cel64 != kCeIgnore);
            ce_bad = ce_on && !(cel64 >= 0 && cel64 < (int64_t)p.V);
            ce_c = (ce_on && !ce_bad) ? (int)cel64 : -1;
            if (next_r < p.n_rows) cel64 = p.ce_labels[next_r];
        }
#if KD_WARM_L2
        warm_l2(next_r);
#endif
        const uint32_t epoch = (uint32_t)r;
        auto load_s = [&](int v) -> uint4 { return lds128_ep(s_s + v * 16, epoch); };
        auto load_t = [&](int v) -> uint4 { return lds128_ep(s_t + v * 16, epoch); };
        if (w_n > 0) mbar_wait_cta(bar_full, ph_full);
        ph_full ^= 1;
        KD_STAMP(1);

        // ================= pass 0: reference maxima (and the student minimum) of the warp's slice ==============
        float m_w, min_w = INFINITY, mt_w = -INFINITY;
        uint32_t thrp = 0u;                 // per-row LoCa: a lower bound of the slice's runner-up teacher logit, in both halves
        float xt_lab = 0.f, xs_lab = 0.f;   // per-row LoCa: the row's label logits (held by the warp that owns the column)
        [[maybe_unused]] float xs_ce = 0.f; // fused CE: the student logit at the row's next-token id (same ownership rule)
        {
            uint32_t smax2 = Pair<T>::kNegInf2, smin2 = Pair<T>::kPosInf2, tmax2 = Pair<T>::kNegInf2;
#pragma unroll kP0Unroll
            for (int v = lane; v < w_n; v += 32) {
                const uint4 rs = load_s(v);
                smax2 = Pair<T>::vmax(smax2, Pair<T>::vmax(Pair<T>::vmax(rs.x, rs.y), Pair<T>::vmax(rs.z, rs.w)));
                if constexpr (CLAMP)
                    smin2 = Pair<T>::vmin(smin2, Pair<T>::vmin(Pair<T>::vmin(rs.x, rs.y), Pair<T>::vmin(rs.z, rs.w)));
                if constexpr (!TAB) {
                    const uint4 rt = load_t(v);
                    tmax2 = Pair<T>::vmax(tmax2, Pair<T>::vmax(Pair<T>::vmax(rt.x, rt.y), Pair<T>::vmax(rt.z, rt.w)));
                }
            }
            // (student max, -student min) [LoCa] or (student max, teacher max) [plain KL] ride one packed butterfly
            smax2 = Pair<T>::vmax(smax2, __byte_perm(smax2, smax2, 0x1032));                 // both halves: the lane's maximum
            uint32_t other;
            if constexpr (CLAMP) other = Pair<T>::vmin(smin2, __byte_perm(smin2, smin2, 0x1032)) ^ 0x80008000u;      // -min (exact)
            else other = Pair<T>::vmax(tmax2, __byte_perm(tmax2, tmax2, 0x1032));
            uint32_t both = __byte_perm(smax2, other, 0x7610);                                // low half: student max, high half: the other
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) both = Pair<T>::vmax(both, __shfl_xor_sync(0xffffffffu, both, o));
            float lo, hi;
            upk(Pair<T>::widen(both), lo, hi);
            m_w = lo;
            if constexpr (CLAMP) min_w = -hi;
            else mt_w = hi;
            if constexpr (PER_ROW) {
                // teacher maximum of the slice, and the second largest LANE maximum: two different lanes hold cells >= it, so
                // it bounds the slice's runner-up from below and the sweep only has to look at vectors that reach it
                float tlo, thi;
                upk(Pair<T>::widen(Pair<T>::vmax(tmax2, __byte_perm(tmax2, tmax2, 0x1032))), tlo, thi);
                const int key = fkey(tlo);
                const int kmax = __reduce_max_sync(0xffffffffu, key);
                mt_w = fkey_inv(kmax);
                const unsigned holders = __ballot_sync(0xffffffffu, key == kmax);
                int k2nd = kmax;
                if (__popc(holders) < 2) k2nd = __reduce_max_sync(0xffffffffu, (lane == __ffs(holders) - 1) ? (int)0x80000000 : key);
                const float thr = fkey_inv(k2nd);
                thrp = Pair<T>::narrow(thr, thr);
                const int vl = (label >> 3) - g_v0;                   // warp-uniform
                if (label >= 0 && vl >= 0 && vl < w_n) {
                    xt_lab = Elem<T>::to_float(*reinterpret_cast<const T*>(smem + (s_t - s_base) + vl * 16 + (label & 7) * 2));
                    xs_lab = Elem<T>::to_float(*reinterpret_cast<const T*>(smem + (s_s - s_base) + vl * 16 + (label & 7) * 2));
                }
            }
            if constexpr (CE) {
                const int vl = (ce_c >> 3) - g_v0;                    // warp-uniform
                if (ce_c >= 0 && vl >= 0 && vl < w_n)
                    xs_ce = Elem<T>::to_float(*reinterpret_cast<const T*>(smem + (s_s - s_base) + vl * 16 + (ce_c & 7) * 2));
            }
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
        [[maybe_unused]] float kl_T2 = 0.f, kl_T3 = 0.f, kl_U0 = 0.f, kl_U2 = 0.f, kl_U3 = 0.f, kl_n = 0.f;      // log_target power sums
        [[maybe_unused]] float S1_w = 0.f;       // fused CE: sum 2^((xs - m_w) log2 e) of the slice
        // the CE needs its own exponentials only when the KD temperature is not 1 (and the row's label is not ignored)
        [[maybe_unused]] const bool ce_extra = CE && ce_on && !ce_same_t && !ce_square;      // own exponentials from the raw student slice
        [[maybe_unused]] const bool ce_sq = CE && ce_on && ce_square;
        uint32_t top_k1 = 0u, top_k2 = 0u;       // per-row LoCa: keys of the slice's best / runner-up teacher cell (0 = none)
        {
            constexpr bool ECACHE = KD_ECACHE;        // every variant of this kernel parks its student exponentials
            // with the exponent cache the student argument carries +15 (e is parked as fp16(e * 2^15)); S_w and P_w are corrected below
            const float esh = ECACHE ? kEScaleLog2 : 0.f;
            const u64 nmk = pk(esh - mk, esh - mk);
            const float tref = TAB ? lse2_t : mtk;
            const u64 nlt = pk(-tref, -tref);
            u64 accS0 = pk(0.f, 0.f), accS1 = pk(0.f, 0.f), accQ = pk(0.f, 0.f), accR = pk(0.f, 0.f), accP = pk(0.f, 0.f);
            [[maybe_unused]] u64 accCE = pk(0.f, 0.f);
            [[maybe_unused]] u64 accT2 = pk(0.f, 0.f), accT3 = pk(0.f, 0.f), accU0 = pk(0.f, 0.f), accU2 = pk(0.f, 0.f), accU3 = pk(0.f, 0.f);
            [[maybe_unused]] const u64 k1p = pk(k1, k1);
            [[maybe_unused]] const float mk1 = __fmul_rn((m_w == -INFINITY) ? 0.f : m_w, k1);
            [[maybe_unused]] const u64 nm1 = pk(-mk1, -mk1);
            uint32_t cand = 0u;            // per-row LoCa: bit j = this lane's vector of iteration j may hold the best / runner-up cell
            auto ab_body = [&](int j, bool act, auto tail_tag, auto poly_tag, auto ce_tag) {
                constexpr bool POLY = decltype(poly_tag)::value;
                constexpr int CE_X = decltype(ce_tag)::value;        // 0: nothing extra, 1: own ex2 per cell, 2: square of the KD exponential
                uint32_t qb[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                uint32_t eb[4] = {0u, 0u, 0u, 0u};
                if (act) {
                    const int v = lane + 32 * j;
                    const uint4 rs = load_s(v), rt = load_t(v);
                    uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
                    if constexpr (TAB) { o0 = lds128_ep(s_ov + v * 16, 0u); o1 = lds128_ep(s_ov + v * 16 + cp.vpc * 16, 0u); }
                    const uint32_t ws[4] = {rs.x, rs.y, rs.z, rs.w}, wt[4] = {rt.x, rt.y, rt.z, rt.w};
#ifndef KD_PR_NOCAND      // timing experiment only (wrong results)
                    if constexpr (PER_ROW) {
                        const uint32_t m4 = Pair<T>::vmax(Pair<T>::vmax(rt.x, rt.y), Pair<T>::vmax(rt.z, rt.w));
                        if (Pair<T>::any_ge(m4, thrp)) cand |= 1u << j;
                    }
#endif
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
                        if constexpr (CE_X == 1) accCE = add2(accCE, ex2_mufu2(fma2(Pair<T>::widen(ws[i]), k1p, nm1)));
                        if constexpr (CE_X == 2) accCE = fma2(e2, e2, accCE);          // (e * 2^15)^2: scaled back below
                        const u64 b2 = fma2(Pair<T>::widen(wt[i]), k2p, nlt);        // log2 tp (table mode) / relative to mt_w otherwise
                        float e_lo, e_hi;
                        if constexpr (TAB) upk(fma2(ovp[i], pushp, b2), e_lo, e_hi);
                        else upk(b2, e_lo, e_hi);
                        const u64 t2 = pk(ex2(e_lo), ex2(e_hi));
                        const u64 q2 = TAB ? add2(t2, ovp[i]) : t2;
                        accQ = add2(accQ, q2);
                        if constexpr (KLLT) {      // power sums T_2, T_3 and cross sums U_0, U_2, U_3 (U_1 is accP, T_1 is accQ)
                            const u64 tt = mul2(t2, t2), ttt = mul2(tt, t2);
                            accT2 = add2(accT2, tt);
                            accT3 = add2(accT3, ttt);
                            accU0 = add2(accU0, a2);
                            accU2 = fma2(tt, a2, accU2);
                            accU3 = fma2(ttt, a2, accU3);
                        } else {
                            accR = fma2(t2, b2, accR);
                        }
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
            const bool poly_ab = TAB && KD_POLY_AB > 0 && (fmaf(min_w, k2, esh - mk) >= kPolyFloor);
            bool token_given = false;
            auto run_ab = [&](auto ce_tag) {
                // the other pipeline may start its pass AB once this warp has `tok_after` iterations left (0 = strict alternation)
                const int j_split = max(0, n_full - cp.tok_after);
#pragma unroll kAbUnroll
                for (int j = 0; j < j_split; ++j) ab_body(j, true, std::false_type{}, std::false_type{}, ce_tag);
                if (NPIPE == 2 && cp.tok_after > 0) {
                    if (lane == 0) mbar_arrive_cta(tok_other);
                    token_given = true;
                }
#pragma unroll kAbUnroll
                for (int j = j_split; j < n_full; ++j) ab_body(j, true, std::false_type{}, std::false_type{}, ce_tag);
                if (n_tail) ab_body(n_full, lane < n_tail, std::true_type{}, std::false_type{}, ce_tag);
            };
            if (poly_ab) {
#pragma unroll 2
                for (int j = 0; j < n_full; ++j) ab_body(j, true, std::false_type{}, std::true_type{}, std::integral_constant<int, 0>{});
                if (n_tail) ab_body(n_full, lane < n_tail, std::true_type{}, std::true_type{}, std::integral_constant<int, 0>{});
            } else if constexpr (CE) {
                if (ce_extra) run_ab(std::integral_constant<int, 1>{});
                else if (ce_sq) run_ab(std::integral_constant<int, 2>{});
                else run_ab(std::integral_constant<int, 0>{});
            } else {
                run_ab(std::integral_constant<int, 0>{});
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
            if constexpr (CE) {
                S1_w = S_w;                         // T == 1: the KD exponentials are the CE's
                if (ce_extra || ce_sq) {
                    upk(accCE, x0, x1);
                    S1_w = warp_sum(x0 + x1) * (ce_sq ? (ECACHE ? 1.f / 1073741824.f : 1.f) : 1.f);
                }
            }
            float q_l, r_l, p_l;                       // this lane's shares of Q_w, R_w, P_w
            upk(accQ, x0, x1);
            q_l = x0 + x1;
            upk(accR, x0, x1);
            r_l = x0 + x1;
            upk(accP, x0, x1);
            p_l = x0 + x1;
            if constexpr (PER_ROW) {
                // Best / runner-up teacher cell of the warp's slice (largest value, lowest column first) from the few candidate
                // vectors the sweep flagged.  Keys = (order-preserving 16 value bits, ~local cell index) are unique: every flagged
                // lane folds its own vectors into a lane-local pair, then two warp-wide integer maxima merge the lanes.
                uint32_t lk1 = 0u, lk2 = 0u;
                while (cand) {
                    const int j = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int v = lane + 32 * j;
                    const uint4 rt = lds128(s_t + v * 16);
                    const uint32_t wt[4] = {rt.x, rt.y, rt.z, rt.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t bits = (i & 1) ? (wt[i >> 1] >> 16) : (wt[i >> 1] & 0xffffu);
                        const uint32_t key = ((bits ^ ((bits & 0x8000u) ? 0xffffu : 0x8000u)) << 16) | (0xffffu - (uint32_t)(v * 8 + i));
                        lk2 = max(lk2, min(lk1, key));
                        lk1 = max(lk1, key);
                    }
                }
                __syncwarp();
                top_k1 = __reduce_max_sync(0xffffffffu, lk1);
                top_k2 = __reduce_max_sync(0xffffffffu, lk1 == top_k1 ? lk2 : lk1);
            }
            Q_w = warp_sum(q_l);
            R_w = warp_sum(r_l) + poison;
            P_w = warp_sum(p_l) - esh * Q_w;          // sum q*(a + 15) -> sum q*a
            if constexpr (KLLT) {                     // the cross sums carry a' + 15 too: U_k = U_k' - 15 T_k  (T_0 = cells of the slice)
                upk(accT2, x0, x1);
                kl_T2 = warp_sum(x0 + x1);
                upk(accT3, x0, x1);
                kl_T3 = warp_sum(x0 + x1);
                kl_n = (float)(w_n * 8);
                upk(accU0, x0, x1);
                kl_U0 = warp_sum(x0 + x1) - esh * kl_n;
                upk(accU2, x0, x1);
                kl_U2 = warp_sum(x0 + x1) - esh * kl_T2;
                upk(accU3, x0, x1);
                kl_U3 = warp_sum(x0 + x1) - esh * kl_T3;
            }
        }
        // per-row LoCa: decode the slice's best / runner-up cell -> (value, global column, student logit in that column)
        float top_a2 = -INFINITY, top_xs1 = 0.f, top_xs2 = 0.f;
        int top_i1 = 0x7fffffff, top_i2 = 0x7fffffff;
        if constexpr (PER_ROW) {
            auto decode = [&](uint32_t key, float& val, int& col, float& xs) {
                if (key == 0u) return;
                const uint32_t u = key >> 16;
                const unsigned short bits = (unsigned short)(u ^ ((u & 0x8000u) ? 0x8000u : 0xffffu));
                const int cell = (int)(0xffffu - (key & 0xffffu));
                val = Elem<T>::to_float(*reinterpret_cast<const T*>(&bits));
                col = g_v0 * 8 + cell;
                xs = Elem<T>::to_float(*reinterpret_cast<const T*>(smem + (s_s - s_base) + cell * 2));
            };
            float a1 = -INFINITY;            // = mt_w
            decode(top_k1, a1, top_i1, top_xs1);
            decode(top_k2, top_a2, top_i2, top_xs2);
            __syncwarp();
            if (next_r < p.n_rows) issue_teacher(next_r);          // the candidate scan was the last reader of the teacher slice
        }
        KD_STAMP(3);
        // ---- the single reduction: push the warp's partial to every CTA of the cluster ---------------------------
        if (DYN) claimed = __shfl_sync(0xffffffffu, claimed, 0);
        if (lane < CS) {
            const uint32_t dst = map_to_cta(slots + (rank * WPP + pw) * SLOTB, lane);
            const uint32_t rbar = map_to_cta(bar_red1, lane);
            st_async_v4(dst, __float_as_uint(m_w), __float_as_uint(S_w), __float_as_uint(Q_w), __float_as_uint(R_w - P_w), rbar);
            if constexpr (TAB) {
                st_async_v2(dst + 16, __float_as_uint(min_w), (uint32_t)claimed, rbar);
                if constexpr (CE) st_async_v2(dst + 32, __float_as_uint(S1_w), __float_as_uint(xs_ce), rbar);
            } else if constexpr (PER_ROW) {
                st_async_v4(dst + 16, __float_as_uint(min_w), (uint32_t)claimed, __float_as_uint(mt_w), (uint32_t)top_i1, rbar);
                st_async_v4(dst + 32, __float_as_uint(top_a2), (uint32_t)top_i2, __float_as_uint(top_xs1), __float_as_uint(top_xs2), rbar);
                st_async_v2(dst + 48, __float_as_uint(xt_lab), __float_as_uint(xs_lab), rbar);
            } else if constexpr (KLLT) {
                // (word [3] was pushed as R_w - P_w with R_w = 0: the fold reads U_1 = -word[3])
                st_async_v4(dst + 16, __float_as_uint(kl_T2), (uint32_t)claimed, __float_as_uint(mt_w), __float_as_uint(kl_T3), rbar);
                st_async_v4(dst + 32, __float_as_uint(kl_U0), __float_as_uint(kl_U2), __float_as_uint(kl_U3), __float_as_uint(kl_n), rbar);
            } else {
                st_async_v4(dst + 16, 0u, (uint32_t)claimed, __float_as_uint(mt_w), 0u, rbar);
            }
        }
        // per-row LoCa: what the fold derives for the row's own two override cells (label -> v1, then runner-up -> v2, P1:184-185
        // applied per row); kept by the folding warp for its deferred work and for the second fold of a clamped row
        bool pr_same = false;
        float pr_v1 = NaN, pr_v2 = NaN, pr_xlogy = 0.f;
        float pr_tl = 0.f, pr_bl = 0.f, pr_al = 0.f, pr_tk = 0.f, pr_bk = 0.f, pr_ak = 0.f;
        int pr_k = 0x7fffffff;
        // Sweep and pass B' treated both cells as q = tp: take that out of (row sum, Q_r) and put the overrides in.  With
        // m = [a >= thr2]: the cell's terms are q*(lq - max(a, thr2)) and q*m; xlogy(v, v) rides in pr_xlogy (log2 units).
        auto patch_q = [&](float Qv) -> float {
            const float ml = (pr_al >= thr2) ? 1.f : 0.f, mkk = (pr_ak >= thr2) ? 1.f : 0.f;
            const float dq = -pr_tl * ml + pr_v2 * mkk;
            const float dq2 = -pr_tk * mkk + pr_v1 * ml;
            return Qv + dq + (pr_same ? 0.f : dq2);
        };
        auto patch_rowsum = [&](float rowsum) -> float {
            const float acl = fmaxf(pr_al, thr2), ack = fmaxf(pr_ak, thr2);
            const float dr = -pr_tl * (pr_bl - acl) - pr_v2 * ack + pr_xlogy;
            const float dr2 = -pr_tk * (pr_bk - ack) - pr_v1 * acl;
            return rowsum + dr + (pr_same ? 0.f : dr2);
        };
        // the two gradient cells of the row's own overrides: g*(sp*Q - ov*m)
        auto patch_grads = [&](float Qv) {
            const float gQv = g * Qv;
            bcast[5] = __int_as_float(pr_k);
            bcast[6] = fmaf(gQv, ex2(pr_al), -g * pr_v1 * ((pr_al >= thr2) ? 1.f : 0.f));
            bcast[8] = fmaf(gQv, ex2(pr_ak), -g * pr_v2 * ((pr_ak >= thr2) ? 1.f : 0.f));
        };
        constexpr int SPL = (NSLOT + 31) / 32;              // slots per lane in a fold (the last ones may not exist: neutral values)
        float pr_rowT = 0.f;                                // per-row LoCa, loss warp: sum_v tp*(b - a) up to the lse terms
        if constexpr (PER_ROW) {
            // Per-row LoCa: THREE warps of the pipeline share the fold — its dependent chain (redux, butterflies, MUFU), not its
            // instruction count, is what every warp of the pipeline waits for (measured: 1.35 us in one warp against 0.56 us for
            // the plain fold).  warp 0: student statistics, then Q_r and the two gradient cells; warp 1: teacher statistics, the
            // row's runner-up cell and the calibration scalars (handed over through the broadcast words + a two-warp barrier);
            // warp 2: the row's loss term, finished after the broadcast.
            if (pw < 3) {
                int lab_slot = 0;
                if (pw == 1) {      // the label logits sit in the slot of the warp that owns the column (known before the wait)
                    const int vl = max(label, 0) >> 3, rank_o = vl / cp.vpc;
                    const int cta_n_o = min(cp.vpc, cp.nvec - rank_o * cp.vpc);
                    const int w_chunk_o = ((cta_n_o + WPP * 32 - 1) / (WPP * 32)) * 32;
                    lab_slot = rank_o * WPP + (vl - rank_o * cp.vpc) / w_chunk_o;
                }
                mbar_wait_cta(bar_red1, ph1);
                if (pw == 0) {
                    KD_STAMP(7);
                    if (lane == 0) mbar_expect_tx(bar_red1, RED1_BYTES);      // arm the next row's phase
                    float4 sa[SPL];
                    float mnv[SPL];
                    int claim0 = 0;
                    float m_s = -INFINITY, mn = INFINITY;
#pragma unroll
                    for (int i = 0; i < SPL; ++i) {
                        const int idx = lane + 32 * i;
                        sa[i] = make_float4(-INFINITY, 0.f, 0.f, 0.f);
                        mnv[i] = INFINITY;
                        if (NSLOT % 32 == 0 || idx < NSLOT) {
                            sa[i] = slot_v[SLOTV * idx];
                            const float4 sb = slot_v[SLOTV * idx + 1];
                            mnv[i] = sb.x;
                            if (i == 0) claim0 = __float_as_int(sb.y);
                        }
                        m_s = fmaxf(m_s, sa[i].x);
                        mn = fminf(mn, mnv[i]);
                    }
                    m_s = warp_max_redux(m_s);
                    mn = warp_min_redux(mn);
                    float s_sum = 0.f;
#pragma unroll
                    for (int i = 0; i < SPL; ++i)
                        if (sa[i].x != -INFINITY) s_sum = fmaf(sa[i].y, ex2((sa[i].x - m_s) * k2), s_sum);
                    s_sum = warp_sum(s_sum);
                    const float lse2_s = fmaf(m_s, k2, __log2f(s_sum));
                    int claim_all = 0;
                    if (DYN) claim_all = __shfl_sync(0xffffffffu, claim0, 0);     // slot 0 = (rank 0, warp 0)
                    const float a_min = fmaf(mn, k2, -lse2_s);
                    const bool all_kept = a_min >= thr2 + 1e-3f * (1.f + fabsf(thr2));       // NaN -> masked path
                    named_bar_sync(bar3, 64);                                  // warp 1's words are in place
                    lse2_t = bcast[1];
                    pr_k = __float_as_int(bcast[5]);
                    pr_v1 = bcast[9]; pr_v2 = bcast[10]; pr_tl = bcast[11]; pr_bl = bcast[12]; pr_tk = bcast[13]; pr_bk = bcast[14];
                    pr_al = fmaf(bcast[15], k2, -lse2_s);
                    pr_ak = fmaf(bcast[16], k2, -lse2_s);
                    pr_same = (pr_k == label);
                    const float Q = patch_q(1.f);                              // sum_v tp = 1
                    if (lane == 0) {
                        bcast[0] = lse2_s;
                        bcast[2] = Q;
                        bcast[3] = all_kept ? 1.f : 0.f;
                        bcast[4] = __int_as_float(claim_all);
                        if (all_kept) patch_grads(Q);
                    }
                } else if (pw == 1) {
                    float tw[SPL];
                    float4 sb[SPL], sc[SPL];
                    float m_t = -INFINITY;
#pragma unroll
                    for (int i = 0; i < SPL; ++i) {
                        const int idx = lane + 32 * i;
                        tw[i] = 0.f;
                        sb[i] = make_float4(INFINITY, 0.f, -INFINITY, __int_as_float(0x7fffffff));
                        sc[i] = make_float4(-INFINITY, __int_as_float(0x7fffffff), 0.f, 0.f);
                        if (NSLOT % 32 == 0 || idx < NSLOT) {
                            tw[i] = slot_v[SLOTV * idx].z;
                            sb[i] = slot_v[SLOTV * idx + 1];
                            sc[i] = slot_v[SLOTV * idx + 2];
                        }
                        m_t = fmaxf(m_t, sb[i].z);
                    }
                    const float4 lab4 = slot_v[SLOTV * lab_slot + 3];
                    m_t = warp_max_redux(m_t);
                    float t_sum = 0.f;
#pragma unroll
                    for (int i = 0; i < SPL; ++i)
                        if (sb[i].z != -INFINITY) t_sum = fmaf(tw[i], ex2((sb[i].z - m_t) * k2), t_sum);
                    // The row's runner-up teacher cell: merge the slots' (best, runner-up) pairs — largest value, lowest column
                    // first.  Branch-free lane-local merge (one basic block with the statistics: the scheduler interleaves the
                    // two chains), then two (value, column) arg-max rounds over the warp with integer redux.
                    float b1a = -INFINITY, b2a = -INFINITY, b1x = 0.f, b2x = 0.f;
                    int b1i = 0x7fffffff, b2i = 0x7fffffff;
                    auto push = [&](float a, int ci, float xs) {
                        const bool p1 = Top2::before(a, ci, b1a, b1i), p2 = Top2::before(a, ci, b2a, b2i);
                        b2a = p1 ? b1a : (p2 ? a : b2a);
                        b2i = p1 ? b1i : (p2 ? ci : b2i);
                        b2x = p1 ? b1x : (p2 ? xs : b2x);
                        b1a = p1 ? a : b1a;
                        b1i = p1 ? ci : b1i;
                        b1x = p1 ? xs : b1x;
                    };
#pragma unroll
                    for (int i = 0; i < SPL; ++i) {
                        push(sb[i].z, __float_as_int(sb[i].w), sc[i].z);
                        push(sc[i].x, __float_as_int(sc[i].y), sc[i].w);
                    }
                    t_sum = warp_sum(t_sum);
                    const int key1 = __reduce_max_sync(0xffffffffu, fkey(b1a));
                    const bool h1 = fkey(b1a) == key1;
                    const int idx1 = __reduce_min_sync(0xffffffffu, h1 ? b1i : 0x7fffffff);
                    const bool win = h1 && b1i == idx1;
                    const float ca = win ? b2a : b1a, cx = win ? b2x : b1x;
                    const int ci = win ? b2i : b1i;
                    const int key2 = __reduce_max_sync(0xffffffffu, fkey(ca));
                    const bool h2 = fkey(ca) == key2;
                    const int kcol = __reduce_min_sync(0xffffffffu, h2 ? ci : 0x7fffffff);
                    const unsigned holder = __ballot_sync(0xffffffffu, h2 && ci == kcol);
                    const float xs_k = __shfl_sync(0xffffffffu, cx, holder ? __ffs(holder) - 1 : 0);
                    const float xt_k = fkey_inv(key2);
                    const float l2t = fmaf(m_t, k2, __log2f(t_sum));
                    float v1, v2;
                    loca_scalars(lab4.x, xt_k, l2t, k2, p.alpha, v1, v2);
                    if (label < 0) { v1 = NaN; v2 = NaN; }            // the reference's gather raises (P1:166): poison the row
                    const float bl = fmaf(lab4.x, k2, -l2t), bk = fmaf(xt_k, k2, -l2t);
                    if (lane == 0) {
                        bcast[1] = l2t;
                        bcast[5] = __int_as_float(kcol);
                        bcast[9] = v1; bcast[10] = v2;
                        bcast[11] = ex2(bl); bcast[12] = bl;
                        bcast[13] = ex2(bk); bcast[14] = bk;
                        bcast[15] = lab4.y; bcast[16] = xs_k;
                    }
                    __syncwarp();
                    named_bar_arrive(bar3, 64);
                } else {
                    // sum_v tp*(b - a) = X / t_sum + (lse2_s - lse2_t) with X = sum_w 2^((mt_w - m_t) k2) * [ (R_w - P_w) + (mt_w - m_w) k2 * T_w ]
                    float m_t = -INFINITY;
                    float4 sa[SPL];
                    float mtv[SPL];
#pragma unroll
                    for (int i = 0; i < SPL; ++i) {
                        const int idx = lane + 32 * i;
                        sa[i] = make_float4(-INFINITY, 0.f, 0.f, 0.f);
                        mtv[i] = -INFINITY;
                        if (NSLOT % 32 == 0 || idx < NSLOT) {
                            sa[i] = slot_v[SLOTV * idx];
                            mtv[i] = slot_v[SLOTV * idx + 1].z;
                        }
                        m_t = fmaxf(m_t, mtv[i]);
                    }
                    m_t = warp_max_redux(m_t);
                    float X = 0.f, t_sum = 0.f;
#pragma unroll
                    for (int i = 0; i < SPL; ++i) {
                        if (mtv[i] != -INFINITY) {
                            const float w = ex2((mtv[i] - m_t) * k2);
                            const float mki = __fmul_rn((sa[i].x == -INFINITY) ? 0.f : sa[i].x, k2);
                            t_sum = fmaf(sa[i].z, w, t_sum);
                            X = fmaf(w, fmaf(__fmul_rn(mtv[i], k2) - mki, sa[i].z, sa[i].w), X);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        X += __shfl_xor_sync(0xffffffffu, X, o);
                        t_sum += __shfl_xor_sync(0xffffffffu, t_sum, o);
                    }
                    pr_rowT = X / t_sum;
                }
            }
        } else if (pw == 0) {
            // one warp per pipeline folds the NSLOT partials and broadcasts the row statistics through shared memory
            mbar_wait_cta(bar_red1, ph1);
            KD_STAMP(7);
            if (lane == 0) mbar_expect_tx(bar_red1, RED1_BYTES);      // arm the next row's phase
            float4 sa[SPL], sb[SPL];
            float m_s = -INFINITY, m_t = -INFINITY, mn = INFINITY;
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const int idx = lane + 32 * i;
                sa[i] = make_float4(-INFINITY, 0.f, 0.f, 0.f);
                sb[i] = make_float4(INFINITY, 0.f, -INFINITY, 0.f);
                if (NSLOT % 32 == 0 || idx < NSLOT) {
                    sa[i] = slot_v[SLOTV * idx];
                    sb[i] = slot_v[SLOTV * idx + 1];
                }
                m_s = fmaxf(m_s, sa[i].x);
                if constexpr (TAB) mn = fminf(mn, sb[i].x);
                else m_t = fmaxf(m_t, sb[i].z);
            }
            m_s = warp_max_redux(m_s);
            if constexpr (TAB) mn = warp_min_redux(mn);
            else m_t = warp_max_redux(m_t);
            float s_sum = 0.f, t_sum = 0.f, Q = 0.f, A = 0.f;
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                if (sa[i].x != -INFINITY) s_sum = fmaf(sa[i].y, ex2((sa[i].x - m_s) * k2), s_sum);
                if constexpr (TAB) {
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
                if constexpr (TAB) {
                    Q += __shfl_xor_sync(0xffffffffu, Q, o);
                    A += __shfl_xor_sync(0xffffffffu, A, o);
                } else {
                    t_sum += __shfl_xor_sync(0xffffffffu, t_sum, o);
                }
            }
            const float lse2_s = fmaf(m_s, k2, __log2f(s_sum));
            [[maybe_unused]] float lse1 = 0.f, ce_row = 0.f;
            if constexpr (CE) {
                if (ce_on) {      // log-sum-exp at temperature 1 from the slices' S1_w, and the logit at the next-token id
                    float s1 = 0.f;
#pragma unroll
                    for (int i = 0; i < SPL; ++i) {
                        const int idx = lane + 32 * i;
                        if ((NSLOT % 32 == 0 || idx < NSLOT) && sa[i].x != -INFINITY)
                            s1 = fmaf(slot_v[SLOTV * idx + 2].x, ex2((sa[i].x - m_s) * k1), s1);
                    }
                    s1 = warp_sum(s1);
                    lse1 = fmaf(m_s, k1, __log2f(s1));
                    const int vl = max(ce_c, 0) >> 3, rank_o = vl / cp.vpc;
                    const int cta_n_o = min(cp.vpc, cp.nvec - rank_o * cp.vpc);
                    const int w_chunk_o = ((cta_n_o + WPP * 32 - 1) / (WPP * 32)) * 32;
                    const float xs_c = slot_v[SLOTV * (rank_o * WPP + (vl - rank_o * cp.vpc) / w_chunk_o) + 2].y;
                    ce_row = ce_bad ? NaN : (lse1 - xs_c * k1) * 0.69314718055994531f;
                }
            }
            float row;
            if constexpr (TAB) {
                row = fmaf(lse2_s, Q, A);
            } else {
                lse2_t = fmaf(m_t, k2, __log2f(t_sum));
                row = 0.f;
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const float mki = __fmul_rn((sa[i].x == -INFINITY) ? 0.f : sa[i].x, k2);
                    const float mti = __fmul_rn((sb[i].z == -INFINITY) ? 0.f : sb[i].z, k2);
                    const float sc_i = ex2(mti - lse2_t);             // this warp's teacher partials are relative to its mt_w
                    if constexpr (KLLT) {
                        // tp = sc_i * t', a = a' + c:  sum e^{tp} = n + s T1 + s^2 T2 / 2 + s^3 T3 / 6,  sum e^{tp} tp = s T1 + s^2 T2 + s^3 T3 / 2,
                        // sum e^{tp} a = (U0 + c n) + s (U1 + c T1) + s^2 (U2 + c T2) / 2 + s^3 (U3 + c T3) / 6      (natural-log row term below)
                        const int idx = lane + 32 * i;
                        const bool live = (NSLOT % 32 == 0 || idx < NSLOT);       // (a missing slot's neutral word [4] is +inf: the LoCa minimum)
                        float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live) sc = slot_v[SLOTV * idx + 2];
                        const float c = mki - lse2_s, T1 = sa[i].z, U1 = -sa[i].w, T2 = live ? sb[i].x : 0.f, T3 = live ? sb[i].w : 0.f;
                        const float s2h = 0.5f * sc_i * sc_i, s3s = s2h * sc_i * (1.f / 3.f);
                        Q += sc.w + fmaf(sc_i, T1, fmaf(s2h, T2, s3s * T3));
                        const float a_tp = fmaf(sc_i, T1, fmaf(2.f * s2h, T2, 3.f * s3s * T3));
                        const float a_a = fmaf(c, sc.w, sc.x) + fmaf(sc_i, fmaf(c, T1, U1), fmaf(s2h, fmaf(c, T2, sc.y), s3s * fmaf(c, T3, sc.z)));
                        row += fmaf(-0.69314718055994531f, a_a, a_tp);
                    } else {
                        Q = fmaf(sc_i, sa[i].z, Q);
                        row = fmaf(sc_i, sa[i].w + ((mti - lse2_t) - (mki - lse2_s)) * sa[i].z, row);
                    }
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
            if constexpr (TAB) {
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
                if constexpr (CE) {
                    bcast[5] = lse1;
                    if (rank == 0) {
                        p.ceaux[r] = make_float2(lse1, ce_row);
                        if (ce_bad) atomicOr(p.status, KD_STATUS_BAD_LABEL);
                    }
                }
            }
        }
        named_bar_sync(bar1, WPP * 32);
        KD_STAMP(4);
        const float lse2_s = bcast[0];
        if constexpr (!TAB) lse2_t = bcast[1];
        float Q = bcast[2];
        const bool all_kept = bcast[3] != 0.f;
        [[maybe_unused]] float lse1_row = 0.f;
        if constexpr (CE) lse1_row = bcast[5];
        int64_t after_next = 0;
        if (DYN) after_next = __float_as_int(bcast[4]);
        ph1 ^= 1;
        if constexpr (PER_ROW) {
            // Off the critical path (the other warps are already in their gradient sweep): the row's loss term and what the
            // backward re-emission needs.
            if (pw == 2) {
                pr_k = __float_as_int(bcast[5]);
                pr_v1 = bcast[9]; pr_v2 = bcast[10]; pr_tl = bcast[11]; pr_bl = bcast[12]; pr_tk = bcast[13]; pr_bk = bcast[14];
                pr_al = fmaf(bcast[15], k2, -lse2_s);
                pr_ak = fmaf(bcast[16], k2, -lse2_s);
                pr_same = (pr_k == label);
                if (rank == 0 && lane == 0) {
                    if (all_kept) {
                        pr_xlogy = own_xlogy2(pr_v1, pr_v2, pr_same, label);
                        p.rowaux[r] = make_float4(lse2_t, lse2_s, Q, patch_rowsum(pr_rowT + (lse2_s - lse2_t)));
                    }
                    kd_token_record rec;
                    rec.label = label; rec.v1 = pr_v1; rec.k = pr_k; rec.v2 = pr_v2;
                    p.records[r] = rec;                       // for the backward re-emission (kd_logit_kd_regrad)
                    if (label < 0) atomicOr(p.status, KD_STATUS_BAD_LABEL);
                }
            }
        }

        // ================= pass B' (rare): exact masked sums for rows the sp >= eps clamp touches =================
        // With the exponent cache the exact mask is applied to the parked q right here (q <- q*[a >= thr2], written back to
        // tensor memory), so the gradient pass of a clamped row is the same cached, MUFU-free sweep as for an unclamped row and
        // the student slice can be re-armed before the second reduction.
        bool student_refilled = false;
        bool q_masked = false;
        if constexpr (CLAMP) {
            if (!all_kept) {
                constexpr bool WRITE_BACK = KD_ECACHE;
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
#pragma unroll kBpUnroll
                for (int j = 0; j < n_full; ++j) bp_body(j, true, std::false_type{});
                if (n_tail) bp_body(n_full, lane < n_tail, std::true_type{});
                if constexpr (WRITE_BACK) {
                    tmem_wait_st();
                    q_masked = true;
                    if (want_grad && !ce_extra) {            // nothing reads the student slice any more: the cached gradient sweep will run
                        __syncwarp();
                        if (next_r < p.n_rows) issue_student(next_r);
                        student_refilled = true;
                    }
                }
                float x0, x1;
                upk(accM, x0, x1);
                float M_w = warp_sum(x0 + x1);
                upk(accN, x0, x1);
                float N_w = warp_sum(x0 + x1) + thr2 * (Q_w - M_w);
                float R_t = R_w;
                if constexpr (PER_ROW) {        // the warp's sums are relative to its teacher reference maximum: bring them to scale
                    const float d = mtk - lse2_t, scw = ex2(d);
                    R_t = scw * fmaf(d, Q_w, R_w);
                    N_w *= scw;
                    M_w *= scw;
                }
                if (lane < CS)
                    st_async_v2(map_to_cta(slots + (rank * WPP + pw) * SLOTB + RED2_OFF, lane), __float_as_uint(R_t - N_w),
                                __float_as_uint(M_w), map_to_cta(bar_red2, lane));
                if (pw == 0) {
                    mbar_wait_cta(bar_red2, ph2);
                    if (lane == 0) mbar_expect_tx(bar_red2, RED2_BYTES);
                    float rr = 0.f, qq = 0.f;
#pragma unroll
                    for (int i = 0; i < (NSLOT + 31) / 32; ++i) {
                        const int idx = lane + 32 * i;
                        if (NSLOT % 32 == 0 || idx < NSLOT) {
                            const float4 sb = slot_v[SLOTV * idx + SLOTV - 1];
                            rr += sb.z;
                            qq += sb.w;
                        }
                    }
                    rr = warp_sum(rr);
                    qq = warp_sum(qq);
                    if constexpr (PER_ROW) {
                        pr_xlogy = own_xlogy2(pr_v1, pr_v2, pr_same, label);
                        rr = patch_rowsum(rr);
                        qq = patch_q(qq);
                    }
                    if (lane == 0) {
                        bcast[2] = qq;
                        if constexpr (PER_ROW) patch_grads(qq);
                        if (rank == 0) p.rowaux[r] = make_float4(lse2_t, lse2_s, qq, rr);
                    }
                }
                named_bar_sync(bar2, WPP * 32);
                Q = bcast[2];
                ph2 ^= 1;
            }
        }
        int pr_kcol = 0x7fffffff;
        float pr_gl = 0.f, pr_gk = 0.f;
        if constexpr (PER_ROW) {
            pr_kcol = __float_as_int(bcast[5]);
            pr_gl = bcast[6];
            pr_gk = bcast[8];
        }

        // ================= pass C: gradient, written straight to HBM ===========================================
        if (want_grad) {
            const float gQ = g * Q;
            const bool fold = (g > 0.f) && (gQ > 0.f) && (gQ < INFINITY);
            // fold: g*Q rides in the exponent, 2^(a + log2 gQ) - g*q;  otherwise gQ * 2^a - g*q
            const float lgQ = fold ? __log2f(gQ) : 0.f;
            const u64 nls = pk(lgQ - lse2_s, lgQ - lse2_s);
            const float thrC = thr2 + lgQ;
            const float gq = (TAB || KLLT) ? g : g * ex2(mtk - lse2_t);         // no phase A: the parked values are relative to mt_w
            // log_target: the parked t' becomes w = e^{tp} = 2^(t' * sigma_w * log2 e), exactly, cell by cell
            [[maybe_unused]] const float sl2e = ex2(mtk - lse2_t) * 1.4426950408889634f;
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
                        if constexpr (STREAM_HINTS) stg_stream_cs(grow + (int64_t)v * 8, make_uint4(out[0], out[1], out[2], out[3]));
                        else stg_stream(grow + (int64_t)v * 8, make_uint4(out[0], out[1], out[2], out[3]));
                    }
                    if constexpr (decltype(tail_tag)::value) __syncwarp();
                };
#pragma unroll 2
                for (int j = 0; j < n_full; ++j) c_body(j, true, std::false_type{});
                if (n_tail) c_body(n_full, lane < n_tail, std::true_type{});
            };
            // exponent cache (unclamped rows, and clamped rows whose q was masked by pass B'): no shared memory, no MUFU
            // ce_tag: the fused cross-entropy at a KD temperature != 1 adds g_ce * 2^(xs log2 e - lse1) per cell, from the raw
            // student slice (which then stays in shared memory until this sweep is over)
            auto sweepC_cached_impl = [&](float ce, auto ce_tag, float gce_row) {
                constexpr int CE_X = decltype(ce_tag)::value;          // 0: no CE term here, 1: from the raw student slice, 2: e^2 from the cache
                const u64 cep = pk(ce, ce);
                [[maybe_unused]] const u64 k1c = pk(k1, k1), nl1 = pk(-lse1_row, -lse1_row), gcp = pk(gce_row, gce_row);
                auto c_step = [&](int j, auto nv_tag, bool act, auto tail_tag) {
                    constexpr int NV = decltype(nv_tag)::value;       // vectors per tensor-memory wait
                    uint32_t qb[NV][8], eb[NV][4];
                    [[maybe_unused]] uint4 rsv[NV];
#pragma unroll
                    for (int u = 0; u < NV; ++u) {
                        tmem_ld8(tq + 8 * (j + u), qb[u]);
                        tmem_ld4(tq + kTmE + 4 * (j + u), eb[u]);
                        if constexpr (CE_X == 1) {
                            rsv[u] = make_uint4(0, 0, 0, 0);
                            if (act) rsv[u] = load_s(lane + 32 * (j + u));
                        }
                    }
                    tmem_wait_ld();
                    if (act) {
#pragma unroll
                        for (int u = 0; u < NV; ++u) {
                            uint32_t out[4];
                            [[maybe_unused]] const uint32_t wsv[4] = {CE_X == 1 ? rsv[u].x : 0u, CE_X == 1 ? rsv[u].y : 0u, CE_X == 1 ? rsv[u].z : 0u, CE_X == 1 ? rsv[u].w : 0u};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 ef = __half22float2(*reinterpret_cast<const __half2*>(&eb[u][i]));
                                u64 q2 = pk(__uint_as_float(qb[u][2 * i]), __uint_as_float(qb[u][2 * i + 1]));
                                if constexpr (KLLT) q2 = ex2_mufu2(mul2(q2, pk(sl2e, sl2e)));
                                float n_lo, n_hi;
                                u64 o2 = fma2(pk(ef.x, ef.y), cep, mul2(q2, negg));
                                if constexpr (CE_X == 1) o2 = fma2(ex2_mufu2(fma2(Pair<T>::widen(wsv[i]), k1c, nl1)), gcp, o2);
                                if constexpr (CE_X == 2) o2 = fma2(mul2(pk(ef.x, ef.y), pk(ef.x, ef.y)), gcp, o2);      // gcp carries 2^(2 m_w k2 - 30 - lse1)
                                upk(o2, n_lo, n_hi);
                                out[i] = Pair<T>::narrow(n_lo, n_hi);
                            }
                            if constexpr (STREAM_HINTS) stg_stream_cs(grow + (int64_t)(lane + 32 * (j + u)) * 8, make_uint4(out[0], out[1], out[2], out[3]));
                            else stg_stream(grow + (int64_t)(lane + 32 * (j + u)) * 8, make_uint4(out[0], out[1], out[2], out[3]));
                        }
                    }
                    if constexpr (decltype(tail_tag)::value) __syncwarp();
                };
                using VM = std::integral_constant<int, KD_C_NV>;
                using V2 = std::integral_constant<int, 2>;
                using V1 = std::integral_constant<int, 1>;
                int j = 0;
                for (; j + KD_C_NV - 1 < n_full; j += KD_C_NV) c_step(j, VM{}, true, std::false_type{});
                for (; j + 1 < n_full; j += 2) c_step(j, V2{}, true, std::false_type{});
                if (j < n_full) c_step(j, V1{}, true, std::false_type{});
                if (n_tail) c_step(n_full, V1{}, lane < n_tail, std::true_type{});
            };
            // g*Q*sp = gQ * e' * 2^(m_w*k2 - 15 - lse2_s): valid for any sign of g and g*Q
            auto sweepC_cached = [&]() { sweepC_cached_impl(gQ * ex2((mk - kEScaleLog2) - lse2_s), std::integral_constant<int, 0>{}, 0.f); };
            using Yes = std::true_type;
            using No = std::false_type;
            // FMA-pipe exponentials in the common (folded, unmasked) sweep when log2(g*Q*sp) >= -126 for the whole slice
            const bool poly_c = TAB && KD_POLY_C > 0 && (fmaf(min_w, k2, lgQ - lse2_s) >= kPolyFloor);
            if ((all_kept || q_masked) && KD_ECACHE) {
                if (!student_refilled && !ce_extra) {
                    if (next_r < p.n_rows) issue_student(next_r);       // the cached sweep reads tensor memory only
                    student_refilled = true;
                }
                if constexpr (CE) {
                    static_assert(!CE || KD_ECACHE, "the fused cross-entropy relies on the cached gradient sweep");
                    // d CE / d xs = g_ce * (softmax(xs) - onehot(c)); a bad label poisons the row like the reference's gather would
                    const float gce_row = ce_bad ? NaN : g_ce;
                    if (!ce_on) sweepC_cached();
                    else if (ce_same_t)        // T == 1: softmax(xs) IS sp, the CE rides the same cached exponentials
                        sweepC_cached_impl((gQ + gce_row) * ex2((mk - kEScaleLog2) - lse2_s), std::integral_constant<int, 0>{}, 0.f);
                    else if (ce_sq)            // T == 2: softmax(xs) = (cached e)^2 * 2^(2 (m_w k2 - 15) - lse1)
                        sweepC_cached_impl(gQ * ex2((mk - kEScaleLog2) - lse2_s), std::integral_constant<int, 2>{},
                                           gce_row * ex2(2.f * (mk - kEScaleLog2) - lse1_row));
                    else
                        sweepC_cached_impl(gQ * ex2((mk - kEScaleLog2) - lse2_s), std::integral_constant<int, 1>{}, gce_row);
                    if (ce_c >= 0) {           // the one-hot term: the lane that stored the vector rewrites that one element
                        const int v = (ce_c >> 3) - g_v0;                   // warp-uniform
                        if (v >= 0 && v < w_n && lane == (v & 31)) {
                            T* cell = grow + v * 8 + (ce_c & 7);
                            *cell = Elem<T>::from_float(Elem<T>::to_float(*cell) - g_ce);
                        }
                    }
                } else {
                    sweepC_cached();
                }
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
        if constexpr (PER_ROW) {
            // the row's own override cells: the sweep wrote g*(sp*Q - tp*m) there; the owning lane (the one that stored the
            // vector) overwrites the two elements with g*(sp*Q - v*m)
            static_assert(!PER_ROW || KD_ECACHE, "per-row LoCa relies on the cached gradient sweep");
            if (want_grad && label >= 0) {
                T* grow = static_cast<T*>(p.grad) + r * p.ld_g + (int64_t)g_v0 * 8;
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const int c = which ? pr_kcol : label;
                    const int v = (c >> 3) - g_v0;                      // warp-uniform
                    if ((which == 1 || pr_kcol != label) && c != 0x7fffffff && v >= 0 && v < w_n && lane == (v & 31))
                        grow[v * 8 + (c & 7)] = Elem<T>::from_float(which ? pr_gk : pr_gl);
                }
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
    if (pw == 0 && lane == 0) reinterpret_cast<volatile float*>(smem + plan.bcast + pipe * BCASTB)[7] = 1.f;      // releases a partner waiting for our pass AB
    tmem_fence_before_sync();
    cluster_sync_all();     // nobody leaves while peers may still push into its shared memory
    tmem_fence_after_sync();
    if (warp == 1) tmem_dealloc_512(tmem_base);
}


// --------------------------------------------------------------------------------------------------------
template <int VARIANT, typename T, int CS, bool DYN, bool CE = false>
int launch_fused_d(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    constexpr int NPIPE = 2;
    auto kfn = row_emit_fused_kernel<VARIANT, T, CS, NPIPE, DYN, CE>;
    const int nvec = e.V / 8;
    const int vpc = (nvec + CS - 1) / CS;
    const FusedPlan<VARIANT, CS, NPIPE, CE> plan(vpc);
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
    cp.tok_after = 0;             // strict alternation: handing the token over 1-3 iterations early measured 1-3 % slower
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

template <int VARIANT, typename T, int CS, bool CE = false>
int launch_fused_c(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    if constexpr (CE) {      // the fused cross-entropy never shares rows with the streaming helper: static row striding only
        return launch_fused_d<VARIANT, T, CS, false, true>(e, st, query_only, capacity, launched, first_free);
    } else {
        if (e.row_queue) return launch_fused_d<VARIANT, T, CS, true>(e, st, query_only, capacity, launched, first_free);
        return launch_fused_d<VARIANT, T, CS, false>(e, st, query_only, capacity, launched, first_free);
    }
}

// cluster size: whichever of 8 / 9 puts more SMs to work for this shape (KD_CLUSTER_SIZE=8|9 forces one) — unless launch_emit asks for
// clusters of 8 because it will run the streaming helper on the SMs they leave free (per-row LoCa, see launch_emit)
int g_last_cluster_size = 8;
thread_local bool t_prefer_cluster8 = false;
template <int VARIANT, typename T, bool CE = false>
int launch_fused_v(const EmitParams& e, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    static int forced = -1;
    if (forced < 0) { const char* ev = getenv("KD_CLUSTER_SIZE"); forced = ev ? atoi(ev) : 0; }
    int cap8 = 0, cap9 = 0;
    const int rc8 = (forced == 9) ? KD_ERR_UNSUPPORTED : launch_fused_c<VARIANT, T, 8, CE>(e, st, true, &cap8, nullptr, nullptr);
    const int rc9 = (forced == 8) ? KD_ERR_UNSUPPORTED : launch_fused_c<VARIANT, T, 9, CE>(e, st, true, &cap9, nullptr, nullptr);
    const bool use9 = (rc9 == KD_OK) && (rc8 != KD_OK || (cap9 * 9 > cap8 * 8 && !(t_prefer_cluster8 && forced == 0)));
    if (!use9 && rc8 != KD_OK) return rc8;
    g_last_cluster_size = use9 ? 9 : 8;
    if (use9) return launch_fused_c<VARIANT, T, 9, CE>(e, st, query_only, capacity, launched, first_free);
    return launch_fused_c<VARIANT, T, 8, CE>(e, st, query_only, capacity, launched, first_free);
}

template <typename T>
int launch_fused_t(const EmitParams& e, int variant, cudaStream_t st, bool query_only, int* capacity, int* launched, int* first_free) {
    if (variant == KD_LOCA_REFERENCE && e.ce_labels) return launch_fused_v<KD_LOCA_REFERENCE, T, true>(e, st, query_only, capacity, launched, first_free);
    if (variant == KD_LOCA_REFERENCE) return launch_fused_v<KD_LOCA_REFERENCE, T>(e, st, query_only, capacity, launched, first_free);
    if (variant == KD_LOCA_PER_ROW) return launch_fused_v<KD_LOCA_PER_ROW, T>(e, st, query_only, capacity, launched, first_free);
    if (variant == KD_KL_LOG_TARGET) return launch_fused_v<KD_KL_LOG_TARGET, T>(e, st, query_only, capacity, launched, first_free);
    return launch_fused_v<KD_KL_PLAIN, T>(e, st, query_only, capacity, launched, first_free);
}

// KD_FUSED=0 selects the three-sweep cluster kernel everywhere (A/B comparisons); read once
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
    if (p.ce_labels && (variant != KD_LOCA_REFERENCE || !p.ce_inv_count || !p.ceaux || !p.status)) return false;
    if (variant == KD_LOCA_REFERENCE) {
        if (!(p.alpha >= 0.f && p.alpha <= 1.f)) return false;      // overrides outside [0, 1]: exact per-cell path
    } else if (variant == KD_LOCA_PER_ROW) {
        if (!p.labels || !p.records || !p.status) return false;     // labels in, token records out (backward re-emission)
    } else if (variant != KD_KL_PLAIN && variant != KD_KL_LOG_TARGET) {
        return false;
    }
    return dtype == KD_DTYPE_BF16 || dtype == KD_DTYPE_F16;
}

int fused_cluster_size() { return g_last_cluster_size; }
void fused_prefer_cluster8(bool on) { t_prefer_cluster8 = on; }

int launch_emit_fused(const EmitParams& p, int variant, int dtype, cudaStream_t stream, bool query_only, int* capacity,
                      int* clusters_launched, int* first_free_row, int* rows_in_flight) {
    g_last_cluster_size = 8;
    if (rows_in_flight) *rows_in_flight = 2;
    if (dtype == KD_DTYPE_BF16) return launch_fused_t<__nv_bfloat16>(p, variant, stream, query_only, capacity, clusters_launched, first_free_row);
    return launch_fused_t<__half>(p, variant, stream, query_only, capacity, clusters_launched, first_free_row);
}

}  // namespace kd
