// logit_kd.cuh — shapes, parameters and per-element math shared by the streaming (logit_kd.cu) and the
// cluster-resident (logit_kd_cluster.cu) logit-KD kernels.
//
// Math (restating P1:141-194 / P1:331-343 / FB:206-219 of the reference), all in the log2 domain:
//   k2 = log2(e)/T,  a = xs*k2 - lse2_s = log2 sp,  b = xt*k2 - lse2_t = log2 tp,  thr2 = log2(eps)
//   LoCa : q = override ? ov : 2^b ; keep = a >= thr2 ; row2 += q*(lq - max(a,thr2)) with lq = override ? 0 : b
//          (the row-independent sum of xlogy(ov,ov) is added once per row at finalisation)
//          Q += keep ? q : 0 ;  grad = g*(2^a * Q - (keep ? q : 0))
//   KL   : q = 2^b ; row2 += q*(b - a) ; Q += q ; grad = g*(2^a * Q - q)
//   KLLT : w = e^(2^b) ; rowNat += w*(2^b - a*ln2) ; Q += w ; grad = g*(2^a * Q - w)
#pragma once

#include "kd_common.cuh"

namespace kd {

struct EmitParams {
    const void* teacher;
    int64_t ld_t;
    const void* student;
    int64_t ld_s;
    const int64_t* labels;      // per-row variants
    const float* lse2_t_in;     // LoCa reference mode: from the teacher-stats pass
    const float* table;         // LoCa reference mode: V-entry override table (NaN = not overridden)
    void* grad;
    int64_t ld_g;
    float4* rowaux;             // per row: lse2_t, lse2_s, Q_r, row sum (log2 units; natural for KLLT)
    kd_token_record* records;   // per-row LoCa: {label, v1, k, v2} written for the backward re-emission (may be null)
    int* row_queue;             // shared row counter: the cluster kernel (120 SMs) and the streaming kernel running on the
                                // SMs a cluster-8 launch cannot reach pull rows from it; null = static striding
    int queue_tail_reserve;     // streaming helper: stop claiming when fewer than this many rows are left
    int helper_ctas;            // streaming helper: grid size (0 = default)
    uint32_t* status;
    const float* upstream;      // device scalar or nullptr
    int64_t n_rows;
    int V;
    float k2;                   // log2(e)/T
    float alpha;
    float thr2;                 // log2(eps)
    float gscale;               // T * inv_norm
    unsigned flags;
    // Fused student cross-entropy (the HF loss the reference adds at P1:194 / P1:260: mean over the valid shifted labels of
    // lse(xs) - xs[c], ignore_index -100, temperature 1) — computed from the student row the fused kernel already holds.
    const int64_t* ce_labels;   // per row: the NEXT token's id (already shifted by the caller) or -100; null = no CE
    const float* ce_inv_count;  // device scalar: 1 / number of valid ce_labels of the whole batch
    float ce_weight;            // loss = KD + ce_weight * CE; the emitted gradient carries upstream * both
    float2* ceaux;              // per row: {log2-sum-exp2 of xs*log2(e), CE term of the row (natural log; 0 when ignored)}
};
constexpr int64_t kCeIgnore = -100;

template <int VARIANT>
struct PointTerm {
    // one (teacher, student) cell in the accumulate sweep: returns the row-sum term and the Q term
    static __device__ __forceinline__ void accum(float xt, float xs, float ov, float k2, float lse2_t, float lse2_s,
                                                 float thr2, float& row, float& Q) {
        const float a = fmaf(xs, k2, -lse2_s);
        const float b = fmaf(xt, k2, -lse2_t);
        if constexpr (VARIANT == KD_LOCA_REFERENCE || VARIANT == KD_LOCA_PER_ROW) {
            const bool over = (ov == ov);           // not NaN
            const float q = over ? ov : ex2(b);
            const float lq = over ? 0.f : b;
            const float ac = fmaxf(a, thr2);
            row = fmaf(q, lq - ac, row);
            Q += (a >= thr2) ? q : 0.f;
        } else if constexpr (VARIANT == KD_KL_PLAIN) {
            const float q = ex2(b);
            row = fmaf(q, b - a, row);
            Q += q;
        } else {
            const float tp = ex2(b);
            const float w = ex2(tp * kLog2e);
            row = fmaf(w, fmaf(-a, kLn2, tp), row);
            Q += w;
        }
    }
    // the same cell in the emit sweep: gradient value g*(sp*Q - w), with gQ = g*Q
    static __device__ __forceinline__ float grad(float xt, float xs, float ov, float k2, float lse2_t, float lse2_s,
                                                 float thr2, float g, float gQ) {
        const float a = fmaf(xs, k2, -lse2_s);
        const float sp = ex2(a);
        const float b = fmaf(xt, k2, -lse2_t);
        if constexpr (VARIANT == KD_LOCA_REFERENCE || VARIANT == KD_LOCA_PER_ROW) {
            const bool over = (ov == ov);
            float q = over ? ov : ex2(b);
            q = (a >= thr2) ? q : 0.f;
            return fmaf(gQ, sp, -g * q);
        } else if constexpr (VARIANT == KD_KL_PLAIN) {
            return fmaf(gQ, sp, -g * ex2(b));
        } else {
            return fmaf(gQ, sp, -g * ex2(ex2(b) * kLog2e));
        }
    }
};

// per-row LoCa calibration scalars from (lse2_t, label logit, runner-up logit) — P1:166-185
__device__ __forceinline__ void loca_scalars(float x_label, float x_runner, float lse2_t, float k2, float alpha,
                                             float& v1, float& v2) {
    // softmax values never exceed 1; 2^(x*k2 - lse2) can by an ulp when the row is a near one-hot, which would turn
    // 1 - p_gt negative and v2 into a tiny negative number (found by tests/test_gpu_logit_kd.py::test_extreme_but_finite_logits)
    const float p_gt = fminf(ex2(fmaf(x_label, k2, -lse2_t)), 1.f);
    const float p_k = fminf(ex2(fmaf(x_runner, k2, -lse2_t)), 1.f);
    const float sc = alpha / (1.f - p_gt + p_k);
    v1 = 1.f - sc * (1.f - p_gt);     // sum_v tp taken as 1
    v2 = sc * p_k;
}

// host-side launchers implemented in logit_kd_cluster.cu; return KD_ERR_UNSUPPORTED when the shape does not fit.
// *first_free_row receives the first row index not statically owned by a cluster pipeline (initial queue value).
int launch_emit_cluster(const EmitParams& p, int variant, int dtype, cudaStream_t stream, int* clusters_launched,
                        int* ctas_per_cluster, int* first_free_row);
// shape the cluster launch would take (no launch): KD_ERR_UNSUPPORTED when the cluster kernel does not apply
// *static_rows = rows per cluster that are assigned statically (the row queue starts behind them)
int cluster_launch_shape(const EmitParams& p, int variant, int dtype, int* n_clusters, int* ctas_per_cluster, int* pipelines,
                         int* static_rows);


// two-sweep kernel with the q values parked in tensor memory (logit_kd_fused.cu): applies to KD_LOCA_REFERENCE with
// 0 <= alpha <= 1 and to KD_KL_PLAIN on 16-bit logits; same launch shape as the cluster kernel (clusters of 8, 2 pipelines)
bool fused_applies(const EmitParams& p, int variant, int dtype);
int fused_cluster_size();      // CTAs per cluster chosen by the last launch_emit_fused call (8 or 9)
void fused_prefer_cluster8(bool on);   // this thread's next launch_emit_fused calls take clusters of 8 where both sizes fit (launch_emit's hint)
int launch_emit_fused(const EmitParams& p, int variant, int dtype, cudaStream_t stream, bool query_only, int* capacity,
                      int* clusters_launched, int* first_free_row, int* rows_in_flight /* per cluster: 1 or 2 pipelines */);

}  // namespace kd
