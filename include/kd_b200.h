/*
 * kd_b200.h — C ABI of the B200-native KD-loss library (libkd_b200.so).
 *
 * The reference (shayekh00/Knowledge_Distillation_for_Sensory_Substitution_in_Multimodal_Models) has no
 * FFI layer: its boundary for this path is a set of Python methods on the KD LightningModule.  Each entry
 * point below names the reference method it replaces (paths relative to the reference root; P1 =
 * distillation/knowledge_distillation7b_double_trouble/phase1/OnlineKnowledgeDistillationLLavaOneVision.py,
 * LG = distillation/knowledge_distillation7b_logit_based/..., FB = ..._feature_based/...,
 * EV = evaluation/onevisionv3/...).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller unless stated;
 *   - the library never allocates or frees device memory and never synchronises the device;
 *   - process-wide state is limited to: the thread-local error string, per-device launch configuration cached on first use
 *     (occupancy of the cluster kernels, opt-in shared-memory size), tuning switches read ONCE from the environment
 *     (KD_FUSED, KD_CLUSTER_SIZE, KD_CLUSTER_CONFIG, KD_SPARE_SMS, KD_HELPER_RESERVE — A/B experiments and the test suite's
 *     way to reach the non-default kernels, never needed in production) and one
 *     internal non-blocking helper stream + event pair per device, used only where clusters of 8 would leave >= 20 SMs idle;
 *     nothing of it changes results;
 *   - every call takes the CUDA stream (a cudaStream_t passed as void*) and is graph-capture safe;
 *   - return value 0 = success, otherwise a KD_ERR_* code; kd_last_error() gives a thread-local message;
 *   - row-major 2-D tensors are described by (pointer, leading dimension in ELEMENTS).
 */
#ifndef KD_B200_H_
#define KD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KD_B200_ABI_VERSION 5

/* element types of logits / features (gradients are written in the same type as the student tensor) */
enum { KD_DTYPE_BF16 = 0, KD_DTYPE_F16 = 1, KD_DTYPE_F32 = 2 };

/* logit-KD variants */
enum {
    KD_LOCA_REFERENCE = 0, /* P1:141-194 as written: one override table shared by every row (P1:184-185)   */
    KD_LOCA_PER_ROW   = 1, /* per-row calibration (the reference applied token by token and averaged)        */
    KD_KL_PLAIN       = 2, /* F.kl_div(log_softmax(s/T), softmax(t/T)) * T^2 — P1:331-343, EV:261-273         */
    KD_KL_LOG_TARGET  = 3  /* same call with log_target=True on a probability target — FB:206-219, LG:115-122 */
};

/* error codes */
enum {
    KD_OK = 0,
    KD_ERR_INVALID_ARGUMENT = 1,
    KD_ERR_UNSUPPORTED = 2,
    KD_ERR_WORKSPACE_TOO_SMALL = 3,
    KD_ERR_CUDA = 4
};

/* bits of the device status word written by the logit-KD kernels (kd_logit_kd_status_offset) */
#define KD_STATUS_BAD_LABEL 1u /* a label outside [0, V): the reference raises (P1:166); loss/grad are NaN-poisoned */

/* flags of kd_logit_kd_fwd_bwd / kd_loca_emit */
#define KD_FLAG_NO_GRAD        1u /* forward only: do not write grad_student                                  */
#define KD_FLAG_FORCE_STREAMING 2u /* use the general multi-sweep kernel even where the cluster kernel applies  */

int kd_abi_version(void);
const char* kd_last_error(void);

/* Device properties the host side sizes its grids with (SM count, max active clusters ...). Host pointers. */
int kd_device_info(int* sm_count, int* max_smem_per_block, int* cluster4_max_active, int* cluster8_max_active);

/* ------------------------------------------------------------------------------------------------------
 * Logit KD — replaces compute_loca_loss (P1:141-194, LG:208-261; call sites P1:254, P1:258, LG:165), the KL
 * term of compute_vision_loss (P1:331-343; call P1:251, EV:236) and of compute_loss (FB:206-219; call FB:165).
 *
 *   loss = T^2 * inv_norm * sum_{r,v} pointwise(q[r,v], sp[r,v])            (one float, *loss_out)
 *   grad = upstream * T * inv_norm * (sp * Q_r - w)                          (written to grad_student)
 *
 * inv_norm = 1/(N*V) reproduces reduction='mean'; 1/B gives 'batchmean' (EV:130-136, P1:378-381).
 * `upstream` (device scalar, may be NULL = 1) is the incoming grad_output folded into the same pass.
 * teacher may have more columns than V (ld_t >= V) — the slice at P1:155 is a stride, not a copy.
 * labels: int64 [n_rows], required for the LoCa variants, must lie in [0, V).
 * loss_out: one float (rows are summed in double on the device before the final rounding); per-row sums
 * are kept in the workspace (kd_logit_kd_rowaux_offset).
 * ---------------------------------------------------------------------------------------------------- */
int kd_logit_kd_workspace_bytes(int64_t n_rows, int64_t V, size_t* bytes);
int kd_logit_kd_status_offset(int64_t n_rows, int64_t V, size_t* offset);   /* uint32 status word   */
int kd_logit_kd_rowaux_offset(int64_t n_rows, int64_t V, size_t* offset);   /* float4[n_rows]: lse2_t, lse2_s, Q_r, row loss */

int kd_logit_kd_fwd_bwd(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s,
                        const int64_t* labels, void* grad_student, int64_t ld_g, float* loss_out,
                        void* workspace, size_t workspace_bytes, int64_t n_rows, int64_t V, int dtype,
                        float T, float alpha, float eps, int variant, double inv_norm,
                        const float* upstream, unsigned flags, void* stream);

/* Phase-split entry points of the LoCa reference mode, for token-sharded execution (one exchange step):
 *   A. kd_loca_teacher_stats : rows -> per-token records {label, v1, k, v2} (16 B each) + lse2_t
 *   (caller all-gathers the records of every rank in global token order)
 *   B. kd_loca_build_table   : all records -> the V-entry override table (last token wins; k pass after labels)
 *   C. kd_loca_emit          : local rows + table -> loss partial (already scaled) + gradient
 *      (kd_loca_emit_ce      : the same + the student's cross-entropy on those rows)
 */
typedef struct {
    int32_t label;
    float v1;      /* 1 - s * (sum tp - p_gt)   (P1:184 rhs) */
    int32_t k;     /* runner-up class index      (P1:170-171) */
    float v2;      /* s * p_k                    (P1:185 rhs) */
} kd_token_record;

int kd_loca_teacher_stats(const void* teacher, int64_t ld_t, const int64_t* labels, kd_token_record* records,
                          float* lse2_t, uint32_t* status, int64_t n_rows, int64_t V, int dtype, float T,
                          float alpha, void* stream);
/* Phase A fused with the record all-gather: every record is stored into peer_records[g][row_offset + r] for all n_peers ranks
 * (device array of device pointers to each rank's [n_tokens_all] record buffer — peer-mapped memory, e.g. CUDA IPC / symmetric
 * memory over NVLink — including this rank's own).  The caller separates phase A from kd_loca_build_table with a cross-GPU
 * barrier; no NCCL all-gather is needed. */
int kd_loca_teacher_stats_peers(const void* teacher, int64_t ld_t, const int64_t* labels, kd_token_record* const* peer_records,
                                int n_peers, int64_t row_offset, float* lse2_t, uint32_t* status, int64_t n_rows, int64_t V,
                                int dtype, float T, float alpha, void* stream);
int kd_loca_table_workspace_bytes(int64_t V, size_t* bytes);
int kd_loca_build_table(const kd_token_record* records_all, int64_t n_tokens_all, int64_t V, float* table,
                        double* table_consts /* [2]: sum xlogy(ov,ov), sum ov */, void* workspace,
                        size_t workspace_bytes, void* stream);
int kd_loca_emit(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* lse2_t,
                 const float* table, const double* table_consts, void* grad_student, int64_t ld_g,
                 float* loss_out, void* rowaux /* float4[n_rows + 1]; the extra entry is scratch (row queue) */,
                 int64_t n_rows, int64_t V, int dtype,
                 float T, float eps, double inv_norm, const float* upstream,
                 const uint32_t* status /* phase A's status word (may be NULL): a bad label NaN-poisons the loss like P1:166 raises */,
                 unsigned flags, void* stream);

/* kd_loca_emit with the student's cross-entropy of the LOCAL rows fused in (token-sharded form of kd_logit_kd_ce_fwd_bwd below, which
 * documents the CE semantics): loss_out = KD partial + ce_weight * CE partial, ce_loss_out (may be NULL) = CE partial =
 * sum over the local valid rows of (lse(xs) - xs[c]) * (*ce_inv_count).  The caller chooses the divisor: 1 / (valid targets of ALL
 * ranks) makes the sum of the ranks' partials the cross-entropy of the concatenated batch.  ceaux = float2[n_rows] scratch;
 * status = phase A's status word, REQUIRED here (a ce_label outside [0, V) other than -100 sets KD_STATUS_BAD_LABEL in it).
 * Same shape limits as kd_logit_kd_ce_fwd_bwd (KD_ERR_UNSUPPORTED otherwise). */
int kd_loca_emit_ce(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* lse2_t,
                    const float* table, const double* table_consts, const int64_t* ce_labels, const float* ce_inv_count,
                    float ce_weight, void* grad_student, int64_t ld_g, float* loss_out, float* ce_loss_out,
                    void* rowaux /* float4[n_rows + 1] */, void* ceaux /* float2[n_rows] */, int64_t n_rows, int64_t V, int dtype,
                    float T, float eps, double inv_norm, const float* upstream, uint32_t* status, unsigned flags, void* stream);

/* Re-emission of the gradient for autograd's backward when the incoming grad_output differs from the scale that
 * was folded into the forward launch: grad = (*upstream) * T * inv_norm * (sp * Q_r - w) recomputed from the row
 * statistics the forward left in `workspace` (one streaming sweep, ONE rounding to the output type — rescaling the
 * already-rounded gradient would round twice).  The kernel returns immediately when *skip_if_one == 1.0f (device
 * scalar, may be NULL = never skip), so the common "grad_output is what forward assumed" case costs one empty launch. */
int kd_logit_kd_regrad(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, void* grad_student,
                       int64_t ld_g, const void* workspace, size_t workspace_bytes, int64_t n_rows, int64_t V,
                       int dtype, float T, float eps, int variant, double inv_norm, const float* upstream,
                       const float* skip_if_one, void* stream);

/* kd_logit_kd_regrad for the phase-split path: re-emits the gradient of the local rows from what kd_loca_emit / kd_loca_emit_ce left in
 * `rowaux` (and `ceaux`) and the override `table`, with *upstream as the new scale — ONE rounding to the output type, where scaling the
 * already rounded gradient (kd_scale_rows) would round twice.  ce_labels == NULL: no cross-entropy terms (a kd_loca_emit forward);
 * otherwise the same ce_labels / ce_inv_count / ce_weight / ceaux as the forward.  Returns at once when *skip_if_one == 1.0f. */
int kd_loca_regrad(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* table, const void* rowaux,
                   const int64_t* ce_labels, const float* ce_inv_count, float ce_weight, const void* ceaux, void* grad_student,
                   int64_t ld_g, int64_t n_rows, int64_t V, int dtype, float T, float eps, double inv_norm, const float* upstream,
                   const float* skip_if_one, void* stream);

/* As-written LoCa KL + the student's cross-entropy in ONE pass over the logits — SURVEY.md §8(f) rank 1.  The reference adds
 * `student_outputs.loss` (P1:194; again at P1:260 in phase 3), which the HF student computes at P1:238-240 from the SAME student
 * logits: shifted labels, ignore_index -100 (the datamodule pads with it, dataset/.../CustomSUNRGBDOneVisionDataModule.py:145-146),
 * mean over the valid positions, temperature 1.  Here the row the fused kernel already holds in shared memory yields it too:
 *   loss_out    = KD + ce_weight * CE            ce_loss_out (may be NULL) = CE = sum_valid (lse(xs) - xs[c]) * (*ce_inv_count)
 *   grad_student = upstream * ( dKD + ce_weight * (softmax(xs) - onehot(c)) * (*ce_inv_count) )   (rows with c == -100: dKD only)
 * ce_labels[r] = id of the token FOLLOWING row r (the caller shifts: labels[b, s + 1], -100 at the last position of a sequence
 * and wherever the target is masked); ce_inv_count = device scalar 1 / (number of valid ce_labels in the whole batch) — the host
 * never reads it.  A ce_label outside [0, V) other than -100 sets KD_STATUS_BAD_LABEL and NaN-poisons the losses
 * (F.cross_entropy raises).  Needs the fused kernel's shape (bf16 / fp16 logits, V % 8 == 0, alpha in [0, 1], 16-byte aligned
 * rows): KD_ERR_UNSUPPORTED otherwise — the caller then adds a separately computed CE.  With T == 1 the CE is free (its
 * exponentials are the KD's); otherwise valid rows pay one more ex2 per cell in each sweep, ignored rows nothing. */
int kd_logit_kd_ce_fwd_bwd(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const int64_t* labels,
                           const int64_t* ce_labels, const float* ce_inv_count, float ce_weight, void* grad_student,
                           int64_t ld_g, float* loss_out, float* ce_loss_out, void* workspace, size_t workspace_bytes,
                           int64_t n_rows, int64_t V, int dtype, float T, float alpha, float eps, double inv_norm,
                           const float* upstream, unsigned flags, void* stream);
/* kd_logit_kd_regrad for a forward made by kd_logit_kd_ce_fwd_bwd (same workspace) */
int kd_logit_kd_ce_regrad(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const int64_t* ce_labels,
                          const float* ce_inv_count, float ce_weight, void* grad_student, int64_t ld_g, const void* workspace,
                          size_t workspace_bytes, int64_t n_rows, int64_t V, int dtype, float T, float eps, double inv_norm,
                          const float* upstream, const float* skip_if_one, void* stream);


/* ------------------------------------------------------------------------------------------------------
 * Vocabulary-sharded logit KD (BASELINE.json configs[4]): every rank holds all n_rows rows but only
 * columns [col0, col0 + V_local) of both logit tensors.  Same reference methods as above (P1:141-194, P1:331-343,
 * FB:206-219); the per-row quantities that couple the shards cross ranks twice, the caller does the two collectives:
 *   kd_vshard_stats   -> stats[n_rows][12] floats  (local max/sum-exp of both rows, local best + runner-up teacher cells
 *                        with GLOBAL column indices, the teacher logit at the label if this shard owns it)
 *   [all-gather the stats blocks in shard order: stats_all[n_shards][n_rows][12]]
 *   kd_vshard_combine -> lse2[n_rows][2] = (lse2_t, lse2_s), token records (global indices, P1:166-185), rowconst[n_rows]
 *                        (per-row mode: xlogy(v,v) of the row's own override cells in log2 units; NaN for a bad label)
 *   [as-written mode: kd_loca_build_table(records, n_rows, V_total, ...) — every rank holds all rows]
 *   kd_vshard_accum   -> partial[n_rows][2] = (local part of the row sum in log2 units [natural for KD_KL_LOG_TARGET],
 *                        local part of Q_r)
 *   [all-reduce(SUM) Q_r over the shards]
 *   kd_vshard_emit    -> gradient of the local columns: upstream * T * inv_norm * (sp * Q_r - w)
 *   loss = T^2 * inv_norm * ( unit * sum over shards and rows of partial.row + ln2 * sum_r rowconst + n_rows * table_consts[0] ),
 *   unit = ln2 (1 for KD_KL_LOG_TARGET).  `table` is the full V_total-entry table (the kernels index it at col0 + c).
 * ---------------------------------------------------------------------------------------------------- */
#define KD_VSHARD_STAT_WORDS 12
int kd_vshard_stats(const void* teacher, int64_t ld_t, const void* student /* NULL: teacher half only */, int64_t ld_s,
                    const int64_t* labels /* may be NULL */,
                    int64_t col0, float* stats, int64_t n_rows, int64_t V_local, int dtype, float T, void* stream);
int kd_vshard_combine(const float* stats_all, int n_shards, int64_t n_rows, const int64_t* labels, int64_t V_total, float T,
                      float alpha, int variant, kd_token_record* records, float* lse2, float* rowconst, uint32_t* status,
                      void* stream);
int kd_vshard_accum(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* lse2,
                    const kd_token_record* records, const float* table, int64_t col0, float* partial, int64_t n_rows,
                    int64_t V_local, int dtype, float T, float eps, int variant, void* stream);
int kd_vshard_emit(const void* teacher, int64_t ld_t, const void* student, int64_t ld_s, const float* lse2, const float* Q_rows,
                   const kd_token_record* records, const float* table, int64_t col0, void* grad_student, int64_t ld_g,
                   int64_t n_rows, int64_t V_local, int dtype, float T, float eps, int variant, double inv_norm,
                   const float* upstream, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Student lm_head feeding the vocabulary-sharded KD — the producer of `student_outputs.logits` (P1:228-240; in the
 * reference the HF Qwen2 lm_head, 896 x 151 936), as a hand-written tcgen05 / tensor-memory GEMM:
 *   logits[n_rows, V_local] = hidden[n_rows, H] . weight[V_local, H]^T     (bf16 / fp16, fp32 accumulation, row-major,
 *                                                                          pitches ld_* in elements)
 * When `stat_partials` is not NULL ([kd_lm_head_stat_tiles(V_local)][n_rows][2] floats) the epilogue also leaves each
 * row's (max, sum 2^((x - max) log2e / T)) over every 128-column tile of the ROUNDED logits; kd_lm_head_stats_merge folds
 * them into words 7, 8 of the rows' kd_vshard_stats blocks, and kd_vshard_stats may then be called with student == NULL
 * (teacher half only) — the student logits are not read again for their statistics.
 * KD_ERR_UNSUPPORTED: fp32 operands, or bases / pitches / V_local that are not multiples of 16 bytes (use a library GEMM).
 * ---------------------------------------------------------------------------------------------------- */
int kd_lm_head_stat_tiles(int64_t V_local);
int kd_lm_head_fwd(const void* hidden, int64_t ld_h, const void* weight, int64_t ld_w, void* logits, int64_t ld_c,
                   float* stat_partials /* may be NULL */, int64_t n_rows, int64_t V_local, int64_t H, int dtype, float T,
                   void* stream);
int kd_lm_head_stats_merge(const float* stat_partials, int64_t n_tiles, float* stats, int64_t n_rows, float T, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * The same tcgen05 kernel as a general product, for the two BACKWARD GEMMs of the lm_head (the reference gets them from autograd
 * through the HF lm_head, P1:228-240) and for the learned projections of BASELINE configs[2] / [3]:
 *   C[M, N] = A[M, K] . B[N, K]^T        bf16 / fp16 operands, fp32 accumulation; C in the operand type or fp32 (c_is_f32)
 * a_mn_major / b_mn_major = 1: the operand is stored TRANSPOSED (A as [K, M] with pitch ld_a, B as [K, N] with pitch ld_b) and is read
 * through an MN-major shared-memory descriptor — nothing is transposed in HBM:
 *   dH[n, H]  = g[n, V] . W[V, H]          -> A = g (K-major), B = W with b_mn_major = 1, c_is_f32 = 1 (summed over vocabulary shards in fp32)
 *   dW[V, H]  = g[n, V]^T . hidden[n, H]   -> A = g with a_mn_major = 1, B = hidden with b_mn_major = 1
 * KD_ERR_UNSUPPORTED: fp32 operands, a_mn_major without b_mn_major, pointers / pitches that are not multiples of 16 bytes.
 * ---------------------------------------------------------------------------------------------------- */
int kd_gemm(const void* A, int64_t ld_a, int a_mn_major, const void* B, int64_t ld_b, int b_mn_major, void* C, int64_t ld_c,
            int c_is_f32, int64_t M, int64_t N, int64_t K, int dtype, void* stream);
/* kd_gemm on at most `max_ctas` CTAs (0 = one per SM): the persistent kernel otherwise holds every SM until it ends, and a
 * collective started beside it (the all-reduce of the partial d hidden under the d weight product, sharded.py) cannot get one. */
int kd_gemm_ex(const void* A, int64_t ld_a, int a_mn_major, const void* B, int64_t ld_b, int b_mn_major, void* C, int64_t ld_c,
            int c_is_f32, int64_t M, int64_t N, int64_t K, int dtype, int max_ctas, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Projected feature MSE (BASELINE configs[3]; mse_loss EV:160-165 applied to `input @ proj`, the intended use at EV:304 / P1:375
 * for a student whose width differs from the teacher's): the projection GEMM with the MSE as its epilogue — the [M, N] projected
 * tensor never reaches HBM.
 *   loss = mean( (round(input . proj) .float() - target.float())^2 )        (the product rounded to the operand type, as a library GEMM hands it on)
 *   grad_projected[M, N] = upstream * 2 (projected - target) / (M N)        in the operand type; the caller forms
 *   d input = grad_projected . proj^T (kd_gemm, both K-major) and d proj = input^T . grad_projected (kd_gemm, both MN-major).
 * input [M, K], proj [K, N], target [M, N] row-major; bf16 / fp16.
 * ---------------------------------------------------------------------------------------------------- */
int kd_proj_mse_workspace_bytes(int64_t M, int64_t N, size_t* bytes);
int kd_proj_mse_fwd_bwd(const void* input, int64_t ld_x, const void* proj, int64_t ld_w, const void* target, int64_t ld_t,
                        void* grad_projected, int64_t ld_g, float* loss_out, void* workspace, size_t workspace_bytes, int64_t M,
                        int64_t N, int64_t K, int dtype, const float* upstream, void* stream);

/* grad *= (*factor) in place, skipped on device when *factor == 1 (kept for callers that accept double rounding). */
int kd_scale_rows(void* grad, int64_t ld_g, int64_t n_rows, int64_t V, int dtype, const float* factor, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * InfoNCE — replaces the pooling at P1:243-248 plus contrastive_loss (P1:393-416; calls P1:345, FB:162).
 *   pooled[m,:] = normalize(mean_p x[m,p,:])   (tokens == 1 treats x as already [M, D])
 *   loss = mean_m( logsumexp_j(us[m].ut[j]/temp) - us[m].ut[m]/temp )
 * ---------------------------------------------------------------------------------------------------- */
/* Token-level calls (M >= 1024, M and D multiples of 8 — e.g. the 46 656 x 1152 un-pooled SigLIP tokens; contrastive_loss accepts
 * any [M, D]) run `student @ teacher.T` (P1:410) on the tcgen05 kernel: fp16 unit vectors, the rows' online log-sum-exp and the
 * diagonal in the GEMM epilogue, softmax - I from a recomputation, G . u_t as a third product; the M x M similarity matrix is
 * never stored (the fp16 gradient matrix is, in the workspace: 2 M^2 bytes).  KD_FLAG_FORCE_STREAMING selects the exact fp32
 * CUDA-core row kernel instead. */
int kd_info_nce_workspace_bytes(int64_t M, int64_t tokens, int64_t D, size_t* bytes);
int kd_info_nce_fwd_bwd(const void* student, const void* teacher, void* grad_student, float* loss_out,
                        void* workspace, size_t workspace_bytes, int64_t M, int64_t tokens, int64_t D,
                        int dtype, float temperature, const float* upstream, unsigned flags, void* stream);
/* backward re-emission from the fp32 pooled gradient kept in `workspace`: grad = (*ratio) * d(mean)/tokens broadcast
 * over tokens (ratio = grad_output / the upstream folded in at forward time), skipped when *skip_if_one == 1 (may be
 * NULL = never skip: a repeated backward that writes a fresh buffer). */
int kd_info_nce_regrad(const void* workspace, size_t workspace_bytes, void* grad_student, int64_t M, int64_t tokens,
                       int64_t D, int dtype, const float* ratio, const float* skip_if_one, void* stream);

/* Pooled contrastive loss through a LEARNED PROJECTION of the student features (BASELINE configs[2]: SigLIP 1152 -> 896 / 3584).
 * Not in the reference, whose two towers are both 1152 wide: the oracle is contrastive_loss(x_s @ proj, x_t) (P1:393-416) after the
 * pooling of P1:243-248.  Mean-pooling is linear, so the projection is applied to the pooled vectors:
 *   ps = mean_p student[m, p, :]  (fp32)      y = ps . proj  [M, D_t]      loss = contrastive_loss(y, mean_p teacher)
 *   grad_proj[D_s, D_t] = upstream * ps^T . dy   (in proj's dtype)     grad_student[m, p, :] = upstream * (dy . proj^T)[m, :] / tokens_s
 * One call = balanced pooling of both tensors (one launch), three fp32 CUDA-core products (M <= 4096 rows: too few for a
 * tensor-core tile; larger M returns KD_ERR_UNSUPPORTED — project with kd_gemm and call kd_info_nce_fwd_bwd), the row kernel
 * and the store-only broadcast.  student [M, tokens_s, D_s] and teacher [M, tokens_t, D_t] share `dtype` (tokens == 1: already
 * pooled); proj [D_s, D_t] row-major in `dtype_proj`.  grad_student / grad_proj may be NULL (not wanted). */
int kd_info_nce_proj_workspace_bytes(int64_t M, int64_t D_s, int64_t D_t, size_t* bytes);
int kd_info_nce_proj_fwd_bwd(const void* student, int64_t tokens_s, const void* teacher, int64_t tokens_t, const void* proj,
                             void* grad_student, void* grad_proj, float* loss_out, void* workspace, size_t workspace_bytes,
                             int64_t M, int64_t D_s, int64_t D_t, int dtype, int dtype_proj, float temperature, const float* upstream,
                             unsigned flags, void* stream);
/* backward re-emission from the fp32 operands kept in `workspace`: both gradients times (*ratio), rounded once; skipped when
 * *skip_if_one == 1 (NULL = never skip: a repeated backward that writes fresh buffers). */
int kd_info_nce_proj_regrad(const void* workspace, size_t workspace_bytes, void* grad_student, void* grad_proj,
                            int64_t M, int64_t tokens_s, int64_t D_s, int64_t D_t, int dtype, int dtype_proj, const float* ratio,
                            const float* skip_if_one, void* stream);

/* The pooling pass on its own — `.mean(dim=1)` over the tokens of the hooked [M, tokens, D] features (P1:243-246) in fp32 — for
 * callers that put a learned projection between pooling and the loss (BASELINE configs[2]: 1152 -> 896 / 3584), and its
 * backward: dx[m, p, :] = (*scale) * d_mean[m, :] for every token p (d_mean = d_pooled / tokens is the caller's; scale NULL = 1). */
int kd_feature_pool_workspace_bytes(int64_t M, int64_t tokens, int64_t D, size_t* bytes);
int kd_feature_pool_fwd(const void* x, float* pooled, void* workspace, size_t workspace_bytes, int64_t M, int64_t tokens,
                        int64_t D, int dtype, void* stream);
int kd_feature_pool_bwd(const float* d_mean, void* dx, int64_t M, int64_t tokens, int64_t D, int dtype, const float* scale,
                        void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Feature MSE — replaces mse_loss (EV:160-165): mean((input.float() - target.float())^2) and
 * grad_input = upstream * 2 (a - b) / numel, written in input's dtype.
 * ---------------------------------------------------------------------------------------------------- */
int kd_feature_mse_workspace_bytes(int64_t numel, size_t* bytes);
int kd_feature_mse_fwd_bwd(const void* input, const void* target, void* grad_input, float* loss_out,
                           void* workspace, size_t workspace_bytes, int64_t numel, int dtype,
                           const float* upstream, unsigned flags, void* stream);
/* backward re-emission (see kd_logit_kd_regrad): grad = (*upstream) * 2 (a - b) / numel unless *skip_if_one == 1 */
int kd_feature_mse_regrad(const void* input, const void* target, void* grad_input, int64_t numel, int dtype,
                          const float* upstream, const float* skip_if_one, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KD_B200_H_ */
