// info_nce.cu — vision-level contrastive loss: replaces the pooling at P1:243-248 and contrastive_loss
// (P1:393-416; copies LG:307-330, FB:288-311, EV:327-350) of the reference.
//
//   u_s[m] = normalize(mean_p x_s[m,p,:]), u_t likewise (eps 1e-12; normalising twice == once)
//   L = u_s u_t^T / temp ; loss = mean_m( logsumexp_j L[m,j] - L[m,m] )
//   G = (softmax(L) - I)/M ; du = G u_t / temp ; dmean = (du - (du.u_s) u_s)/|mean| ; dx[m,p,:] = dmean[m,:]/P
//
// Four launches: balanced pooling of both tensors (HBM-bound, the only pass that reads the [M,P,D] tensors) + its finishing
// kernel (mean, F.normalize), one CTA per similarity row for logits -> log-softmax -> gradient row (the last CTA also writes the
// loss), and the store-only broadcast of dx.  With a learned projection of the student features (BASELINE configs[2]): three
// few-row fp32 products around the row kernel (kd_info_nce_proj_fwd_bwd, below).
//
// Token-level M (>= 1024 rows, M and D multiples of 8): the similarity matrix runs on the TENSOR CORES and never reaches HBM —
// kd_gemm.cu's tcgen05 kernel with (1) an epilogue that folds the rows' online (max, sum-exp) and picks up the diagonal,
// (2) a recomputation whose epilogue writes softmax - I as fp16, (3) G . u_t with an fp32 result; the unit vectors are fp16
// operands (|x| <= 1 keeps 11 mantissa bits), everything else (pooling, norms, the projection off u_s) stays fp32.
#include "kd_common.cuh"

namespace kd {
namespace {

constexpr int kPoolThreads = 512;
constexpr int kRowThreads = 1024;   // 32 warps: two similarity columns per warp at M = 64, one gradient column per thread at D ~ 1k
constexpr float kNormEps = 1e-12f;   // F.normalize default

// ---- pooling, split over tokens: CTA (split, image|tensor) sums its token chunk into an fp32 partial ---------------
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
pool_partial_kernel(const T* __restrict__ xs, const T* __restrict__ xt, float* __restrict__ partials, int M, int P, int D,
                    int nsplit, bool vec) {
    extern __shared__ float sm[];        // [groups][D] partial sums
    constexpr int E = Elem<T>::kPack;
    const int mt = blockIdx.x / nsplit, split = blockIdx.x % nsplit;      // x only: 2 M can exceed the 65 535 limit of y
    const int m = mt % M;
    const T* x = (mt >= M ? xt : xs) + (int64_t)m * P * D;
    const int p0 = (int)(((int64_t)P * split) / nsplit), p1 = (int)(((int64_t)P * (split + 1)) / nsplit);

    const int ncol = vec ? D / E : D;                       // column slots (vector or scalar)
    const int lanes = ncol < kPoolThreads ? ncol : kPoolThreads;
    const int groups = kPoolThreads / lanes;                // row groups working on disjoint token subsets
    const int col0 = threadIdx.x % lanes, grp = threadIdx.x / lanes;
    for (int i = threadIdx.x; i < groups * D; i += kPoolThreads) sm[i] = 0.f;
    __syncthreads();
    if (grp < groups) {
        for (int c = col0; c < ncol; c += lanes) {
            if (vec) {
                float acc[E];
#pragma unroll
                for (int i = 0; i < E; ++i) acc[i] = 0.f;
#pragma unroll 8
                for (int p = p0 + grp; p < p1; p += groups) {
                    float v[E];
                    Elem<T>::unpack(ldg_stream(reinterpret_cast<const uint4*>(x + (int64_t)p * D) + c), v);
#pragma unroll
                    for (int i = 0; i < E; ++i) acc[i] += v[i];
                }
#pragma unroll
                for (int i = 0; i < E; ++i) sm[grp * D + c * E + i] = acc[i];
            } else {
                float acc = 0.f;
                for (int p = p0 + grp; p < p1; p += groups) acc += Elem<T>::to_float(x[(int64_t)p * D + c]);
                sm[grp * D + c] = acc;
            }
        }
    }
    __syncthreads();
    float* out = partials + ((int64_t)split * 2 * M + mt) * D;
    for (int d = threadIdx.x; d < D; d += kPoolThreads) {
        float a = 0.f;
        for (int gI = 0; gI < groups; ++gI) a += sm[gI * D + d];
        out[d] = a;
    }
}

// ---- finish: add the partials, take the mean, L2-normalise (F.normalize eps) -------------------------------------
__global__ void __launch_bounds__(256)
pool_finish_kernel(const float* __restrict__ partials, float* __restrict__ us, float* __restrict__ ut, float* __restrict__ norm_s,
                   int M, int P, int D, int nsplit, __half* __restrict__ us16 = nullptr, __half* __restrict__ ut16 = nullptr,
                   bool normalize = true) {
    extern __shared__ float sm[];        // [D] pooled vector
    __shared__ float red[8];
    const int mt = blockIdx.x, m = mt % M;
    const bool is_teacher = mt >= M;
    float* u = (is_teacher ? ut : us) + (int64_t)m * D;
    const float invP = 1.f / (float)P;
    float ss = 0.f;
    for (int d = threadIdx.x; d < D; d += 256) {
        float a = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) a += partials[((int64_t)sp * 2 * M + mt) * D + d];
        a *= invP;
        ss = fmaf(a, a, ss);
        sm[d] = a;
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float nrm = sqrtf(tot);
    const float inv = normalize ? 1.f / fmaxf(nrm, kNormEps) : 1.f;      // kd_feature_pool: the plain mean
    for (int d = threadIdx.x; d < D; d += 256) u[d] = sm[d] * inv;
    if (us16) {          // tensor-core similarity: the unit vectors as fp16 operands (|x| <= 1: 11 mantissa bits)
        __half* h = (is_teacher ? ut16 : us16) + (int64_t)m * D;
        for (int d = threadIdx.x; d < D; d += 256) h[d] = __float2half_rn(sm[d] * inv);
    }
    if (!is_teacher && threadIdx.x == 0 && norm_s) norm_s[m] = nrm;
}

// ---- one similarity row per CTA: logits, log-softmax, gradient w.r.t. the pooled student vector -----
__global__ void __launch_bounds__(kRowThreads, 1)
similarity_row_kernel(const float* __restrict__ us, const float* __restrict__ ut, const float* __restrict__ norm_s,
                      float* __restrict__ dmean, float* __restrict__ rowloss, int M, int D, float inv_temp,
                      float inv_tokens /* 1/P */, const float* upstream, bool want_grad,
                      int raw_parts = 0 /* > 0: `us` holds that many [M, D] partials of the UN-normalised student rows */,
                      float* __restrict__ dmeanT = nullptr /* optional [D, M] copy of the gradient rows */,
                      unsigned* __restrict__ done_counter = nullptr /* zeroed by the pooling's finishing kernel ... */,
                      float* __restrict__ loss_out = nullptr /* ... the LAST CTA to get here writes mean(rowloss) */) {
    extern __shared__ float sm[];   // [D] u_s row | [M] logits / G row
    float* su = sm;
    float* sl = sm + D;
    __shared__ float red[kRowThreads / 32];
    __shared__ float bc[2];
    const int i = blockIdx.x;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    float nrm_raw = 0.f;
    if (raw_parts > 0) {
        // projected student row: add the split-reduction partials in order, then F.normalize it here
        float ss = 0.f;
        for (int d = threadIdx.x; d < D; d += kRowThreads) {
            float a = 0.f;
            for (int z = 0; z < raw_parts; ++z) a += us[((int64_t)z * M + i) * D + d];
            su[d] = a;
            ss = fmaf(a, a, ss);
        }
        ss = warp_sum(ss);
        if (l == 0) red[w] = ss;
        __syncthreads();
        float tot = 0.f;
        for (int k = 0; k < kRowThreads / 32; ++k) tot += red[k];
        nrm_raw = sqrtf(tot);
        const float inv = 1.f / fmaxf(nrm_raw, kNormEps);
        for (int d = threadIdx.x; d < D; d += kRowThreads) su[d] *= inv;
    } else {
        for (int d = threadIdx.x; d < D; d += kRowThreads) su[d] = us[(int64_t)i * D + d];
    }
    __syncthreads();
    const bool vec4 = (D % 4 == 0);
    for (int j = w; j < M; j += kRowThreads / 32) {
        const float* t = ut + (int64_t)j * D;
        float acc = 0.f;
        if (vec4) {
            const float4* t4 = reinterpret_cast<const float4*>(t);
            const float4* s4 = reinterpret_cast<const float4*>(su);
#pragma unroll 12
            for (int d = l; d < D / 4; d += 32) {      // (12 loads in flight per lane: the whole row at D = 1152)
                const float4 a = s4[d], b = __ldg(t4 + d);
                acc = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
            }
        } else {
            for (int d = l; d < D; d += 32) acc = fmaf(su[d], t[d], acc);
        }
        acc = warp_sum(acc);
        if (l == 0) sl[j] = acc * inv_temp;
    }
    __syncthreads();
    // log-sum-exp over the row
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < M; j += kRowThreads) mx = fmaxf(mx, sl[j]);
    mx = warp_max(mx);
    if (l == 0) red[w] = mx;
    __syncthreads();
    mx = red[0];
    for (int k = 1; k < kRowThreads / 32; ++k) mx = fmaxf(mx, red[k]);
    __syncthreads();
    // sum over the OFF-diagonal entries only: 1 - p_ii = r/(exp(L_ii - mx) + r) and lse - L_ii = log1p-style, so a
    // well-separated diagonal (p_ii -> 1) does not lose the gradient to cancellation
    const float lii = sl[i];
    float se = 0.f;
    for (int j = threadIdx.x; j < M; j += kRowThreads) se += (j == i) ? 0.f : __expf(sl[j] - mx);
    se = warp_sum(se);
    if (l == 0) red[w] = se;
    __syncthreads();
    float r = 0.f;
    for (int k = 0; k < kRowThreads / 32; ++k) r += red[k];
    const float ed = __expf(lii - mx);                 // diagonal term relative to the max
    const float denom = ed + r;
    // loss_i = lse - L_ii = log(denom) - (L_ii - mx);  when the diagonal is the max (ed == 1) use log1p(r)
    const float loss_i = (lii == mx) ? log1pf(r) : (logf(denom) - (lii - mx));
    if (threadIdx.x == 0) rowloss[i] = loss_i;
    if (done_counter && w == 0) {
        // the last CTA adds the rows' terms (fixed lane assignment + fixed shuffle tree: deterministic) — saves the finalize launch
        unsigned prev = 0;
        if (l == 0) {
            __threadfence();
            prev = atomicAdd(done_counter, 1u);
        }
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev == (unsigned)M - 1u) {
            __threadfence();
            double a = 0.0;
            for (int j = l; j < M; j += 32) a += (double)__ldcg(rowloss + j);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (l == 0) *loss_out = static_cast<float>(a / (double)M);
        }
    }
    if (!want_grad) return;
    __syncthreads();
    // G row: (softmax - onehot)/M, folded with 1/temp; diagonal = -(sum of the others)/denom
    const float gsc = inv_temp / ((float)M * denom);
    for (int j = threadIdx.x; j < M; j += kRowThreads)
        sl[j] = (j == i) ? -r * gsc : __expf(sl[j] - mx) * gsc;
    __syncthreads();
    // du[d] = sum_j G[j] * ut[j,d]  (coalesced over d; ut is L2 resident)
    const float nrm = raw_parts > 0 ? nrm_raw : norm_s[i];
    float dots = 0.f;
    if (D % 2 == 0 && D <= 2 * kRowThreads) {
        // one pass: a thread owns two adjacent columns and keeps 16 rows of u_t in flight (D = 1152 on 1024 threads used to take
        // two passes of eight 8-deep batches: this column sum, not the dot products, was most of the kernel's 20 us)
        const int d = 2 * threadIdx.x;
        if (d < D) {
            float2 acc = make_float2(0.f, 0.f);
            const float2* t2 = reinterpret_cast<const float2*>(ut + d);
#pragma unroll 16
            for (int j = 0; j < M; ++j) {
                const float2 v = __ldg(t2 + (int64_t)j * (D / 2));
                acc.x = fmaf(sl[j], v.x, acc.x);
                acc.y = fmaf(sl[j], v.y, acc.y);
            }
            dots = fmaf(acc.x, su[d], acc.y * su[d + 1]);
            *reinterpret_cast<float2*>(dmean + (int64_t)i * D + d) = acc;
        }
    } else {
        for (int d = threadIdx.x; d < D; d += kRowThreads) {
            float acc = 0.f;
#pragma unroll 8
            for (int j = 0; j < M; ++j) acc = fmaf(sl[j], __ldg(ut + (int64_t)j * D + d), acc);
            dots = fmaf(acc, su[d], dots);
            dmean[(int64_t)i * D + d] = acc;
        }
    }
    dots = warp_sum(dots);
    if (l == 0) red[w] = dots;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < kRowThreads / 32; ++k) t += red[k];
        bc[0] = t;
    }
    __syncthreads();
    const float dot = (nrm > kNormEps) ? bc[0] : 0.f;      // below eps normalize() is a plain scaling
    const float sc = inv_tokens * (upstream ? *upstream : 1.f) / fmaxf(nrm, kNormEps);
    for (int d = threadIdx.x; d < D; d += kRowThreads) {
        const float du = dmean[(int64_t)i * D + d];
        const float v = (du - dot * su[d]) * sc;
        dmean[(int64_t)i * D + d] = v;
        if (dmeanT) dmeanT[(int64_t)d * M + i] = v;
    }
}

__global__ void __launch_bounds__(1024) nce_finalize_kernel(const float* __restrict__ rowloss, int M, float* __restrict__ loss) {
    __shared__ double sh[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) a += (double)rowloss[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += sh[k];      // fixed order: deterministic
        *loss = static_cast<float>(tot / (double)M);
    }
}

// ---- dx[m,p,:] = dmean[m,:] -----------------------------------------------------------------------
// A thread owns one 16-byte column slot of the row: it converts its source values once and then only stores, token after
// token (the previous version recomputed idx % ncol and re-read the source for every store: 1.8 TB/s).
template <typename T>
__global__ void __launch_bounds__(256)
broadcast_rows_kernel(const float* __restrict__ dmean, T* __restrict__ dx, int P, int D, bool vec, const float* ratio,
                      const float* skip_if_one, int nsrc = 1 /* the source is the sum of this many partials ... */,
                      int64_t src_stride = 0 /* ... this many floats apart */, float host_scale = 1.f) {
    constexpr int E = Elem<T>::kPack;
    if (skip_if_one && *skip_if_one == 1.f) return;      // backward re-emission: grad_output is what forward assumed
    const float f = (ratio ? *ratio : 1.f) * host_scale;
    const int m = blockIdx.x;                              // rows on x (no 65 535 limit), token chunks on y
    const float* src = dmean + (int64_t)m * D;
    T* dst = dx + (int64_t)m * P * D;
    if (vec) {
        const int ncol = D / E;
        const int lanes = ncol < 256 ? ncol : 256;          // column slots handled side by side
        const int slots = 256 / lanes;                       // token rows handled side by side
        const int c0 = threadIdx.x % lanes, slot = threadIdx.x / lanes;
        if (slot >= slots) return;
        for (int c = c0; c < ncol; c += lanes) {
            float v[E];
#pragma unroll
            for (int i = 0; i < E; ++i) {
                float a = src[c * E + i];
                for (int z = 1; z < nsrc; ++z) a += src[z * src_stride + c * E + i];
                v[i] = a * f;
            }
            const uint4 packed = Elem<T>::pack(v);
            uint4* col = reinterpret_cast<uint4*>(dst) + c;
#pragma unroll 4
            for (int p = blockIdx.y * slots + slot; p < P; p += gridDim.y * slots) stg_stream(col + (int64_t)p * ncol, packed);
        }
    } else {
        const int64_t total = (int64_t)P * D;
        for (int64_t idx = blockIdx.y * 256ll + threadIdx.x; idx < total; idx += (int64_t)gridDim.y * 256) {
            float a = src[idx % D];
            for (int z = 1; z < nsrc; ++z) a += src[z * src_stride + idx % D];
            dst[idx] = Elem<T>::from_float(a * f);
        }
    }
}


// ---- tensor-core similarity (token-level M) ------------------------------------------------------------------------
// P == 1 (features are already [M, D]): one warp normalises student row i AND teacher row i (fp32 math; fp32 + fp16 copies out)
// and takes their exact fp32 dot product — the diagonal similarity the loss subtracts.
template <typename T>
__global__ void __launch_bounds__(256)
nce_normalize_pair_kernel(const T* __restrict__ xs, const T* __restrict__ xt, float* __restrict__ us, float* __restrict__ ut,
                          __half* __restrict__ us16, __half* __restrict__ ut16, float* __restrict__ norm_s,
                          float* __restrict__ diag_exact, int M, int D) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (i >= M) return;
    const T* a = xs + (int64_t)i * D;
    const T* b = xt + (int64_t)i * D;
    float ss = 0.f, tt = 0.f, st = 0.f;
    for (int d = l; d < D; d += 32) {
        const float x = Elem<T>::to_float(a[d]), y = Elem<T>::to_float(b[d]);
        ss = fmaf(x, x, ss);
        tt = fmaf(y, y, tt);
        st = fmaf(x, y, st);
    }
    ss = warp_sum(ss);
    tt = warp_sum(tt);
    st = warp_sum(st);
    const float ns = sqrtf(ss), nt = sqrtf(tt);
    const float is = 1.f / fmaxf(ns, kNormEps), it = 1.f / fmaxf(nt, kNormEps);
    for (int d = l; d < D; d += 32) {
        const float x = Elem<T>::to_float(a[d]) * is, y = Elem<T>::to_float(b[d]) * it;
        us[(int64_t)i * D + d] = x;
        ut[(int64_t)i * D + d] = y;
        us16[(int64_t)i * D + d] = __float2half_rn(x);
        ut16[(int64_t)i * D + d] = __float2half_rn(y);
    }
    if (l == 0) {
        norm_s[i] = ns;
        diag_exact[i] = st * is * it;
    }
}
// pooled features (P > 1): the exact diagonal from the fp32 unit vectors, one warp per row
__global__ void __launch_bounds__(256)
nce_diag_kernel(const float* __restrict__ us, const float* __restrict__ ut, float* __restrict__ diag_exact, int M, int D) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (i >= M) return;
    float st = 0.f;
    for (int d = l; d < D; d += 32) st = fmaf(us[(int64_t)i * D + d], ut[(int64_t)i * D + d], st);
    st = warp_sum(st);
    if (l == 0) diag_exact[i] = st;
}

// Rows' log-sum-exp and loss terms from the GEMM epilogue's partials.  L[i, j] = acc[i, j] / temp (log2 units: acc * k2).
//   lse2[i] = log2 sum_j 2^(acc k2)  — from the tensor-core values only, so that the gradient epilogue's softmax is consistent;
//   loss_i  = log1p( sum_{j != i} 2^((acc_ij - d_i) k2) ),  d_i = the EXACT fp32 diagonal similarity.
// The tensor core accumulates in fp32 with truncation: a sum of 1152 same-sign terms (the diagonal of related features) comes
// out ~4e-6 low, which shifts lse - L_ii by 5e-5 relative when the off-diagonal entries (mixed signs, unbiased) are measured
// against it.  Measuring them against the exact diagonal removes the bias.  The tiles are merged in double and the part of the
// sum away from the maximum is kept apart: with a dominant diagonal the sum is 1 + r, r ~ 1e-2 ... 1e-6, and r's digits matter.
__global__ void __launch_bounds__(256)
nce_tc_rows_kernel(const float2* __restrict__ partials, int n_tiles, const float* __restrict__ diag_tc,
                   const float* __restrict__ diag_exact, int M, float k2, float* __restrict__ lse2, float* __restrict__ rowloss) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= M) return;
    float mx = -INFINITY;
    for (int t = 0; t < n_tiles; ++t) mx = fmaxf(mx, partials[(size_t)t * M + i].x);
    double main_s = 0.0, rest = 0.0;          // tiles sitting at the maximum / all others, relative to mx
    for (int t = 0; t < n_tiles; ++t) {
        const float2 p = partials[(size_t)t * M + i];
        if (p.x == -INFINITY) continue;
        if (p.x == mx) main_s += (double)p.y;
        else rest += (double)(p.y * exp2f((p.x - mx) * k2));
    }
    lse2[i] = (float)((double)mx * (double)k2 + log2(main_s + rest));
    // off-diagonal sum relative to mx: take the tensor core's own diagonal term out of the tile it sits in
    const double dterm = exp2((double)(diag_tc[i] - mx) * (double)k2);
    const double off = (main_s - dterm) + rest;
    rowloss[i] = (float)log1p(fmax(off, 0.0) * exp2((double)(mx - diag_exact[i]) * (double)k2));
}

// dmean[i, :] = (du - (du . u_s) u_s) * upstream / (P * max(|mean|, eps)),  du = dU[i, :] * (1 / (2^14 * temp * M))
__global__ void __launch_bounds__(256)
nce_tc_finish_kernel(const float* __restrict__ dU, const float* __restrict__ us, const float* __restrict__ norm_s,
                     float* __restrict__ dmean, int D, float du_scale, float inv_tokens, const float* upstream) {
    __shared__ float red[8];
    const int i = blockIdx.x;
    const float* du = dU + (int64_t)i * D;
    const float* u = us + (int64_t)i * D;
    float dots = 0.f;
    for (int d = threadIdx.x; d < D; d += 256) dots = fmaf(du[d] * du_scale, u[d], dots);
    dots = warp_sum(dots);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dots;
    __syncthreads();
    float tot = 0.f;
    for (int k = 0; k < 8; ++k) tot += red[k];
    const float nrm = norm_s[i];
    const float dot = (nrm > kNormEps) ? tot : 0.f;
    const float sc = inv_tokens * (upstream ? *upstream : 1.f) / fmaxf(nrm, kNormEps);
    for (int d = threadIdx.x; d < D; d += 256) dmean[(int64_t)i * D + d] = (du[d] * du_scale - dot * u[d]) * sc;
}

// ---- balanced pooling (the CUDA-core forms: pooled M = 64 and friends) ----------------------------------------------
// A tensor's token rows ([M, P, D] contiguous = M * P rows of D elements) are cut into G contiguous, equally long ranges,
// one per CTA: every SM streams the same number of bytes whatever M is (the (image, split) grid above leaves 80 of its 592
// CTA slots empty at M = 64), and two tensors of different width share one launch with CTA counts in proportion to their
// bytes.  G >= M, so a range touches at most two images: the CTA writes one fp32 partial per image piece and
// pool_ranges_finish_kernel adds an image's pieces in CTA order (deterministic).  The loads go out in explicit batches of
// kPoolBatch 16-byte vectors per thread before any is consumed: an unrolled loop with an exit test per iteration compiles to
// one load in flight per thread (SASS of pool_partial_kernel), 27 KB per SM — half of what the HBM latency asks for.
#ifndef KD_POOL_BATCH
#define KD_POOL_BATCH 8
#endif
constexpr int kPoolBatch = KD_POOL_BATCH;
struct PoolTensor { const void* x; int M, P, D, G, cta0; };
struct PoolPlan { PoolTensor t[2]; int n, dmax; };

template <typename T, bool VEC>
__global__ void __launch_bounds__(kPoolThreads, 2)
pool_ranges_kernel(const PoolPlan plan, float* __restrict__ partials) {
    extern __shared__ float sm[];        // [groups][D] partial sums
    constexpr int E = VEC ? Elem<T>::kPack : 1;
    const PoolTensor t = (plan.n > 1 && (int)blockIdx.x >= plan.t[1].cta0) ? plan.t[1] : plan.t[0];      // (a select, not an indexed copy)
    const int b = (int)blockIdx.x - t.cta0;
    const int D = t.D, P = t.P;
    const int64_t R = (int64_t)t.M * P;
    const int64_t r0 = R * b / t.G, r1 = R * (b + 1) / t.G;
    const int64_t m0 = r0 / P;
    const int ncol = D / E;
    const int lanes = ncol < kPoolThreads ? ncol : kPoolThreads;
    const int groups = kPoolThreads / lanes;                // row groups working on disjoint token subsets
    const int c0 = threadIdx.x % lanes, grp = threadIdx.x / lanes;
    const T* x = static_cast<const T*>(t.x);
    for (int s = 0; s < 2; ++s) {
        const int64_t lo = (m0 + s) * P, hi = lo + P;
        const int64_t a = r0 > lo ? r0 : lo, e = r1 < hi ? r1 : hi;
        if (a >= e) break;                                  // (uniform) the range ends inside its first image
        if (grp < groups) {
            for (int c = c0; c < ncol; c += lanes) {
                float acc[E];
#pragma unroll
                for (int i = 0; i < E; ++i) acc[i] = 0.f;
                for (int64_t r = a + grp; r < e; r += (int64_t)groups * kPoolBatch) {
                    if constexpr (VEC) {
                        uint4 v[kPoolBatch];
#pragma unroll
                        for (int u = 0; u < kPoolBatch; ++u) {
                            const int64_t rr = r + (int64_t)u * groups;
                            v[u] = make_uint4(0u, 0u, 0u, 0u);              // all-zero bits are 0.0 in the three element types
                            if (rr < e) v[u] = ldg_stream(reinterpret_cast<const uint4*>(x + rr * D) + c);
                        }
#pragma unroll
                        for (int u = 0; u < kPoolBatch; ++u) {
                            float f[Elem<T>::kPack];
                            Elem<T>::unpack(v[u], f);
#pragma unroll
                            for (int i = 0; i < E; ++i) acc[i] += f[i];
                        }
                    } else {
                        float v[kPoolBatch];
#pragma unroll
                        for (int u = 0; u < kPoolBatch; ++u) {
                            const int64_t rr = r + (int64_t)u * groups;
                            v[u] = (rr < e) ? Elem<T>::to_float(x[rr * D + c]) : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < kPoolBatch; ++u) acc[0] += v[u];
                    }
                }
#pragma unroll
                for (int i = 0; i < E; ++i) sm[grp * D + c * E + i] = acc[i];
            }
        }
        __syncthreads();
        float* out = partials + ((int64_t)blockIdx.x * 2 + s) * plan.dmax;
        for (int d = threadIdx.x; d < D; d += kPoolThreads) {
            float v = 0.f;
            for (int gI = 0; gI < groups; ++gI) v += sm[gI * D + d];
            out[d] = v;
        }
        __syncthreads();
    }
}

// One CTA per (tensor, image): add the image's pieces in CTA order, take the mean and (bit k of normalize_mask) L2-normalise
// tensor k's rows (F.normalize eps); the student's norm goes to norm0 when that is given.  A thread owns 4 consecutive
// features and fetches its float4 of every piece (at most 6: G <= 4 M) before adding them: one round trip to L2, not one per piece.
constexpr int kFinishPieces = 8;
__global__ void __launch_bounds__(1024)
pool_ranges_finish_kernel(const PoolPlan plan, const float* __restrict__ partials, float* __restrict__ out0, float* __restrict__ out1,
                          float* __restrict__ norm0, unsigned normalize_mask, unsigned* __restrict__ zero_word) {
    __shared__ float red[32];
    __shared__ long long offs[kFinishPieces];
    if (zero_word && blockIdx.x == 0 && threadIdx.x == 0) *zero_word = 0u;      // the row kernel's "CTAs done" counter
    __shared__ int s_first, s_count;
    const int k = (plan.n > 1 && (int)blockIdx.x >= plan.t[0].M) ? 1 : 0;
    const PoolTensor t = k ? plan.t[1] : plan.t[0];
    const int m = (int)blockIdx.x - (k ? plan.t[0].M : 0);
    const int D = t.D, P = t.P, G = t.G;
    const int64_t R = (int64_t)t.M * P;
    const int64_t lo = (int64_t)m * P, hi = lo + P;
    float* const u_out = (k ? out1 : out0) + (int64_t)m * D;
    const bool vec4 = (D % 4 == 0) && (plan.dmax % 4 == 0) && (reinterpret_cast<uintptr_t>(u_out) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(partials) & 15u) == 0;
    const int ncol = vec4 ? D / 4 : D;
    if (threadIdx.x == 0) {
        int b = (int)(lo * G / R);
        b = b > 0 ? b - 1 : 0;
        while (R * (b + 1) / G <= lo) ++b;                   // first CTA whose range reaches into the image
        int n = 0;
        while (b + n < G && R * (b + n) / G < hi) ++n;
        s_first = b;
        s_count = n;
    }
    __syncthreads();
    const int first = s_first, count = s_count;
    float4 acc[4];       // up to 4 column slots per thread (D <= 16 384 with 1024 threads)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c0 = 0; c0 < count; c0 += kFinishPieces) {
        const int nb = min(kFinishPieces, count - c0);
        __syncthreads();
        if ((int)threadIdx.x < nb) {
            const int b = first + c0 + threadIdx.x;
            const int seg = ((R * b / G) / P == m) ? 0 : 1;  // the image is the first / second one of that CTA's range
            offs[threadIdx.x] = ((long long)(t.cta0 + b) * 2 + seg) * plan.dmax;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = threadIdx.x + q * blockDim.x;
            if (c >= ncol) break;
            float4 v[kFinishPieces];
#pragma unroll
            for (int bi = 0; bi < kFinishPieces; ++bi) {
                v[bi] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bi < nb) {
                    if (vec4) v[bi] = *reinterpret_cast<const float4*>(partials + offs[bi] + 4 * c);
                    else v[bi].x = partials[offs[bi] + c];
                }
            }
#pragma unroll
            for (int bi = 0; bi < kFinishPieces; ++bi) {      // fixed order: deterministic
                acc[q].x += v[bi].x; acc[q].y += v[bi].y; acc[q].z += v[bi].z; acc[q].w += v[bi].w;
            }
        }
    }
    const float invP = 1.f / (float)P;
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        acc[q].x *= invP; acc[q].y *= invP; acc[q].z *= invP; acc[q].w *= invP;
        ss += acc[q].x * acc[q].x + acc[q].y * acc[q].y + acc[q].z * acc[q].z + acc[q].w * acc[q].w;      // unused slots are 0
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    const float nrm = sqrtf(tot);
    const float inv = ((normalize_mask >> k) & 1u) ? 1.f / fmaxf(nrm, kNormEps) : 1.f;
    float* u = u_out;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = threadIdx.x + q * blockDim.x;
        if (c >= ncol) break;
        if (vec4) *reinterpret_cast<float4*>(u + 4 * c) = make_float4(acc[q].x * inv, acc[q].y * inv, acc[q].z * inv, acc[q].w * inv);
        else u[c] = acc[q].x * inv;
    }
    if (k == 0 && threadIdx.x == 0 && norm0) norm0[m] = nrm;
}

// CTA counts of the balanced pooling: two CTAs per SM shared out by bytes, at least M (a range then spans at most two images) and at
// most 4 M / one row per CTA (the partials fit the block the (image, split) form already reserves: 8 * 2 M rows of D floats)
int sm_count_cached() {
    static int dev_cached = -1, sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != dev_cached) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        dev_cached = dev;
    }
    return sms > 0 ? sms : 148;
}
void pool_plan_fill(PoolPlan& pl, int elem_bytes) {
    const int budget = 2 * sm_count_cached();
    double total = 0.0;
    for (int k = 0; k < pl.n; ++k) total += (double)pl.t[k].M * pl.t[k].P * pl.t[k].D * elem_bytes;
    int cta0 = 0;
    pl.dmax = 0;
    for (int k = 0; k < pl.n; ++k) {
        PoolTensor& t = pl.t[k];
        const int64_t R = (int64_t)t.M * t.P;
        int64_t G = (int64_t)(budget * ((double)t.M * t.P * t.D * elem_bytes / total) + 0.5);
        if (G < t.M) G = t.M;
        const int64_t cap = R < 4ll * t.M ? R : 4ll * t.M;
        if (G > cap) G = cap;
        t.G = (int)G;
        t.cta0 = cta0;
        cta0 += t.G;
        if (t.D > pl.dmax) pl.dmax = t.D;
    }
}
// upper bound for workspace queries (before the pointers are known)
size_t pool_ranges_partial_bound(int64_t M0, int64_t D0, int64_t M1, int64_t D1) {
    return (size_t)(4 * (M0 + M1)) * 2 * (size_t)(D0 > D1 ? D0 : D1);
}

template <typename T>
int launch_pool_ranges(PoolPlan& pl, float* partials, float* out0, float* out1, float* norm0, unsigned normalize_mask, cudaStream_t st,
                       unsigned* zero_word = nullptr) {
    constexpr int E = Elem<T>::kPack;
    bool vec = true;
    for (int k = 0; k < pl.n; ++k) vec = vec && (pl.t[k].D % E == 0) && (reinterpret_cast<uintptr_t>(pl.t[k].x) & 15u) == 0;
    pool_plan_fill(pl, (int)sizeof(T));
    size_t smem = 0;
    int ctas = 0, images = 0;
    for (int k = 0; k < pl.n; ++k) {
        const int ncol = vec ? pl.t[k].D / E : pl.t[k].D;
        const int lanes = ncol < kPoolThreads ? ncol : kPoolThreads;
        const size_t s = sizeof(float) * (size_t)(kPoolThreads / lanes) * pl.t[k].D;
        smem = s > smem ? s : smem;
        ctas += pl.t[k].G;
        images += pl.t[k].M;
    }
    KD_CHECK_ARG(smem <= 200 * 1024 && sizeof(float) * (size_t)pl.dmax <= 200 * 1024, "pooling: feature dimension %d too large", pl.dmax);
    if (vec) {
        auto kfn = pool_ranges_kernel<T, true>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kfn<<<ctas, kPoolThreads, smem, st>>>(pl, partials);
    } else {
        auto kfn = pool_ranges_kernel<T, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kfn<<<ctas, kPoolThreads, smem, st>>>(pl, partials);
    }
    KD_CHECK_LAUNCH("pool_ranges_kernel");
    int fcols = 0;        // column slots (float4 or scalar) of the widest tensor: one per thread, up to 4 when a CTA cannot hold them
    for (int k = 0; k < pl.n; ++k) {
        const float* o = k ? out1 : out0;
        const bool v4 = pl.t[k].D % 4 == 0 && pl.dmax % 4 == 0 && (reinterpret_cast<uintptr_t>(o) & 15u) == 0 &&
                        (reinterpret_cast<uintptr_t>(partials) & 15u) == 0;
        const int c = v4 ? pl.t[k].D / 4 : pl.t[k].D;
        fcols = c > fcols ? c : fcols;
    }
    KD_CHECK_ARG(fcols <= 4096, "pooling: feature dimension %d too large for the finishing kernel", pl.dmax);
    int fthreads = (fcols + 31) / 32 * 32;
    if (fthreads > 1024) fthreads = 1024;
    pool_ranges_finish_kernel<<<images, fthreads, 0, st>>>(pl, partials, out0, out1, norm0, normalize_mask, zero_word);
    KD_CHECK_LAUNCH("pool_ranges_finish_kernel");
    return KD_OK;
}

// ---- few-row fp32 products around the pooled contrastive loss (learned projection, BASELINE configs[2]) ---------------
// C[i, j] = sum_r A(i, r) * B(j, r) on the CUDA cores, fp32 accumulation: with M = 64 rows these products (66 MFMA each) are
// too small for a 256-row tensor-core tile, and fp32 FMAs keep the pooled features' digits without hi + lo operand splitting.
// *_RM ("reduction-major") says the operand is stored [r][i]: its 64-wide tile row is read with 128-bit shared-memory loads
// by the fast thread index; an operand stored [i][r] is read as warp-wide broadcasts by the slow index — so no operand is ever
// transposed in memory as long as one of the two is reduction-major.  64 x 64 tile, 4 x 4 outputs per thread, blockIdx.z
// splits the reduction (partials [z][I][J] for the consumer to add in order: deterministic).
constexpr int kSgTile = 64, kSgRc = 32;
template <typename TA, typename TB, bool A_RM, bool B_RM, typename TC>
__global__ void __launch_bounds__(256)
small_gemm_kernel(const TA* __restrict__ A, int64_t lda, const TB* __restrict__ B, int64_t ldb, TC* __restrict__ C, int64_t ldc,
                  int64_t split_stride, int I, int J, int Rr, int rsplit, const float* out_scale = nullptr,
                  const float* skip_if_one = nullptr) {
    static_assert(A_RM || B_RM, "one operand must be reduction-major");
    if (skip_if_one && *skip_if_one == 1.f) return;      // backward re-emission: grad_output is what forward assumed
    const float oscale = out_scale ? *out_scale : 1.f;
    constexpr int kPitch = kSgRc + 4;       // [i][r] operand: rows 36 floats apart — 16-byte aligned float4s of 4 consecutive r, conflict-free fills
    __shared__ __align__(16) float As[A_RM ? kSgRc * kSgTile : kSgTile * kPitch];
    __shared__ __align__(16) float Bs[B_RM ? kSgRc * kSgTile : kSgTile * kPitch];
    const int tid = threadIdx.x;
    const int ti = A_RM ? tid % 16 : tid / 16;       // the broadcast operand takes the slow index
    const int tj = A_RM ? tid / 16 : tid % 16;
    const int i0 = blockIdx.x * kSgTile, j0 = blockIdx.y * kSgTile;
    const int rbeg = blockIdx.z * rsplit, rend = min(Rr, rbeg + rsplit);
    float acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[q][p] = 0.f;
    // the next reduction chunk's operands are fetched into registers while the current one is multiplied (the global round trip
    // of every chunk used to sit in front of its 32 steps of arithmetic)
    constexpr int NQ = (kSgRc * kSgTile) / 256;
    float pa[NQ], pb[NQ];
    auto fetch = [&](int rc) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int idx = tid + 256 * q;
            if constexpr (A_RM) {
                const int r = idx / kSgTile, i = idx % kSgTile;
                pa[q] = (rc + r < rend && i0 + i < I) ? Elem<TA>::to_float(A[(int64_t)(rc + r) * lda + i0 + i]) : 0.f;
            } else {
                const int i = idx / kSgRc, r = idx % kSgRc;
                pa[q] = (rc + r < rend && i0 + i < I) ? Elem<TA>::to_float(A[(int64_t)(i0 + i) * lda + rc + r]) : 0.f;
            }
            if constexpr (B_RM) {
                const int r = idx / kSgTile, j = idx % kSgTile;
                pb[q] = (rc + r < rend && j0 + j < J) ? Elem<TB>::to_float(B[(int64_t)(rc + r) * ldb + j0 + j]) : 0.f;
            } else {
                const int j = idx / kSgRc, r = idx % kSgRc;
                pb[q] = (rc + r < rend && j0 + j < J) ? Elem<TB>::to_float(B[(int64_t)(j0 + j) * ldb + rc + r]) : 0.f;
            }
        }
    };
    if (rbeg < rend) fetch(rbeg);
    for (int rc = rbeg; rc < rend; rc += kSgRc) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int idx = tid + 256 * q;
            if constexpr (A_RM) As[(idx / kSgTile) * kSgTile + idx % kSgTile] = pa[q];
            else As[(idx / kSgRc) * kPitch + idx % kSgRc] = pa[q];
            if constexpr (B_RM) Bs[(idx / kSgTile) * kSgTile + idx % kSgTile] = pb[q];
            else Bs[(idx / kSgRc) * kPitch + idx % kSgRc] = pb[q];
        }
        if (rc + kSgRc < rend) fetch(rc + kSgRc);
        __syncthreads();
        // four reduction steps per iteration: eight 128-bit shared loads up front, then 64 independent FMAs
#pragma unroll 2
        for (int r = 0; r < kSgRc; r += 4) {
            float a[4][4], bb[4][4];       // [reduction step][output]
            if constexpr (A_RM) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v = *reinterpret_cast<const float4*>(&As[(r + k) * kSgTile + ti * 4]);
                    a[k][0] = v.x; a[k][1] = v.y; a[k][2] = v.z; a[k][3] = v.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4*>(&As[(ti * 4 + q) * kPitch + r]);
                    a[0][q] = v.x; a[1][q] = v.y; a[2][q] = v.z; a[3][q] = v.w;
                }
            }
            if constexpr (B_RM) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v = *reinterpret_cast<const float4*>(&Bs[(r + k) * kSgTile + tj * 4]);
                    bb[k][0] = v.x; bb[k][1] = v.y; bb[k][2] = v.z; bb[k][3] = v.w;
                }
            } else {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float4 v = *reinterpret_cast<const float4*>(&Bs[(tj * 4 + p) * kPitch + r]);
                    bb[0][p] = v.x; bb[1][p] = v.y; bb[2][p] = v.z; bb[3][p] = v.w;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int p = 0; p < 4; ++p) acc[q][p] = fmaf(a[k][q], bb[k][p], acc[q][p]);
        }
        __syncthreads();
    }
    TC* out = C + (int64_t)blockIdx.z * split_stride;
    const bool vec_st = (ldc % 4 == 0) && (split_stride % 4 == 0) && (reinterpret_cast<uintptr_t>(C) & 15u) == 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = i0 + ti * 4 + q;
        if (i < I) {
            const int j = j0 + tj * 4;
            TC* dst = out + (int64_t)i * ldc + j;
            if (vec_st && j + 3 < J) {      // the thread's four consecutive outputs of the row as one store
                if constexpr (sizeof(TC) == 4) {
                    *reinterpret_cast<float4*>(dst) = make_float4(acc[q][0] * oscale, acc[q][1] * oscale, acc[q][2] * oscale, acc[q][3] * oscale);
                } else {
                    const uint2 pk2 = make_uint2(Pair<TC>::narrow(acc[q][0] * oscale, acc[q][1] * oscale),
                                                 Pair<TC>::narrow(acc[q][2] * oscale, acc[q][3] * oscale));
                    *reinterpret_cast<uint2*>(dst) = pk2;
                }
            } else {
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (j + p < J) dst[p] = Elem<TC>::from_float(acc[q][p] * oscale);
            }
        }
    }
}

template <typename TA, typename TB, bool A_RM, bool B_RM, typename TC>
int launch_small_gemm(const TA* A, int64_t lda, const TB* B, int64_t ldb, TC* C, int64_t ldc, int I, int J, int Rr, int nsplit,
                      cudaStream_t st, const float* out_scale = nullptr, const float* skip_if_one = nullptr) {
    const int rsplit = ((Rr + nsplit - 1) / nsplit + kSgRc - 1) / kSgRc * kSgRc;
    const dim3 grid((I + kSgTile - 1) / kSgTile, (J + kSgTile - 1) / kSgTile, (Rr + rsplit - 1) / rsplit);
    small_gemm_kernel<TA, TB, A_RM, B_RM, TC><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, (int64_t)I * ldc, I, J, Rr, rsplit, out_scale,
                                                                    skip_if_one);
    KD_CHECK_LAUNCH("small_gemm_kernel");
    return KD_OK;
}
// how many ways a reduction of length Rr is split so that the grid reaches ~one CTA per SM; the consumers add the partials in order
int small_gemm_splits(int I, int J, int Rr) {
    const int tiles = ((I + kSgTile - 1) / kSgTile) * ((J + kSgTile - 1) / kSgTile);
    int want = (sm_count_cached() + tiles - 1) / tiles;
    const int most = (Rr + 127) / 128;                     // at least 128 reduction steps per CTA
    if (want > most) want = most;
    if (want > 16) want = 16;
    if (want < 1) want = 1;
    const int rsplit = ((Rr + want - 1) / want + kSgRc - 1) / kSgRc * kSgRc;
    return (Rr + rsplit - 1) / rsplit;
}

constexpr int kMaxPoolSplit = 8;
struct NceWorkspace { size_t us, ut, norm_s, dmean, rowloss, counter, partials, total; };
NceWorkspace nce_layout(int64_t M, int64_t D) {
    NceWorkspace w{};
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = (off + b + 255) / 256 * 256; return o; };
    w.us = take(sizeof(float) * M * D);
    w.ut = take(sizeof(float) * M * D);
    w.norm_s = take(sizeof(float) * M);
    w.dmean = take(sizeof(float) * M * D);
    w.rowloss = take(sizeof(float) * M);
    w.counter = take(sizeof(unsigned));
    w.partials = take(sizeof(float) * kMaxPoolSplit * 2 * M * D);
    w.total = off;
    return w;
}
// Tensor-core form: the same leading blocks (kd_info_nce_regrad finds dmean at the same offset), pooling partials for the one
// split large M gets, then the fp16 operands, the epilogue's statistics, the [M, M] gradient matrix and the fp32 product G . ut
struct NceTcWorkspace { NceWorkspace base; size_t us16, ut16, stat, diag, diag_exact, lse2, g16, du, total; };
NceTcWorkspace nce_tc_layout(int64_t M, int64_t D) {
    NceTcWorkspace w{};
    w.base = nce_layout(M, D);
    size_t off = w.base.partials;
    auto take = [&](size_t b) { size_t o = off; off = (off + b + 255) / 256 * 256; return o; };
    take(sizeof(float) * 2 * M * D);                       // pooling partials, nsplit = 1
    w.us16 = take(sizeof(__half) * M * D);
    w.ut16 = take(sizeof(__half) * M * D);
    w.stat = take(sizeof(float2) * (size_t)nce_stat_tiles(M) * M);
    w.diag = take(sizeof(float) * M);
    w.diag_exact = take(sizeof(float) * M);
    w.lse2 = take(sizeof(float) * M);
    w.g16 = take(sizeof(__half) * M * M);
    w.du = take(sizeof(float) * M * D);
    w.total = off;
    return w;
}
constexpr int kTcMinRows = 1024;        // below this the CUDA-core row kernel (exact fp32 similarities) is the better tool
bool nce_tc_applies(int64_t M, int64_t D) { return M >= kTcMinRows && M % 8 == 0 && D % 8 == 0; }

template <typename T>
int run(const void* xs, const void* xt, void* dx, float* loss, char* ws, const NceWorkspace& w, int M, int P, int D,
        float temperature, const float* upstream, unsigned flags, cudaStream_t st, const NceTcWorkspace* tc = nullptr) {
    constexpr int E = Elem<T>::kPack;
    const bool want_grad = !(flags & KD_FLAG_NO_GRAD);
    float* us = reinterpret_cast<float*>(ws + w.us);
    float* ut = reinterpret_cast<float*>(ws + w.ut);
    float* norm_s = reinterpret_cast<float*>(ws + w.norm_s);
    float* dmean = reinterpret_cast<float*>(ws + w.dmean);
    float* rowloss = reinterpret_cast<float*>(ws + w.rowloss);
    const bool vec_in = (D % E == 0) && ((reinterpret_cast<uintptr_t>(xs) | reinterpret_cast<uintptr_t>(xt)) & 15u) == 0;
    if (tc && P == 1) {
        nce_normalize_pair_kernel<T><<<(M + 7) / 8, 256, 0, st>>>(static_cast<const T*>(xs), static_cast<const T*>(xt), us, ut,
                                                                 reinterpret_cast<__half*>(ws + tc->us16), reinterpret_cast<__half*>(ws + tc->ut16),
                                                                 norm_s, reinterpret_cast<float*>(ws + tc->diag_exact), M, D);
        KD_CHECK_LAUNCH("nce_normalize_pair_kernel");
    } else if (!tc) {
        // pooled M (the 64-image case): balanced contiguous ranges, both tensors in one launch
        PoolPlan pl{};
        pl.n = 2;
        pl.t[0] = PoolTensor{xs, M, P, D, 0, 0};
        pl.t[1] = PoolTensor{xt, M, P, D, 0, 0};
        if (int rc = launch_pool_ranges<T>(pl, reinterpret_cast<float*>(ws + w.partials), us, ut, norm_s, 3u, st,
                                           reinterpret_cast<unsigned*>(ws + w.counter)))
            return rc;
    } else {
        float* partials = reinterpret_cast<float*>(ws + w.partials);
        const int ncol = vec_in ? D / E : D;
        const int lanes = ncol < kPoolThreads ? ncol : kPoolThreads;
        const size_t smem = sizeof(float) * (size_t)(kPoolThreads / lanes) * D;
        KD_CHECK_ARG(smem <= 200 * 1024, "info_nce: feature dimension %d too large for the pooling kernel", D);
        const int nsplit = 1;        // token-level M: one CTA per (tensor, row) already fills the machine
        auto kfn = pool_partial_kernel<T>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kfn<<<(unsigned)(nsplit * 2 * (int64_t)M), kPoolThreads, smem, st>>>(static_cast<const T*>(xs), static_cast<const T*>(xt), partials, M, P, D, nsplit, vec_in);
        KD_CHECK_LAUNCH("pool_partial_kernel");
        KD_CHECK_ARG(sizeof(float) * (size_t)D <= 200 * 1024, "info_nce: feature dimension %d too large", D);
        if (sizeof(float) * (size_t)D > 48 * 1024)
            cudaFuncSetAttribute(pool_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * D));
        pool_finish_kernel<<<2 * M, 256, sizeof(float) * (size_t)D, st>>>(partials, us, ut, norm_s, M, P, D, nsplit,
                                                                          reinterpret_cast<__half*>(ws + tc->us16),
                                                                          reinterpret_cast<__half*>(ws + tc->ut16));
        KD_CHECK_LAUNCH("pool_finish_kernel");
    }
    if (tc) {
        // Similarities on the tensor cores (P1:410 `torch.mm(student_features, teacher_features.T)`), never stored: one product
        // for the rows' log-sum-exp + diagonal, one recomputation whose epilogue writes softmax - I, one product with u_t.
        const float k2 = kLog2e / temperature;
        const __half* us16 = reinterpret_cast<const __half*>(ws + tc->us16);
        const __half* ut16 = reinterpret_cast<const __half*>(ws + tc->ut16);
        float* stat = reinterpret_cast<float*>(ws + tc->stat);
        float* diag = reinterpret_cast<float*>(ws + tc->diag);
        float* lse2 = reinterpret_cast<float*>(ws + tc->lse2);
        float* diag_exact = reinterpret_cast<float*>(ws + tc->diag_exact);
        if (P != 1) {
            nce_diag_kernel<<<(M + 7) / 8, 256, 0, st>>>(us, ut, diag_exact, M, D);
            KD_CHECK_LAUNCH("nce_diag_kernel");
        }
        if (int rc = nce_gemm_stats(us16, ut16, M, D, k2, stat, diag, st)) return rc;
        nce_tc_rows_kernel<<<(M + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float2*>(stat), nce_stat_tiles(M), diag, diag_exact, M, k2, lse2, rowloss);
        KD_CHECK_LAUNCH("nce_tc_rows_kernel");
        if (want_grad) {
            __half* g16 = reinterpret_cast<__half*>(ws + tc->g16);
            float* du = reinterpret_cast<float*>(ws + tc->du);
            if (int rc = nce_gemm_grad(us16, ut16, M, D, k2, lse2, g16, st)) return rc;
            if (int rc = nce_gemm_du(g16, ut16, du, M, D, st)) return rc;
            nce_tc_finish_kernel<<<M, 256, 0, st>>>(du, us, norm_s, dmean, D, 1.f / (nce_grad_scale() * temperature * (float)M),
                                                    1.f / (float)P, upstream);
            KD_CHECK_LAUNCH("nce_tc_finish_kernel");
        }
    } else {
        const size_t smem = sizeof(float) * ((size_t)D + M);
        KD_CHECK_ARG(smem <= 220 * 1024, "info_nce: M + D = %d floats exceed the row kernel's shared memory; the tensor-core form needs M and D to be multiples of 8", M + D);
        if (smem > 48 * 1024) cudaFuncSetAttribute(similarity_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        // (the pooling's finishing kernel zeroed the counter: the last row CTA writes the loss)
        similarity_row_kernel<<<M, kRowThreads, smem, st>>>(us, ut, norm_s, dmean, rowloss, M, D, 1.f / temperature,
                                                            1.f / (float)P, upstream, want_grad, 0, nullptr,
                                                            reinterpret_cast<unsigned*>(ws + w.counter), loss);
        KD_CHECK_LAUNCH("similarity_row_kernel");
    }
    if (tc) {
        nce_finalize_kernel<<<1, M >= 1024 ? 1024 : 32, 0, st>>>(rowloss, M, loss);
        KD_CHECK_LAUNCH("nce_finalize_kernel");
    }
    if (want_grad) {
        const bool vec_out = (D % E == 0) && (reinterpret_cast<uintptr_t>(dx) & 15u) == 0;
        int bx = vec_out ? (P + 15) / 16 : (int)(((int64_t)P * D + 255) / 256);      // >= 16 tokens per CTA and row
        if (bx > 32) bx = 32;
        if (bx < 1) bx = 1;
        broadcast_rows_kernel<T><<<dim3(M, bx), 256, 0, st>>>(dmean, static_cast<T*>(dx), P, D, vec_out, nullptr, nullptr);
        KD_CHECK_LAUNCH("broadcast_rows_kernel");
    }
    return KD_OK;
}

template <typename T>
int run_regrad(const float* dmean, void* dx, int M, int P, int D, const float* ratio, const float* skip, cudaStream_t st) {
    constexpr int E = Elem<T>::kPack;
    const bool vec_out = (D % E == 0) && (reinterpret_cast<uintptr_t>(dx) & 15u) == 0;
    int bx = vec_out ? (P + 15) / 16 : (int)(((int64_t)P * D + 255) / 256);
    bx = bx > 32 ? 32 : (bx < 1 ? 1 : bx);
    broadcast_rows_kernel<T><<<dim3(M, bx), 256, 0, st>>>(dmean, static_cast<T*>(dx), P, D, vec_out, ratio, skip);
    KD_CHECK_LAUNCH("broadcast_rows_kernel");
    return KD_OK;
}


// ---- pooled contrastive loss through a learned projection (BASELINE configs[2]: 1152 -> 896 / 3584) ------------------
//   ps = mean_p x_s [M, Ds] (fp32)    y = ps . W [M, Dt]    u_s = normalize(y)    u_t = normalize(mean_p x_t)
//   loss = contrastive_loss(y, mean_p x_t)  (P1:393-416 on the projected student vectors — pooling is linear, so projecting the
//   pooled vectors equals pooling the projected tokens);  dy from the row kernel,  dW = ps^T . dy,  d ps = dy . W^T,
//   dx_s[m, p, :] = d ps[m, :] / P.  Eight launches, every product in fp32 on the CUDA cores.
struct NceProjWorkspace { size_t ps, ut, ypart, dy, dyT, rowloss, counter, dpspart, pool, total; };
constexpr int kProjMaxSplit = 16;
NceProjWorkspace nce_proj_layout(int64_t M, int64_t Ds, int64_t Dt) {
    NceProjWorkspace w{};
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = (off + b + 255) / 256 * 256; return o; };
    w.ps = take(sizeof(float) * M * Ds);
    w.ut = take(sizeof(float) * M * Dt);
    w.dy = take(sizeof(float) * M * Dt);
    w.dyT = take(sizeof(float) * M * Dt);
    w.rowloss = take(sizeof(float) * M);
    w.counter = take(sizeof(unsigned));
    w.dpspart = take(sizeof(float) * kProjMaxSplit * M * Ds);
    w.ypart = take(sizeof(float) * kProjMaxSplit * M * Dt);
    w.pool = take(sizeof(float) * pool_ranges_partial_bound(M, Ds, M, Dt));
    w.total = off;
    return w;
}
constexpr int64_t kProjMaxRows = 4096;      // beyond this the similarity rows belong on the tensor cores (kd_info_nce_fwd_bwd)

template <typename T, typename TW>
int run_proj(const void* xs, int Ps, const void* xt, int Pt, const TW* W, void* dx, TW* dW, float* loss, char* ws,
             const NceProjWorkspace& w, int M, int Ds, int Dt, float temperature, const float* upstream, bool want_dx, bool want_dw,
             cudaStream_t st) {
    constexpr int E = Elem<T>::kPack;
    float* ps = reinterpret_cast<float*>(ws + w.ps);
    float* ut = reinterpret_cast<float*>(ws + w.ut);
    float* ypart = reinterpret_cast<float*>(ws + w.ypart);
    float* dy = reinterpret_cast<float*>(ws + w.dy);
    float* dyT = reinterpret_cast<float*>(ws + w.dyT);
    float* rowloss = reinterpret_cast<float*>(ws + w.rowloss);
    float* dpspart = reinterpret_cast<float*>(ws + w.dpspart);
    PoolPlan pl{};
    pl.n = 2;
    pl.t[0] = PoolTensor{xs, M, Ps, Ds, 0, 0};
    pl.t[1] = PoolTensor{xt, M, Pt, Dt, 0, 0};
    // student: the plain mean (the projection comes first); teacher: mean + F.normalize
    if (int rc = launch_pool_ranges<T>(pl, reinterpret_cast<float*>(ws + w.pool), ps, ut, nullptr, 2u, st,
                                       reinterpret_cast<unsigned*>(ws + w.counter)))
        return rc;
    const int sy = small_gemm_splits(M, Dt, Ds);
    if (int rc = launch_small_gemm<float, TW, false, true, float>(ps, Ds, W, Dt, ypart, Dt, M, Dt, Ds, sy, st)) return rc;
    const bool want_grad = want_dx || want_dw;
    const size_t smem = sizeof(float) * ((size_t)Dt + M);
    KD_CHECK_ARG(smem <= 220 * 1024, "info_nce (projected): M + D_t = %d floats exceed the row kernel's shared memory", M + Dt);
    if (smem > 48 * 1024) cudaFuncSetAttribute(similarity_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    similarity_row_kernel<<<M, kRowThreads, smem, st>>>(ypart, ut, nullptr, dy, rowloss, M, Dt, 1.f / temperature, 1.f, upstream, want_grad,
                                                        sy, dyT, reinterpret_cast<unsigned*>(ws + w.counter), loss);
    KD_CHECK_LAUNCH("similarity_row_kernel");
    if (want_dw) {       // dW[k, n] = sum_m ps[m, k] dy[m, n]: both operands reduction-major as stored
        if (int rc = launch_small_gemm<float, float, true, true, TW>(ps, Ds, dy, Dt, dW, Dt, Ds, Dt, M, 1, st)) return rc;
    }
    if (want_dx) {       // d ps[m, k] = sum_n dy[m, n] W[k, n]: the transposed copy of dy is the reduction-major operand
        const int sd = small_gemm_splits(M, Ds, Dt);
        if (int rc = launch_small_gemm<float, TW, true, false, float>(dyT, M, W, Dt, dpspart, Ds, M, Ds, Dt, sd, st)) return rc;
        const bool vec_out = (Ds % E == 0) && (reinterpret_cast<uintptr_t>(dx) & 15u) == 0;
        int bx = vec_out ? (Ps + 15) / 16 : (int)(((int64_t)Ps * Ds + 255) / 256);
        bx = bx > 32 ? 32 : (bx < 1 ? 1 : bx);
        if (sd > 1 && bx > 8) bx = 8;       // every CTA adds the sd partials of its row first: fewer, longer CTAs (32 -> 18 us)
        broadcast_rows_kernel<T><<<dim3(M, bx), 256, 0, st>>>(dpspart, static_cast<T*>(dx), Ps, Ds, vec_out, nullptr, nullptr, sd,
                                                              (int64_t)M * Ds, 1.f / (float)Ps);
        KD_CHECK_LAUNCH("broadcast_rows_kernel");
    }
    return KD_OK;
}

template <typename T, typename TW>
int run_proj_regrad(const char* ws, const NceProjWorkspace& w, void* dx, TW* dW, int M, int Ps, int Ds, int Dt,
                    const float* ratio, const float* skip, cudaStream_t st) {
    constexpr int E = Elem<T>::kPack;
    if (dx) {
        const float* dpspart = reinterpret_cast<const float*>(ws + w.dpspart);
        const int sd = small_gemm_splits(M, Ds, Dt);
        const bool vec_out = (Ds % E == 0) && (reinterpret_cast<uintptr_t>(dx) & 15u) == 0;
        int bx = vec_out ? (Ps + 15) / 16 : (int)(((int64_t)Ps * Ds + 255) / 256);
        bx = bx > 32 ? 32 : (bx < 1 ? 1 : bx);
        if (sd > 1 && bx > 8) bx = 8;
        broadcast_rows_kernel<T><<<dim3(M, bx), 256, 0, st>>>(dpspart, static_cast<T*>(dx), Ps, Ds, vec_out, ratio, skip, sd,
                                                              (int64_t)M * Ds, 1.f / (float)Ps);
        KD_CHECK_LAUNCH("broadcast_rows_kernel");
    }
    if (dW) {      // rebuilt from the fp32 operands with the ratio applied before the one rounding to proj's dtype
        const float* ps = reinterpret_cast<const float*>(ws + w.ps);
        const float* dy = reinterpret_cast<const float*>(ws + w.dy);
        if (int rc = launch_small_gemm<float, float, true, true, TW>(ps, Ds, dy, Dt, dW, Dt, Ds, Dt, M, 1, st, ratio, skip)) return rc;
    }
    return KD_OK;
}

// (feature dtype, projection dtype) -> instantiation
template <typename F>
int dispatch2(int dtype, int dtype_w, F&& f) {
    auto inner = [&](auto t) -> int {
        switch (dtype_w) {
            case KD_DTYPE_BF16: return f(t, __nv_bfloat16{});
            case KD_DTYPE_F16: return f(t, __half{});
            default: return f(t, float{});
        }
    };
    switch (dtype) {
        case KD_DTYPE_BF16: return inner(__nv_bfloat16{});
        case KD_DTYPE_F16: return inner(__half{});
        default: return inner(float{});
    }
}

}  // namespace
}  // namespace kd

extern "C" {

int kd_info_nce_workspace_bytes(int64_t M, int64_t tokens, int64_t D, size_t* bytes) {
    KD_CHECK_ARG(bytes != nullptr && M >= 1 && tokens >= 1 && D >= 1, "kd_info_nce_workspace_bytes: bad arguments");
    *bytes = kd::nce_layout(M, D).total;
    if (kd::nce_tc_applies(M, D)) {        // either form may run (KD_FLAG_FORCE_STREAMING picks the row kernel): room for both
        const size_t tc = kd::nce_tc_layout(M, D).total;
        if (tc > *bytes) *bytes = tc;
    }
    return KD_OK;
}

int kd_info_nce_fwd_bwd(const void* student, const void* teacher, void* grad_student, float* loss_out, void* workspace,
                        size_t workspace_bytes, int64_t M, int64_t tokens, int64_t D, int dtype, float temperature,
                        const float* upstream, unsigned flags, void* stream) {
    KD_CHECK_ARG(student && teacher && loss_out && workspace, "kd_info_nce_fwd_bwd: null pointer");
    KD_CHECK_ARG((flags & KD_FLAG_NO_GRAD) || grad_student, "kd_info_nce_fwd_bwd: grad_student is null");
    KD_CHECK_ARG(M >= 1 && tokens >= 1 && D >= 1 && M < (1 << 20) && D < (1 << 20) && tokens < (1 << 24), "kd_info_nce_fwd_bwd: bad shape");
    KD_CHECK_ARG(temperature > 0.f, "temperature must be positive");
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32, "unknown dtype %d", dtype);
    const kd::NceWorkspace w = kd::nce_layout(M, D);
    // token-level M: similarities on the tensor cores; small M (the pooled 64 x 64 case): one CTA per row on the CUDA cores
    const bool use_tc = kd::nce_tc_applies(M, D) && !(flags & KD_FLAG_FORCE_STREAMING);
    const kd::NceTcWorkspace wtc = use_tc ? kd::nce_tc_layout(M, D) : kd::NceTcWorkspace{};
    const size_t need = use_tc ? wtc.total : w.total;
    if (workspace_bytes < need) {
        kd::set_error("kd_info_nce_fwd_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    const kd::NceTcWorkspace* tc = use_tc ? &wtc : nullptr;
    switch (dtype) {
        case KD_DTYPE_BF16: return kd::run<__nv_bfloat16>(student, teacher, grad_student, loss_out, ws, w, (int)M, (int)tokens, (int)D, temperature, upstream, flags, st, tc);
        case KD_DTYPE_F16: return kd::run<__half>(student, teacher, grad_student, loss_out, ws, w, (int)M, (int)tokens, (int)D, temperature, upstream, flags, st, tc);
        default: return kd::run<float>(student, teacher, grad_student, loss_out, ws, w, (int)M, (int)tokens, (int)D, temperature, upstream, flags, st, tc);
    }
}

int kd_info_nce_regrad(const void* workspace, size_t workspace_bytes, void* grad_student, int64_t M, int64_t tokens,
                       int64_t D, int dtype, const float* ratio, const float* skip_if_one, void* stream) {
    KD_CHECK_ARG(workspace && grad_student && ratio, "kd_info_nce_regrad: null pointer");
    KD_CHECK_ARG(M >= 1 && tokens >= 1 && D >= 1, "kd_info_nce_regrad: bad shape");
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32, "unknown dtype %d", dtype);
    const kd::NceWorkspace w = kd::nce_layout(M, D);
    if (workspace_bytes < w.partials) {          // only the leading blocks (pooled gradient) are read; both forms share them
        kd::set_error("kd_info_nce_regrad: workspace %zu < %zu bytes", workspace_bytes, w.partials);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    const float* dmean = reinterpret_cast<const float*>(static_cast<const char*>(workspace) + w.dmean);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case KD_DTYPE_BF16: return kd::run_regrad<__nv_bfloat16>(dmean, grad_student, (int)M, (int)tokens, (int)D, ratio, skip_if_one, st);
        case KD_DTYPE_F16: return kd::run_regrad<__half>(dmean, grad_student, (int)M, (int)tokens, (int)D, ratio, skip_if_one, st);
        default: return kd::run_regrad<float>(dmean, grad_student, (int)M, (int)tokens, (int)D, ratio, skip_if_one, st);
    }
}

/* mean over the token dimension in fp32 (the `.mean(dim=1)` of P1:243-246) — the pooling pass of kd_info_nce_fwd_bwd on its own,
 * for callers that put a learned projection between pooling and the loss (BASELINE configs[2]) */
int kd_feature_pool_workspace_bytes(int64_t M, int64_t tokens, int64_t D, size_t* bytes) {
    KD_CHECK_ARG(bytes != nullptr && M >= 1 && tokens >= 1 && D >= 1, "kd_feature_pool_workspace_bytes: bad arguments");
    *bytes = sizeof(float) * (size_t)kd::kMaxPoolSplit * 2 * M * D + 256;
    return KD_OK;
}

int kd_feature_pool_fwd(const void* x, float* pooled, void* workspace, size_t workspace_bytes, int64_t M, int64_t tokens, int64_t D,
                        int dtype, void* stream) {
    using namespace kd;
    KD_CHECK_ARG(x && pooled && workspace, "kd_feature_pool_fwd: null pointer");
    KD_CHECK_ARG(M >= 1 && tokens >= 1 && D >= 1 && M < (1 << 20) && D < (1 << 20) && tokens < (1 << 24), "kd_feature_pool_fwd: bad shape");
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32, "unknown dtype %d", dtype);
    size_t need = 0;
    kd_feature_pool_workspace_bytes(M, tokens, D, &need);
    if (workspace_bytes < need) {
        set_error("kd_feature_pool_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* partials = static_cast<float*>(workspace);
    const int Mi = (int)M, P = (int)tokens, Di = (int)D;
    auto go = [&](auto tag) -> int {
        using T = decltype(tag);
        PoolPlan pl{};
        pl.n = 1;
        pl.t[0] = PoolTensor{x, Mi, P, Di, 0, 0};
        return launch_pool_ranges<T>(pl, partials, pooled, nullptr, nullptr, 0u, st);      // the plain mean
    };
    switch (dtype) {
        case KD_DTYPE_BF16: return go(__nv_bfloat16{});
        case KD_DTYPE_F16: return go(__half{});
        default: return go(float{});
    }
}

/* backward of the pooling: dx[m, p, :] = (*scale) * d_mean[m, :] for every token p, in x's dtype; d_mean = d_pooled / tokens is
 * the caller's (an [M, D] fp32 tensor); scale: device scalar or NULL = 1 */
int kd_feature_pool_bwd(const float* d_pooled, void* dx, int64_t M, int64_t tokens, int64_t D, int dtype, const float* scale, void* stream) {
    using namespace kd;
    KD_CHECK_ARG(d_pooled && dx, "kd_feature_pool_bwd: null pointer");
    KD_CHECK_ARG(M >= 1 && tokens >= 1 && D >= 1 && M < (1 << 20) && D < (1 << 20) && tokens < (1 << 24), "kd_feature_pool_bwd: bad shape");
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32, "unknown dtype %d", dtype);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case KD_DTYPE_BF16: return run_regrad<__nv_bfloat16>(d_pooled, dx, (int)M, (int)tokens, (int)D, scale, nullptr, st);
        case KD_DTYPE_F16: return run_regrad<__half>(d_pooled, dx, (int)M, (int)tokens, (int)D, scale, nullptr, st);
        default: return run_regrad<float>(d_pooled, dx, (int)M, (int)tokens, (int)D, scale, nullptr, st);
    }
}

/* Pooled contrastive loss through a learned projection of the student features — see include/kd_b200.h */
int kd_info_nce_proj_workspace_bytes(int64_t M, int64_t D_s, int64_t D_t, size_t* bytes) {
    KD_CHECK_ARG(bytes != nullptr && M >= 1 && D_s >= 1 && D_t >= 1, "kd_info_nce_proj_workspace_bytes: bad arguments");
    *bytes = kd::nce_proj_layout(M, D_s, D_t).total;
    return KD_OK;
}

int kd_info_nce_proj_fwd_bwd(const void* student, int64_t tokens_s, const void* teacher, int64_t tokens_t, const void* proj,
                             void* grad_student, void* grad_proj, float* loss_out, void* workspace, size_t workspace_bytes,
                             int64_t M, int64_t D_s, int64_t D_t, int dtype, int dtype_proj, float temperature, const float* upstream,
                             unsigned flags, void* stream) {
    using namespace kd;
    KD_CHECK_ARG(student && teacher && proj && loss_out && workspace, "kd_info_nce_proj_fwd_bwd: null pointer");
    KD_CHECK_ARG(M >= 1 && tokens_s >= 1 && tokens_t >= 1 && D_s >= 1 && D_t >= 1 && D_s < (1 << 20) && D_t < (1 << 20) &&
                 tokens_s < (1 << 24) && tokens_t < (1 << 24), "kd_info_nce_proj_fwd_bwd: bad shape");
    if (M > kProjMaxRows) {
        set_error("kd_info_nce_proj_fwd_bwd: M = %lld rows: project with kd_gemm and call kd_info_nce_fwd_bwd (tensor-core similarity)", (long long)M);
        return KD_ERR_UNSUPPORTED;
    }
    KD_CHECK_ARG(temperature > 0.f, "temperature must be positive");
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32 && dtype_proj >= KD_DTYPE_BF16 && dtype_proj <= KD_DTYPE_F32, "unknown dtype");
    const NceProjWorkspace w = nce_proj_layout(M, D_s, D_t);
    if (workspace_bytes < w.total) {
        set_error("kd_info_nce_proj_fwd_bwd: workspace %zu < %zu bytes", workspace_bytes, w.total);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    const bool no_grad = (flags & KD_FLAG_NO_GRAD) != 0;
    const bool want_dx = !no_grad && grad_student, want_dw = !no_grad && grad_proj;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    return dispatch2(dtype, dtype_proj, [&](auto t, auto tw) -> int {
        using T = decltype(t);
        using TW = decltype(tw);
        return run_proj<T, TW>(student, (int)tokens_s, teacher, (int)tokens_t, static_cast<const TW*>(proj), grad_student,
                               static_cast<TW*>(grad_proj), loss_out, ws, w, (int)M, (int)D_s, (int)D_t, temperature, upstream,
                               want_dx, want_dw, st);
    });
}

int kd_info_nce_proj_regrad(const void* workspace, size_t workspace_bytes, void* grad_student, void* grad_proj,
                            int64_t M, int64_t tokens_s, int64_t D_s, int64_t D_t, int dtype, int dtype_proj, const float* ratio,
                            const float* skip_if_one, void* stream) {
    using namespace kd;
    KD_CHECK_ARG(workspace && ratio && (grad_student || grad_proj), "kd_info_nce_proj_regrad: null pointer");
    KD_CHECK_ARG(M >= 1 && M <= kProjMaxRows && tokens_s >= 1 && D_s >= 1 && D_t >= 1, "kd_info_nce_proj_regrad: bad shape");
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32 && dtype_proj >= KD_DTYPE_BF16 && dtype_proj <= KD_DTYPE_F32, "unknown dtype");
    const NceProjWorkspace w = nce_proj_layout(M, D_s, D_t);
    if (workspace_bytes < w.total) {
        set_error("kd_info_nce_proj_regrad: workspace %zu < %zu bytes", workspace_bytes, w.total);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const char* ws = static_cast<const char*>(workspace);
    return dispatch2(dtype, dtype_proj, [&](auto t, auto tw) -> int {
        using T = decltype(t);
        using TW = decltype(tw);
        return run_proj_regrad<T, TW>(ws, w, grad_student, static_cast<TW*>(grad_proj), (int)M, (int)tokens_s,
                                      (int)D_s, (int)D_t, ratio, skip_if_one, st);
    });
}


}  // extern "C"
