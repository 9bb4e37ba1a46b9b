// feature_mse.cu — feature-distillation MSE, replaces mse_loss (EV:160-165 of the reference):
//   loss = mean((input.float() - target.float())^2),  grad_input = upstream * 2 (a - b) / numel.
// One streaming pass: 2 reads + 1 write per element (6 B/element in bf16), HBM-bound.
#include "kd_common.cuh"

namespace kd {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8;
#ifndef KD_MSE_BATCH
#define KD_MSE_BATCH 8       // tools/knob_sweep.py at 65 536 x 3584 bf16: 2 -> 232.8 us, 4 -> 233, 6 -> 224.9, 8 -> 223.7
#endif
constexpr int kBatch = KD_MSE_BATCH;        // 16-byte vector pairs in flight per thread

template <typename T>
__global__ void __launch_bounds__(kThreads)
mse_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ g, int64_t numel, float gscale,
           const float* upstream, bool want_grad, bool vec, double* __restrict__ partials) {
    constexpr int P = Elem<T>::kPack;
    __shared__ double sh[kThreads / 32];
    const float gs = gscale * (upstream ? *upstream : 1.f);
    float acc = 0.f;          // per-thread fp32 partial over <= a few thousand terms, then double
    double dacc = 0.0;
    const int64_t nvec = vec ? numel / P : 0;
    const uint4* av = reinterpret_cast<const uint4*>(a);
    const uint4* bv = reinterpret_cast<const uint4*>(b);
    uint4* gv = reinterpret_cast<uint4*>(g);
    int since = 0;
    // kBatch vector pairs per thread are fetched before any is consumed (the volatile streaming store would otherwise fence
    // every iteration's loads behind the previous iteration's store: two loads in flight per thread)
    for (int64_t v0 = blockIdx.x * (int64_t)(kThreads * kBatch) + threadIdx.x; v0 < nvec; v0 += (int64_t)gridDim.x * kThreads * kBatch) {
        uint4 xa[kBatch], xb[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int64_t v = v0 + (int64_t)u * kThreads;
            xa[u] = make_uint4(0u, 0u, 0u, 0u);
            xb[u] = xa[u];
            if (v < nvec) { xa[u] = ldg_stream(av + v); xb[u] = ldg_stream(bv + v); }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int64_t v = v0 + (int64_t)u * kThreads;
            float x[P], y[P], o[P];
            Elem<T>::unpack(xa[u], x);
            Elem<T>::unpack(xb[u], y);
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const float d = x[i] - y[i];            // (out-of-range slots hold 0 - 0)
                acc = fmaf(d, d, acc);
                o[i] = gs * d;
            }
            if (want_grad && v < nvec) stg_stream(gv + v, Elem<T>::pack(o));
        }
        if (++since == 16) { dacc += acc; acc = 0.f; since = 0; }
    }
    for (int64_t c = nvec * P + blockIdx.x * (int64_t)kThreads + threadIdx.x; c < numel; c += (int64_t)gridDim.x * kThreads) {
        const float d = Elem<T>::to_float(a[c]) - Elem<T>::to_float(b[c]);
        acc = fmaf(d, d, acc);
        if (want_grad) g[c] = Elem<T>::from_float(gs * d);
    }
    dacc += acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = dacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < kThreads / 32; ++i) t += sh[i];
        partials[blockIdx.x] = t;
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
mse_regrad_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ g, int64_t numel, float gscale,
                  const float* upstream, const float* skip_if_one, bool vec) {
    if (skip_if_one && *skip_if_one == 1.f) return;
    constexpr int P = Elem<T>::kPack;
    const float gs = gscale * (upstream ? *upstream : 1.f);
    const int64_t nvec = vec ? numel / P : 0;
    const uint4* av = reinterpret_cast<const uint4*>(a);
    const uint4* bv = reinterpret_cast<const uint4*>(b);
    uint4* gv = reinterpret_cast<uint4*>(g);
    for (int64_t v0 = blockIdx.x * (int64_t)(kThreads * kBatch) + threadIdx.x; v0 < nvec; v0 += (int64_t)gridDim.x * kThreads * kBatch) {
        uint4 xa[kBatch], xb[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int64_t v = v0 + (int64_t)u * kThreads;
            if (v < nvec) { xa[u] = ldg_stream(av + v); xb[u] = ldg_stream(bv + v); }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int64_t v = v0 + (int64_t)u * kThreads;
            if (v < nvec) {
                float x[P], y[P], o[P];
                Elem<T>::unpack(xa[u], x);
                Elem<T>::unpack(xb[u], y);
#pragma unroll
                for (int i = 0; i < P; ++i) o[i] = gs * (x[i] - y[i]);
                stg_stream(gv + v, Elem<T>::pack(o));
            }
        }
    }
    for (int64_t c = nvec * P + blockIdx.x * (int64_t)kThreads + threadIdx.x; c < numel; c += (int64_t)gridDim.x * kThreads)
        g[c] = Elem<T>::from_float(gs * (Elem<T>::to_float(a[c]) - Elem<T>::to_float(b[c])));
}

template <typename T>
int launch_regrad(const void* a, const void* b, void* g, int64_t numel, const float* upstream, const float* skip, cudaStream_t st) {
    constexpr int P = Elem<T>::kPack;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(g)) & 15u) == 0;
    int64_t blocks = (numel / P + kThreads * kBatch - 1) / (kThreads * kBatch);
    blocks = blocks < 1 ? 1 : (blocks > kMaxBlocks ? kMaxBlocks : blocks);
    mse_regrad_kernel<T><<<(int)blocks, kThreads, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b), static_cast<T*>(g), numel,
                                                           static_cast<float>(2.0 / (double)numel), upstream, skip, vec);
    KD_CHECK_LAUNCH("mse_regrad_kernel");
    return KD_OK;
}

__global__ void mse_finalize_kernel(const double* __restrict__ partials, int n, double inv_numel, float* __restrict__ loss) {
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) a += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (threadIdx.x == 0) *loss = static_cast<float>(a * inv_numel);
}

template <typename T>
int launch(const void* a, const void* b, void* g, float* loss, double* partials, int64_t numel, const float* upstream,
           unsigned flags, cudaStream_t st) {
    constexpr int P = Elem<T>::kPack;
    const bool want_grad = !(flags & KD_FLAG_NO_GRAD);
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                       (want_grad ? reinterpret_cast<uintptr_t>(g) : 0)) & 15u) == 0;
    int64_t blocks = (numel / P + kThreads * kBatch - 1) / (kThreads * kBatch);
    if (blocks < 1) blocks = 1;
    if (blocks > kMaxBlocks) blocks = kMaxBlocks;
    mse_kernel<T><<<(int)blocks, kThreads, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b), static_cast<T*>(g),
                                                    numel, static_cast<float>(2.0 / (double)numel), upstream, want_grad, vec,
                                                    partials);
    KD_CHECK_LAUNCH("mse_kernel");
    mse_finalize_kernel<<<1, 32, 0, st>>>(partials, (int)blocks, 1.0 / (double)numel, loss);
    KD_CHECK_LAUNCH("mse_finalize_kernel");
    return KD_OK;
}

}  // namespace
}  // namespace kd

extern "C" {

int kd_feature_mse_workspace_bytes(int64_t numel, size_t* bytes) {
    KD_CHECK_ARG(bytes != nullptr && numel >= 0, "kd_feature_mse_workspace_bytes: bad arguments");
    *bytes = sizeof(double) * kd::kMaxBlocks;
    return KD_OK;
}

int kd_feature_mse_fwd_bwd(const void* input, const void* target, void* grad_input, float* loss_out, void* workspace,
                           size_t workspace_bytes, int64_t numel, int dtype, const float* upstream, unsigned flags,
                           void* stream) {
    KD_CHECK_ARG(input && target && loss_out && workspace, "kd_feature_mse_fwd_bwd: null pointer");
    KD_CHECK_ARG((flags & KD_FLAG_NO_GRAD) || grad_input, "kd_feature_mse_fwd_bwd: grad_input is null");
    KD_CHECK_ARG(numel > 0, "kd_feature_mse_fwd_bwd: numel must be positive (mean of an empty tensor is undefined)");
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32, "unknown dtype %d", dtype);
    if (workspace_bytes < sizeof(double) * kd::kMaxBlocks) {
        kd::set_error("kd_feature_mse_fwd_bwd: workspace %zu < %zu bytes", workspace_bytes, sizeof(double) * kd::kMaxBlocks);
        return KD_ERR_WORKSPACE_TOO_SMALL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* partials = static_cast<double*>(workspace);
    switch (dtype) {
        case KD_DTYPE_BF16: return kd::launch<__nv_bfloat16>(input, target, grad_input, loss_out, partials, numel, upstream, flags, st);
        case KD_DTYPE_F16: return kd::launch<__half>(input, target, grad_input, loss_out, partials, numel, upstream, flags, st);
        default: return kd::launch<float>(input, target, grad_input, loss_out, partials, numel, upstream, flags, st);
    }
}

int kd_feature_mse_regrad(const void* input, const void* target, void* grad_input, int64_t numel, int dtype,
                          const float* upstream, const float* skip_if_one, void* stream) {
    KD_CHECK_ARG(input && target && grad_input, "kd_feature_mse_regrad: null pointer");
    KD_CHECK_ARG(numel > 0, "kd_feature_mse_regrad: numel must be positive");
    KD_CHECK_ARG(dtype >= KD_DTYPE_BF16 && dtype <= KD_DTYPE_F32, "unknown dtype %d", dtype);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case KD_DTYPE_BF16: return kd::launch_regrad<__nv_bfloat16>(input, target, grad_input, numel, upstream, skip_if_one, st);
        case KD_DTYPE_F16: return kd::launch_regrad<__half>(input, target, grad_input, numel, upstream, skip_if_one, st);
        default: return kd::launch_regrad<float>(input, target, grad_input, numel, upstream, skip_if_one, st);
    }
}

}  // extern "C"
