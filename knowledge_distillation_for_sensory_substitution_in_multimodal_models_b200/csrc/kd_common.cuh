// kd_common.cuh — device helpers shared by the KD-loss kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kd_b200.h"

namespace kd {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kWarp = 32;

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
// kd_gemm.cu: the tensor-core products of contrastive_loss (fp16 unit vectors), driven by info_nce.cu
int nce_stat_tiles(int64_t M);
int nce_gemm_stats(const void* us16, const void* ut16, int64_t M, int64_t D, float k2, float* partials, float* diag, cudaStream_t st);
int nce_gemm_grad(const void* us16, const void* ut16, int64_t M, int64_t D, float k2, const float* lse2, void* G16, cudaStream_t st);
int nce_gemm_du(const void* G16, const void* ut16, float* dU, int64_t M, int64_t D, cudaStream_t st);
float nce_grad_scale();

#define KD_CHECK_ARG(cond, ...)                      \
    do {                                             \
        if (!(cond)) {                               \
            ::kd::set_error(__VA_ARGS__);            \
            return KD_ERR_INVALID_ARGUMENT;          \
        }                                            \
    } while (0)

#define KD_CHECK_LAUNCH(what)                                         \
    do {                                                              \
        cudaError_t e__ = cudaGetLastError();                         \
        if (e__ != cudaSuccess) return ::kd::cuda_fail(e__, what);    \
    } while (0)

// ---------------------------------------------------------------------------------------------
// fast math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2(float x) {   // MUFU.EX2, flush-to-zero; ex2(-inf) = 0
#ifdef KD_ABLATE_MUFU      // timing experiment only (wrong results): take the XU pipe out of the picture
    return x * 0.5f;
#endif
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------------
// 16-byte vector access for the three element types: a "pack" is what one 128-bit access moves
// ---------------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kPack = 8;
    static __device__ __forceinline__ void unpack(const uint4& v, float (&x)[8]) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x[2 * i] = __uint_as_float(w[i] << 16);
            x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ uint4 pack(const float (&x)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Elem<__half> {
    static constexpr int kPack = 8;
    static __device__ __forceinline__ void unpack(const uint4& v, float (&x)[8]) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            x[2 * i] = f.x;
            x[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ uint4 pack(const float (&x)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
};
template <> struct Elem<float> {
    static constexpr int kPack = 4;
    static __device__ __forceinline__ void unpack(const uint4& v, float (&x)[4]) {
        x[0] = __uint_as_float(v.x);
        x[1] = __uint_as_float(v.y);
        x[2] = __uint_as_float(v.z);
        x[3] = __uint_as_float(v.w);
    }
    static __device__ __forceinline__ uint4 pack(const float (&x)[4]) {
        return make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3]));
    }
    static __device__ __forceinline__ float to_float(float v) { return v; }
    static __device__ __forceinline__ float from_float(float v) { return v; }
};

// streaming 128-bit global accesses
__device__ __forceinline__ uint4 ldg_stream(const void* p) {          // read-once data: do not pollute L1
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_keep(const void* p) {            // data re-read shortly (L2 resident)
    uint4 v;
    asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {  // write-once data
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_cs(void* p, const uint4& v) {   // write-once data, cache-streaming (evict-first in L2)
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// running (max, sum of 2^((x-max)*k2)) pair in the raw-logit domain
struct MaxSum {
    float m, s;
    __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
    __device__ __forceinline__ void merge(float om, float os, float k2) {
        float nm = fmaxf(m, om);
        float a = (m == -INFINITY) ? 0.f : s * ex2((m - nm) * k2);
        float b = (om == -INFINITY) ? 0.f : os * ex2((om - nm) * k2);
        m = nm;
        s = a + b;
    }
};
__device__ __forceinline__ void warp_merge(MaxSum& v, float k2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, v.m, o);
        float os = __shfl_xor_sync(0xffffffffu, v.s, o);
        v.merge(om, os, k2);
    }
}

// best / runner-up with "largest value, lowest index" ordering (our topk tie rule)
struct Top2 {
    float a1, a2;
    int i1, i2;
    __device__ __forceinline__ void init() { a1 = a2 = -INFINITY; i1 = i2 = 0x7fffffff; }
    static __device__ __forceinline__ bool before(float va, int ia, float vb, int ib) {
        return va > vb || (va == vb && ia < ib);
    }
    __device__ __forceinline__ void push(float v, int i) {
        if (before(v, i, a1, i1)) { a2 = a1; i2 = i1; a1 = v; i1 = i; }
        else if (before(v, i, a2, i2)) { a2 = v; i2 = i; }
    }
    __device__ __forceinline__ void merge(const Top2& o) { push(o.a1, o.i1); push(o.a2, o.i2); }
};
__device__ __forceinline__ void warp_merge(Top2& t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Top2 p;
        p.a1 = __shfl_xor_sync(0xffffffffu, t.a1, o);
        p.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
        p.a2 = __shfl_xor_sync(0xffffffffu, t.a2, o);
        p.i2 = __shfl_xor_sync(0xffffffffu, t.i2, o);
        t.merge(p);
    }
}

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2) --------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// ---- exp2 on the FMA pipe (Cody-Waite + polynomial), two values at a time ---------------------------------
// The fused emit kernels are bound by the XU pipe (MUFU.EX2: 16 lanes/clk/SM) inside their sweeps while issue slots are
// free, so a fraction of the exponentials is evaluated with packed FMAs instead:  t = x + 1.5*2^23 rounds x to an integer
// kept in t's low mantissa bits, f = x - (t - 1.5*2^23) is in [-0.5, 0.5], 2^x = p(f) with the integer added to p's
// exponent field.  Valid for -126 <= x <= 126 (callers check the range per warp and row).  Relative error (float32
// Horner, measured): degree 3: 7.5e-5, degree 4: 2.7e-6 (mean 3.5e-8), degree 5: 2.3e-7.
template <int DEG>
__device__ __forceinline__ u64 ex2_poly2(u64 x2) {
    const u64 t = add2(x2, pk(12582912.f, 12582912.f));
    const u64 r = add2(t, pk(-12582912.f, -12582912.f));
    const u64 f = fma2(r, pk(-1.f, -1.f), x2);
    u64 p;
    if constexpr (DEG == 3) {
        p = fma2(pk(5.517080426e-02f, 5.517080426e-02f), f, pk(2.426092774e-01f, 2.426092774e-01f));
        p = fma2(p, f, pk(6.932609677e-01f, 6.932609677e-01f));
        p = fma2(p, f, pk(9.999281764e-01f, 9.999281764e-01f));
    } else if constexpr (DEG == 4) {
        p = fma2(pk(9.570019320e-03f, 9.570019320e-03f), f, pk(5.591766536e-02f, 5.591766536e-02f));
        p = fma2(p, f, pk(2.402474433e-01f, 2.402474433e-01f));
        p = fma2(p, f, pk(6.931218505e-01f, 6.931218505e-01f));
        p = fma2(p, f, pk(9.999992847e-01f, 9.999992847e-01f));
    } else {
        p = fma2(pk(1.327640144e-03f, 1.327640144e-03f), f, pk(9.675524198e-03f, 9.675524198e-03f));
        p = fma2(p, f, pk(5.550713465e-02f, 5.550713465e-02f));
        p = fma2(p, f, pk(2.402212024e-01f, 2.402212024e-01f));
        p = fma2(p, f, pk(6.931469440e-01f, 6.931469440e-01f));
        p = fma2(p, f, pk(1.000000119e+00f, 1.000000119e+00f));
    }
    float p_lo, p_hi, t_lo, t_hi;
    upk(p, p_lo, p_hi);
    upk(t, t_lo, t_hi);
    return pk(__int_as_float(__float_as_int(p_lo) + (__float_as_int(t_lo) << 23)),
              __int_as_float(__float_as_int(p_hi) + (__float_as_int(t_hi) << 23)));
}
__device__ __forceinline__ u64 ex2_mufu2(u64 x2) {
    float lo, hi;
    upk(x2, lo, hi);
    return pk(ex2(lo), ex2(hi));
}

__device__ __forceinline__ float set_ge(float a, float b) {      // 1.0f if a >= b else 0.0f, one FSET
    float r;
    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// 16-bit element pair -> packed fp32x2
template <typename T> struct Pair;
template <> struct Pair<__nv_bfloat16> {
    static constexpr uint32_t kPosInf2 = 0x7f807f80u;
    static constexpr uint32_t kNegInf2 = 0xff80ff80u;
    static __device__ __forceinline__ uint32_t vmin(uint32_t a, uint32_t b) {
        __nv_bfloat162 r = __hmin2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
    static __device__ __forceinline__ u64 widen(uint32_t w) { return pk(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
    static __device__ __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) {
        __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
    static __device__ __forceinline__ uint32_t narrow(float lo, float hi) {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ bool any_ge(uint32_t a, uint32_t b) {      // either half of a >= the same half of b
        __nv_bfloat162 r = __hge2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        return *reinterpret_cast<uint32_t*>(&r) != 0u;
    }
};
template <> struct Pair<__half> {
    static constexpr uint32_t kPosInf2 = 0x7c007c00u;
    static constexpr uint32_t kNegInf2 = 0xfc00fc00u;
    static __device__ __forceinline__ uint32_t vmin(uint32_t a, uint32_t b) {
        __half2 r = __hmin2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
    static __device__ __forceinline__ u64 widen(uint32_t w) {
        float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
        return pk(f.x, f.y);
    }
    static __device__ __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) {
        __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
    static __device__ __forceinline__ uint32_t narrow(float lo, float hi) {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ bool any_ge(uint32_t a, uint32_t b) {
        __half2 r = __hge2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
        return *reinterpret_cast<uint32_t*>(&r) != 0u;
    }
};


}  // namespace kd
