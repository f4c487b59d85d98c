// Shared helpers for libcartnet_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cartnet_b200.h"

namespace cartnet {

void set_error(const char* fmt, ...);
void count_launch();

#define CN_CHECK_ARG(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            cartnet::set_error(__VA_ARGS__); \
            return 2;                        \
        }                                    \
    } while (0)

#define CN_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            cartnet::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                \
                               cudaGetErrorString(_e));                                            \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define CN_LAUNCH_CHECK()                                                                          \
    do {                                                                                           \
        cudaError_t _e = cudaGetLastError();                                                       \
        cartnet::count_launch();                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            cartnet::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,            \
                               cudaGetErrorString(_e));                                            \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

constexpr int kNumSMs = 148;  // B200

// destination of a split-K weight gradient: up to 4 equal row blocks, each with its own base pointer
struct TnDst {
    float* c[4];
    int rows_per_blk;
};

// mean / biased variance (+ running-statistics update) from fp64 partial sums [nblocks][2][C] (edge_ops.cu)
int launch_colstats_final(const double* partial, int nblocks, int C, int64_t rows, const float* shift, float* mean, float* var,
                          float* running_mean, float* running_var, float momentum, cudaStream_t st);

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- operand type helpers -------------------------------------------------------------
// Storage type of the TF32 tensor-core mode: an fp32 word whose STORES round to nearest tf32 (10-bit mantissa,
// cvt.rna). tcgen05 kind::tf32 simply ignores the low 13 mantissa bits, i.e. truncates; rounding where the operand
// is produced removes that bias. Loads are plain fp32 loads.
struct tf32_t { float v; };
__device__ __forceinline__ float round_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <>
__device__ __forceinline__ float to_f32<tf32_t>(tf32_t v) { return v.v; }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ tf32_t from_f32<tf32_t>(float v) { tf32_t t; t.v = round_tf32(v); return t; }
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements, 16-byte (float) / 8-byte (bf16) aligned
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <>
__device__ __forceinline__ float4 load4<tf32_t>(const tf32_t* p) { return *reinterpret_cast<const float4*>(p); }
// read-only (non-coherent) variants: the compiler may hoist them above unrelated stores
template <typename T>
__device__ __forceinline__ float4 ldg4(const T* p);
template <>
__device__ __forceinline__ float4 ldg4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <>
__device__ __forceinline__ float4 ldg4<tf32_t>(const tf32_t* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <>
__device__ __forceinline__ float4 ldg4<__nv_bfloat16>(const __nv_bfloat16* p) {
    uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
// raw (unconverted) 4-element loads so that all global reads of a chunk can be issued before any math / store
template <typename T> struct Raw4;
template <> struct Raw4<tf32_t> { using type = float4; };
template <> struct Raw4<float> { using type = float4; };
template <> struct Raw4<__nv_bfloat16> { using type = uint2; };
template <typename T> __device__ __forceinline__ typename Raw4<T>::type ld_raw4(const T* p);
template <> __device__ __forceinline__ float4 ld_raw4<tf32_t>(const tf32_t* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ float4 ld_raw4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ uint2 ld_raw4<__nv_bfloat16>(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ float4 cvt_raw4(const float4& r) { return r; }
__device__ __forceinline__ float4 cvt_raw4(const uint2& r) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}

template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<tf32_t>(tf32_t* p, float4 v) {
    *reinterpret_cast<float4*>(p) = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
}
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}

// ---- activations (full-precision expf: the fp32 path is held to 1e-5) -------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float siluf_(float v) { return v * sigmoidf_(v); }
// d/dz [z * sigmoid(z)] = s * (1 + z * (1 - s))
__device__ __forceinline__ float dsiluf_(float z) {
    float s = sigmoidf_(z);
    return s * (1.0f + z * (1.0f - s));
}
// models/utils.py:87-91 (cutoff_lower == 0 branch)
__device__ __forceinline__ float cosine_cutoff(float d, float upper) {
    float c = 0.5f * (cosf(d * 3.14159265358979323846f / upper) + 1.0f);
    return d < upper ? c : 0.0f;
}

#define CN_DISPATCH_PREC(prec, ...)                                         \
    do {                                                                    \
        if ((prec) == CARTNET_PREC_FP32) {                                  \
            using T = float;                                                \
            __VA_ARGS__                                                     \
        } else if ((prec) == CARTNET_PREC_BF16) {                           \
            using T = __nv_bfloat16;                                        \
            __VA_ARGS__                                                     \
        } else if ((prec) == CARTNET_PREC_TF32) {                           \
            using T = cartnet::tf32_t;                                      \
            __VA_ARGS__                                                     \
        } else {                                                            \
            cartnet::set_error("unknown prec %d", (int)(prec));             \
            return 2;                                                       \
        }                                                                   \
    } while (0)

}  // namespace cartnet
