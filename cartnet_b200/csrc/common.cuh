// Shared helpers for libcartnet_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// Storage type of the split-precision mode (CARTNET_PREC_BF16X3): every element occupies 4 bytes, but the words are
// not individually meaningful. Each run of 64 consecutive elements (256 bytes, 256-byte aligned) holds the 64 bf16 HIGH
// parts (128 bytes) followed by the 64 bf16 LOW parts: value = hi + lo with hi = bf16(v), lo = bf16(v - hi), i.e. ~16
// mantissa bits. Seen as a bf16 matrix of twice the width, the hi / lo halves of a chunk are two 128-byte TMA boxes,
// which is what lets tcgen05 contract a pair tensor with three kind::f16 MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulate).
// Pointer arithmetic on bf16p_t* is in ELEMENTS, exactly like float*; the helpers below map an element address to the
// addresses of its two halves. Requirements: 256-byte aligned bases, row pitches and column offsets multiples of 64.
struct bf16p_t { uint32_t raw; };
__device__ __forceinline__ char* pair_hi_addr(const void* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    return reinterpret_cast<char*>((a & ~(uintptr_t)255) + ((a & (uintptr_t)255) >> 1));      // lo half: + 128
}
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ float2 bf162_to_float2(uint32_t w) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
// 4 floats -> hi (2 words) and lo (2 words) bf16 pairs
__device__ __forceinline__ void split4_bf16(const float4& v, uint2& hi, uint2& lo) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
    lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
}
__device__ __forceinline__ float4 join4_bf16(const uint2& hi, const uint2& lo) {
    const float2 a = bf162_to_float2(hi.x), b = bf162_to_float2(hi.y), c = bf162_to_float2(lo.x), d = bf162_to_float2(lo.y);
    return make_float4(a.x + c.x, a.y + c.y, b.x + d.x, b.y + d.y);
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

// one element at an element address (weight packing; not a hot path)
template <typename T>
__device__ __forceinline__ void store1(T* p, float v) { *p = from_f32<T>(v); }
template <>
__device__ __forceinline__ void store1<bf16p_t>(bf16p_t* p, float v) {
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    char* h = pair_hi_addr(p);
    *reinterpret_cast<__nv_bfloat16*>(h) = hi;
    *reinterpret_cast<__nv_bfloat16*>(h + 128) = lo;
}

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
template <>
__device__ __forceinline__ float4 load4<bf16p_t>(const bf16p_t* p) {
    const char* h = pair_hi_addr(p);
    return join4_bf16(*reinterpret_cast<const uint2*>(h), *reinterpret_cast<const uint2*>(h + 128));
}
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
template <>
__device__ __forceinline__ float4 ldg4<bf16p_t>(const bf16p_t* p) {
    const char* h = pair_hi_addr(p);
    return join4_bf16(__ldg(reinterpret_cast<const uint2*>(h)), __ldg(reinterpret_cast<const uint2*>(h + 128)));
}
// raw (unconverted) 4-element loads so that all global reads of a chunk can be issued before any math / store
template <typename T> struct Raw4;
template <> struct Raw4<tf32_t> { using type = float4; };
template <> struct Raw4<float> { using type = float4; };
template <> struct Raw4<__nv_bfloat16> { using type = uint2; };
template <> struct Raw4<bf16p_t> { using type = uint4; };      // {hi.x, hi.y, lo.x, lo.y}
template <typename T> __device__ __forceinline__ typename Raw4<T>::type ld_raw4(const T* p);
template <> __device__ __forceinline__ float4 ld_raw4<tf32_t>(const tf32_t* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ float4 ld_raw4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ uint2 ld_raw4<__nv_bfloat16>(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
template <> __device__ __forceinline__ uint4 ld_raw4<bf16p_t>(const bf16p_t* p) {
    const char* h = pair_hi_addr(p);
    const uint2 hi = __ldg(reinterpret_cast<const uint2*>(h)), lo = __ldg(reinterpret_cast<const uint2*>(h + 128));
    return make_uint4(hi.x, hi.y, lo.x, lo.y);
}
__device__ __forceinline__ float4 cvt_raw4(const uint4& r) { return join4_bf16(make_uint2(r.x, r.y), make_uint2(r.z, r.w)); }
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

template <>
__device__ __forceinline__ void store4<bf16p_t>(bf16p_t* p, float4 v) {
    uint2 hi, lo;
    split4_bf16(v, hi, lo);
    char* h = pair_hi_addr(p);
    *reinterpret_cast<uint2*>(h) = hi;
    *reinterpret_cast<uint2*>(h + 128) = lo;
}

// ---- fp16 storage of pre-activations (CARTNET_PREC_BF16X3 only) ------------------------------------------------------
// A pre-activation z is only ever re-read to evaluate silu'(z) in the backward pass. |d silu'(z)/dz * z| <= 0.25, so the
// 2^-11 relative rounding of an fp16 word moves silu'(z) by at most 1.2e-4 (absolute, on a factor of order 1): far inside
// the mode's 2e-3 budget, and half the bytes of the widest tensor a layer stores. Stores SATURATE (cvt.rn.satfinite), which
// is exact for the purpose: silu' is 1 / 0 to fp32 precision long before |z| = 65504.
struct half4raw { uint2 v; };
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));      // first source -> upper half
    return r;
}
__device__ __forceinline__ float4 cvt_raw4(const half4raw& r) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <> struct Raw4<__half> { using type = half4raw; };
template <> __device__ __forceinline__ half4raw ld_raw4<__half>(const __half* p) { half4raw r; r.v = __ldg(reinterpret_cast<const uint2*>(p)); return r; }
template <> __device__ __forceinline__ float4 load4<__half>(const __half* p) { half4raw r; r.v = *reinterpret_cast<const uint2*>(p); return cvt_raw4(r); }
template <> __device__ __forceinline__ float4 ldg4<__half>(const __half* p) { return cvt_raw4(ld_raw4<__half>(p)); }
template <> __device__ __forceinline__ void store4<__half>(__half* p, float4 v) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_half2_sat(v.x, v.y), pack_half2_sat(v.z, v.w));
}
// element type of the stored pre-activations (z_out / z_in of a GEMM epilogue, dsilu_mul): T, except fp16 in the pair mode
template <typename T> struct ZOf { using type = T; };
template <> struct ZOf<bf16p_t> { using type = __half; };

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
        } else if ((prec) == CARTNET_PREC_BF16X3) {                         \
            using T = cartnet::bf16p_t;                                     \
            __VA_ARGS__                                                     \
        } else {                                                            \
            cartnet::set_error("unknown prec %d", (int)(prec));             \
            return 2;                                                       \
        }                                                                   \
    } while (0)

}  // namespace cartnet
