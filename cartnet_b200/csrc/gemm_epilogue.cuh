// Fused GEMM epilogue: parameter block shared by the fp32 SIMT kernel and the tcgen05 kernel (whose compile-time
// specialised epilogue lives in gemm_tc.cu) and the fp64 epilogue of the fp32 parity GEMM.
// Order (see cartnet_gemm_t in include/cartnet_b200.h):
//   v = acc + bias[col] + gather0[gidx0[row], col] + gather1[gidx1[row], col]
//   z_out = v ; v = act(v, z_in) ; v += resid ; out_f32 = v ; out_t = (T) v
#pragma once
#include "common.cuh"

namespace cartnet {

// element type of the tensors that the epilogue only touches ELEMENTWISE -- the gathered projections (added) and the
// pre-activations z (stored for, and read by, the SiLU' of the backward pass): T, except for bf16 pairs, where they are
// plain fp32 words. They are never contracted, so a hi|lo pair would cost the same 4 bytes, be less exact, and need a
// split on every store and a join on every load.
template <typename T> struct GatherOf { using type = T; };
template <> struct GatherOf<bf16p_t> { using type = float; };

template <typename T>
struct EpiParams {
    const float* bias;
    const typename GatherOf<T>::type* gather0;
    const int32_t* gidx0;
    const typename GatherOf<T>::type* gather1;
    const int32_t* gidx1;
    int64_t ldg;
    typename ZOf<T>::type* z_out;
    int64_t ldz;
    int act;
    const typename ZOf<T>::type* z_in;
    int64_t ldzin;
    const float* resid;
    int64_t ldr;
    float* out_f32;
    int64_t ldo;
    T* out_t;
    int64_t ldt;
};

template <typename T>
inline EpiParams<T> make_epi(const cartnet_gemm_t& d) {
    EpiParams<T> p;
    p.bias = d.bias;
    p.gather0 = (const typename GatherOf<T>::type*)d.gather0; p.gidx0 = d.gidx0;
    p.gather1 = (const typename GatherOf<T>::type*)d.gather1; p.gidx1 = d.gidx1;
    p.ldg = d.ldg;
    p.z_out = (typename ZOf<T>::type*)d.z_out; p.ldz = d.ldz;
    p.act = d.act;
    p.z_in = (const typename ZOf<T>::type*)d.z_in; p.ldzin = d.ldzin;
    p.resid = d.resid; p.ldr = d.ldr;
    p.out_f32 = d.out_f32; p.ldo = d.ldo;
    p.out_t = (T*)d.out_t; p.ldt = d.ldt;
    return p;
}

// per-row gather bases are resolved once per row by the caller
template <typename T>
struct EpiRow {
    const typename GatherOf<T>::type* g0;   // gather0 + gidx0[row]*ldg  (or null)
    const typename GatherOf<T>::type* g1;
};

template <typename T>
__device__ __forceinline__ EpiRow<T> epi_row(const EpiParams<T>& p, int64_t row) {
    EpiRow<T> r;
    r.g0 = p.gather0 ? p.gather0 + (int64_t)p.gidx0[row] * p.ldg : nullptr;
    r.g1 = p.gather1 ? p.gather1 + (int64_t)p.gidx1[row] * p.ldg : nullptr;
    return r;
}

// fp64 variant used by the fp32 parity GEMM: the accumulator arrives in double and bias / gathered projections /
// SiLU are applied before the single rounding to fp32 (the reference rounds after every op; its accumulated
// rounding noise is what the 1e-5 budget is mostly spent on)
__device__ __forceinline__ double silu_d(double v) { return v / (1.0 + exp(-v)); }
__device__ __forceinline__ double dsilu_d(double z) {
    const double s = 1.0 / (1.0 + exp(-z));
    return s * (1.0 + z * (1.0 - s));
}
__device__ __forceinline__ void epi_apply4_f64(const EpiParams<float>& p, const EpiRow<float>& r, int64_t row, int col,
                                               const double* acc) {
    double v[4] = {acc[0], acc[1], acc[2], acc[3]};
    if (p.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += (double)p.bias[col + j];
    }
    if (r.g0) {
        const float4 a = *reinterpret_cast<const float4*>(r.g0 + col);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
    }
    if (r.g1) {
        const float4 a = *reinterpret_cast<const float4*>(r.g1 + col);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
    }
    if (p.z_out) *reinterpret_cast<float4*>(p.z_out + row * p.ldz + col) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    if (p.act == CARTNET_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = silu_d(v[j]);
    } else if (p.act == CARTNET_ACT_MUL_DSILU) {
        const float4 z = *reinterpret_cast<const float4*>(p.z_in + row * p.ldzin + col);
        v[0] *= dsilu_d(z.x); v[1] *= dsilu_d(z.y); v[2] *= dsilu_d(z.z); v[3] *= dsilu_d(z.w);
    }
    if (p.resid) {
        const float4 a = *reinterpret_cast<const float4*>(p.resid + row * p.ldr + col);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
    }
    const float4 o = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + row * p.ldo + col) = o;
    if (p.out_t) *reinterpret_cast<float4*>(p.out_t + row * p.ldt + col) = o;
}

}  // namespace cartnet
