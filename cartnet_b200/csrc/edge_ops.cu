// HBM-bound passes of the CartNet layer: edge featuriser, BatchNorm statistics, the gate /
// aggregate edge pass with its deterministic per-destination reduction, the node update, and the
// matching backward passes. Replaces the eager elementwise / scatter chains of
// /root/reference/models/cartnet.py:230-274 and models/utils.py:56-61,87-91.
//
// Layout: every [rows, C] tensor is row-major; a row is covered by C/4 consecutive threads with one
// 128-bit access each (C = 256 -> 64 threads read 1 KB contiguous). Reductions over rows use fp64
// accumulators in a fixed order: per-thread over a strided row set, then over row-lanes in shared
// memory, then over blocks in a second launch -- no atomics anywhere, results are reproducible.
#include "common.cuh"

namespace cartnet {

constexpr int kRedBlocksMax = 4 * kNumSMs;   // 592 partial rows at most

struct BnCoef {   // y = x * scale + shift ; xhat = (x - mean) * rstd
    float4 mean, rstd, scale, shift;
};
__device__ __forceinline__ BnCoef bn_coef(const float* mean, const float* var, const float* w, const float* b,
                                          float eps, int col) {
    BnCoef c;
    float4 m = *reinterpret_cast<const float4*>(mean + col);
    float4 v = *reinterpret_cast<const float4*>(var + col);
    float4 ww = w ? *reinterpret_cast<const float4*>(w + col) : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 bb = b ? *reinterpret_cast<const float4*>(b + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    c.mean = m;
    c.rstd = make_float4(1.0f / sqrtf(v.x + eps), 1.0f / sqrtf(v.y + eps), 1.0f / sqrtf(v.z + eps), 1.0f / sqrtf(v.w + eps));
    c.scale = make_float4(ww.x * c.rstd.x, ww.y * c.rstd.y, ww.z * c.rstd.z, ww.w * c.rstd.w);
    c.shift = bb;
    return c;
}
// BN(x) = (x - mean) * (w * rstd) + b   (same association as ATen's batch_norm CPU kernel: alpha*x+beta form
// differs only in the last ulp; the parity budget is 1e-5)
__device__ __forceinline__ float4 bn_apply(const BnCoef& c, float4 x) {
    return make_float4((x.x - c.mean.x) * c.scale.x + c.shift.x, (x.y - c.mean.y) * c.scale.y + c.shift.y,
                       (x.z - c.mean.z) * c.scale.z + c.shift.z, (x.w - c.mean.w) * c.scale.w + c.shift.w);
}
__device__ __forceinline__ float4 bn_hat(const BnCoef& c, float4 x) {
    return make_float4((x.x - c.mean.x) * c.rstd.x, (x.y - c.mean.y) * c.rstd.y, (x.z - c.mean.z) * c.rstd.z,
                       (x.w - c.mean.w) * c.rstd.w);
}

// ------------------------------------------------------------------------------------------
// generic column reduction: F(row, col) -> NV (2 or 3) float4 contributions; partial[blk][NV][C] fp64
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) colreduce_kernel(F f, int64_t rows, int C, double* __restrict__ partial) {
    constexpr int NV = F::NV;
    constexpr int U = 4;                     // rows per inner batch: loads of all U rows are issued together
    extern __shared__ double sm[];           // [lanes][NV][C]
    const int tpr = C >> 2;                  // threads per row
    const int lanes = 256 / tpr;             // row lanes per block
    const int rl = threadIdx.x / tpr, col = (threadIdx.x % tpr) * 4;
    const int64_t per_block = ceil_div64(rows, gridDim.x);
    const int64_t r0 = blockIdx.x * per_block;
    const int64_t r1 = (r0 + per_block < rows) ? r0 + per_block : rows;
    const typename F::State st = f.init(col);   // per-column constants (BatchNorm coefficients) hoisted out of the row loop
    double acc[NV][4];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[v][j] = 0.0;
    int64_t r = r0 + rl;
    for (; r + (int64_t)(U - 1) * lanes < r1; r += (int64_t)U * lanes) {
        typename F::In in[U];
#pragma unroll
        for (int u = 0; u < U; ++u) f.load(r + (int64_t)u * lanes, col, in[u]);
        if (F::F32_PARTIAL) {                // U values are summed in fp32 (fixed order), the running total stays in fp64
            float4 part[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) part[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float4 val[NV];
                f.compute(st, r + (int64_t)u * lanes, col, in[u], val);
#pragma unroll
                for (int v = 0; v < NV; ++v) { part[v].x += val[v].x; part[v].y += val[v].y; part[v].z += val[v].z; part[v].w += val[v].w; }
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                acc[v][0] += (double)part[v].x; acc[v][1] += (double)part[v].y; acc[v][2] += (double)part[v].z; acc[v][3] += (double)part[v].w;
            }
        } else {                             // BatchNorm statistics: every term goes straight into fp64 (E[x^2]-E[x]^2 cancels)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float4 val[NV];
                f.compute(st, r + (int64_t)u * lanes, col, in[u], val);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    acc[v][0] += (double)val[v].x; acc[v][1] += (double)val[v].y; acc[v][2] += (double)val[v].z; acc[v][3] += (double)val[v].w;
                }
            }
        }
    }
    for (; r < r1; r += lanes) {
        typename F::In in;
        f.load(r, col, in);
        float4 val[NV];
        f.compute(st, r, col, in, val);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            acc[v][0] += (double)val[v].x; acc[v][1] += (double)val[v].y; acc[v][2] += (double)val[v].z; acc[v][3] += (double)val[v].w;
        }
    }
    f.finish(col, acc);                      // per-column fix-up of the thread's totals (a no-op for most functors)
    double* mine = sm + (size_t)rl * NV * C;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int j = 0; j < 4; ++j) mine[v * C + col + j] = acc[v][j];
    __syncthreads();
    for (int j = threadIdx.x; j < NV * C; j += 256) {
        double s = 0;
        for (int l = 0; l < lanes; ++l) s += sm[(size_t)l * NV * C + j];
        partial[(size_t)blockIdx.x * NV * C + j] = s;
    }
}

enum { FIN_STATS = 0, FIN_SUMS = 1 };
// Sum of partial[b][j] over blocks b, one WARP per column j: lane l adds blocks l, l+32, ... in order, then a
// fixed-shape butterfly -- deterministic, and 32x less serial latency than one thread walking all blocks.
__device__ __forceinline__ double warp_block_sum(const double* __restrict__ partial, int nblocks, int stride, int j, int lane) {
    double s = 0;
    for (int b = lane; b < nblocks; b += 32) s += partial[(size_t)b * stride + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
// FIN_STATS: mean/var (+ running update) from (sum, sumsq); FIN_SUMS: out0[j] = sum of value j (j < n_out)
__global__ void __launch_bounds__(256)
colreduce_final_kernel(const double* __restrict__ partial, int nblocks, int C, int64_t rows, int mode,
                       float* __restrict__ out0, float* __restrict__ out1, float* __restrict__ run_mean,
                       float* __restrict__ run_var, float momentum, int n_out, int nv, const float* __restrict__ shift) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // one warp per output column
    if (mode == FIN_STATS) {
        if (c >= C) return;
        const double s = warp_block_sum(partial, nblocks, nv * C, c, lane);
        const double q = warp_block_sum(partial, nblocks, nv * C, C + c, lane);
        if (lane != 0) return;
        const double n = (double)rows;
        const double mean = s / n;
        double var = q / n - mean * mean;
        if (var < 0) var = 0;
        out0[c] = (float)mean;
        out1[c] = (float)var;
        if (run_mean) {
            const double unb = rows > 1 ? var * n / (n - 1.0) : var;
            // the statistics are those of the stored (possibly centred) tensor; the running mean tracks x + shift
            const double true_mean = mean + (shift ? (double)shift[c] : 0.0);
            run_mean[c] = (float)((1.0 - (double)momentum) * (double)run_mean[c] + (double)momentum * true_mean);
            run_var[c] = (float)((1.0 - (double)momentum) * (double)run_var[c] + (double)momentum * unb);
        }
    } else {
        if (c >= n_out) return;
        const double s = warp_block_sum(partial, nblocks, nv * C, c, lane);
        if (lane == 0) out0[c] = (float)s;
    }
}

static inline bool colreduce_shape_ok(int C) {
    int tpr = C / 4;
    return C % 4 == 0 && tpr >= 1 && tpr <= 256 && (tpr & (tpr - 1)) == 0;
}
static inline int colreduce_blocks(int64_t rows, int C) {
    int lanes = 256 / (C / 4);
    int64_t b = ceil_div64(rows, (int64_t)lanes * 16);
    if (b > kRedBlocksMax) b = kRedBlocksMax;
    if (b < 1) b = 1;
    return (int)b;
}

template <class F>
static int run_colreduce(F f, int64_t rows, int C, double* partial, int mode, float* out0, float* out1, float* rm,
                         float* rv, float momentum, int n_out, cudaStream_t st, const float* shift = nullptr) {
    const int nb = colreduce_blocks(rows, C);
    const int lanes = 256 / (C / 4);
    const size_t smem = (size_t)lanes * F::NV * C * sizeof(double);
    colreduce_kernel<F><<<nb, 256, smem, st>>>(f, rows, C, partial);
    CN_LAUNCH_CHECK();
    const int nthreads = mode == FIN_STATS ? C : n_out;
    colreduce_final_kernel<<<ceil_div(nthreads, 8), 256, 0, st>>>(partial, nb, C, rows, mode, out0, out1, rm, rv,
                                                                 momentum, n_out, F::NV, shift);
    CN_LAUNCH_CHECK();
    return 0;
}

int launch_colstats_final(const double* partial, int nblocks, int C, int64_t rows, const float* shift, float* mean, float* var,
                          float* running_mean, float* running_var, float momentum, cudaStream_t st) {
    colreduce_final_kernel<<<ceil_div(C, 8), 256, 0, st>>>(partial, nblocks, C, rows, FIN_STATS, mean, var, running_mean,
                                                          running_var, momentum, 0, 2, shift);
    CN_LAUNCH_CHECK();
    return 0;
}

// ---- functors: init(col) -> State ; load(row, col, In&) ; compute(State, row, col, In, float4 out[NV]) --------
struct NoState {};
template <typename TX>
struct StatsF {
    static constexpr int NV = 2;
    static constexpr bool F32_PARTIAL = false;
    using State = NoState;
    using In = float4;
    const TX* x; int64_t ld;
    __device__ State init(int) const { return State{}; }
    __device__ void load(int64_t r, int col, In& in) const { in = ldg4<TX>(x + r * ld + col); }
    __device__ void compute(const State&, int64_t, int, const In& a, float4* o) const {
        o[0] = a;
        o[1] = make_float4(a.x * a.x, a.y * a.y, a.z * a.z, a.w * a.w);
    }
    __device__ void finish(int, double (*)[4]) const {}
};
template <typename TX>
struct SumF {
    static constexpr int NV = 1;
    static constexpr bool F32_PARTIAL = true;
    using State = NoState;
    using In = float4;
    const TX* x; int64_t ld;
    __device__ State init(int) const { return State{}; }
    __device__ void load(int64_t r, int col, In& in) const { in = load4<TX>(x + r * ld + col); }
    __device__ void compute(const State&, int64_t, int, const In& a, float4* o) const { o[0] = a; }
    __device__ void finish(int, double (*)[4]) const {}
};
// y = (T)(dy * silu'(z)) written on the fly, column sums of y (a bias gradient) as the reduction
// z (a pre-activation) is only ever used elementwise: fp16 words in the bf16 pair mode (common.cuh::ZOf)

template <typename T>
struct DsiluMulF {
    static constexpr int NV = 1;
    static constexpr bool F32_PARTIAL = true;
    using State = NoState;
    struct In { float4 d, z; };
    const float* dy; int64_t ld_dy; const typename ZOf<T>::type* z; int64_t ldz; T* y; int64_t ldy;
    __device__ State init(int) const { return State{}; }
    __device__ void load(int64_t r, int col, In& in) const {
        in.d = __ldg(reinterpret_cast<const float4*>(dy + r * ld_dy + col));
        in.z = ldg4<typename ZOf<T>::type>(z + r * ldz + col);
    }
    __device__ void compute(const State&, int64_t r, int col, const In& in, float4* o) const {
        const float4 v = make_float4(in.d.x * dsiluf_(in.z.x), in.d.y * dsiluf_(in.z.y), in.d.z * dsiluf_(in.z.z), in.d.w * dsiluf_(in.z.w));
        store4<T>(y + r * ldy + col, v);
        o[0] = v;
    }
    __device__ void finish(int, double (*)[4]) const {}
};
struct NodeBwdF {
    static constexpr int NV = 2;
    static constexpr bool F32_PARTIAL = false;
    using State = BnCoef;
    struct In { float4 m, g; };
    const float* dx; const float* m; int D;
    const float *mean, *var, *w, *bias; float eps;
    __device__ State init(int col) const { return bn_coef(mean, var, w, bias, eps, col); }
    __device__ void load(int64_t r, int col, In& in) const {
        in.m = __ldg(reinterpret_cast<const float4*>(m + r * D + col));
        in.g = __ldg(reinterpret_cast<const float4*>(dx + r * D + col));
    }
    __device__ void compute(const State& c, int64_t, int, const In& in, float4* o) const {
        const float4 y = bn_apply(c, in.m), h = bn_hat(c, in.m);
        const float4 a = make_float4(in.g.x * dsiluf_(y.x), in.g.y * dsiluf_(y.y), in.g.z * dsiluf_(y.z), in.g.w * dsiluf_(y.w));
        o[0] = a;
        o[1] = make_float4(a.x * h.x, a.y * h.y, a.z * h.z, a.w * h.w);
    }
    __device__ void finish(int, double (*)[4]) const {}
};
// FAST: MUFU-based sigmoid / cosine for the tensor-core modes (their operands carry >= 2^-11 rounding anyway);
// the fp32 parity mode keeps expf / cosf / IEEE division.
template <bool FAST> __device__ __forceinline__ float sigmoid_sel(float v) {
    if (FAST) { float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v)); return fmaf(0.5f, t, 0.5f); }   // one MUFU op
    return sigmoidf_(v);
}
template <bool FAST> __device__ __forceinline__ float cutoff_sel(float d, float upper) {
    if (FAST) return d < upper ? 0.5f * (__cosf(d * 3.14159265358979323846f / upper) + 1.0f) : 0.0f;
    return cosine_cutoff(d, upper);
}
// Inputs are the T-typed tensors the forward pass saved: gn = (g - mean) * rstd (the normalised gate pre-activation,
// so no statistics are needed here) and s; the affine + sigmoid is recomputed in registers.
template <typename T, bool FAST>
struct EdgeBwdF {
    static constexpr int NV = 3;      // sum dghat | sum dghat*gn | sum ds  (the last one is d(bias) of MLP_aggr[2])
    static constexpr bool F32_PARTIAL = true;
    struct State { float4 w, b; };
    struct In { float4 gn, s, de, dmd; float dist; };
    const T *gn, *s; const float* dist; const int32_t* dst; const float *de, *dm; int D;
    const float *w, *bias; float radius; int use_env;
    T *ds_t, *dghat_t;
    // gvar != null: `gn` holds the stored (centred) pre-activation g, gn = (g - gmean) * rsqrt(gvar + eps) (gmean null = 0)
    // -- the forward pass then need not write a normalised copy. The affine BN(g) = gn w + b is folded into the
    // per-column constants (no extra registers in the row loop); the second sum is taken over dghat * g and turned into
    // sum dghat * gn = rstd (sum dghat g - mean sum dghat) once per thread in finish(), on the fp64 totals.
    const float *gmean, *gvar; float eps;
    __device__ State init(int col) const {
        State c;
        c.w = w ? *reinterpret_cast<const float4*>(w + col) : make_float4(1.f, 1.f, 1.f, 1.f);
        c.b = bias ? *reinterpret_cast<const float4*>(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (gvar) {
            const float4 v = *reinterpret_cast<const float4*>(gvar + col);
            const float4 mu = gmean ? *reinterpret_cast<const float4*>(gmean + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 rs = make_float4(1.0f / sqrtf(v.x + eps), 1.0f / sqrtf(v.y + eps), 1.0f / sqrtf(v.z + eps), 1.0f / sqrtf(v.w + eps));
            c.w = make_float4(c.w.x * rs.x, c.w.y * rs.y, c.w.z * rs.z, c.w.w * rs.w);
            c.b = make_float4(c.b.x - mu.x * c.w.x, c.b.y - mu.y * c.w.y, c.b.z - mu.z * c.w.z, c.b.w - mu.w * c.w.w);
        }
        return c;
    }
    __device__ void finish(int col, double (*acc)[4]) const {
        if (!gvar) return;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double rs = 1.0 / sqrt((double)gvar[col + j] + (double)eps);
            const double mu = gmean ? (double)gmean[col + j] : 0.0;
            acc[1][j] = rs * (acc[1][j] - mu * acc[0][j]);
        }
    }
    __device__ void load(int64_t r, int col, In& in) const {
        // read-only (non-coherent) loads of U rows are all in flight before the first store
        const int64_t o = r * D + col;
        in.gn = ldg4<T>(gn + o);
        in.s = ldg4<T>(s + o);
        in.de = de ? __ldg(reinterpret_cast<const float4*>(de + o)) : make_float4(0.f, 0.f, 0.f, 0.f);   // null = no gradient into e_out
        in.dmd = __ldg(reinterpret_cast<const float4*>(dm + (int64_t)__ldg(dst + r) * D + col));
        in.dist = use_env ? __ldg(dist + r) : 0.f;
    }
    __device__ void compute(const State& c, int64_t r, int col, const In& in, float4* out) const {
        const float env = use_env ? cutoff_sel<FAST>(in.dist, radius) : 1.0f;
        const int64_t o = r * D + col;
        const float sg[4] = {sigmoid_sel<FAST>(fmaf(in.gn.x, c.w.x, c.b.x)), sigmoid_sel<FAST>(fmaf(in.gn.y, c.w.y, c.b.y)),
                             sigmoid_sel<FAST>(fmaf(in.gn.z, c.w.z, c.b.z)), sigmoid_sel<FAST>(fmaf(in.gn.w, c.w.w, c.b.w))};
        // ds = sig * dm[dst] ; dsig = de_out + s * dm[dst] ; dghat = dsig * env * sg (1 - sg)
        const float4 ds = make_float4(env * sg[0] * in.dmd.x, env * sg[1] * in.dmd.y, env * sg[2] * in.dmd.z, env * sg[3] * in.dmd.w);
        store4<T>(ds_t + o, ds);
        const float4 a = make_float4((in.de.x + in.s.x * in.dmd.x) * env * sg[0] * (1.f - sg[0]),
                                     (in.de.y + in.s.y * in.dmd.y) * env * sg[1] * (1.f - sg[1]),
                                     (in.de.z + in.s.z * in.dmd.z) * env * sg[2] * (1.f - sg[2]),
                                     (in.de.w + in.s.w * in.dmd.w) * env * sg[3] * (1.f - sg[3]));
        store4<T>(dghat_t + o, a);
        out[0] = a;
        out[1] = make_float4(a.x * in.gn.x, a.y * in.gn.y, a.z * in.gn.z, a.w * in.gn.w);
        out[2] = ds;
    }
};

// ------------------------------------------------------------------------------------------
// forward edge pass: gate, residual, deterministic per-destination sum
// ------------------------------------------------------------------------------------------
// R = math type: double for the fp32 parity mode (the reference's own fp32 rounding already uses most of the
// 1e-5 budget in training mode), float for the tensor-core modes.
template <typename R> __device__ __forceinline__ R sigmoid_r(R v);
template <> __device__ __forceinline__ float sigmoid_r<float>(float v) { return sigmoidf_(v); }
template <> __device__ __forceinline__ double sigmoid_r<double>(double v) { return 1.0 / (1.0 + exp(-v)); }
template <typename R> __device__ __forceinline__ R cutoff_r(float d, float upper);
template <> __device__ __forceinline__ float cutoff_r<float>(float d, float upper) { return cosine_cutoff(d, upper); }
template <> __device__ __forceinline__ double cutoff_r<double>(float d, float upper) {
    // same operation order as models/utils.py:88 in fp32 for the argument, then a double cosine
    const float arg = d * 3.14159265358979323846f / upper;
    return d < upper ? 0.5 * (cos((double)arg) + 1.0) : 0.0;
}

// FAST (tensor-core modes): MUFU sigmoid / cosine, like the backward pass; the fp32 parity mode computes in double.
template <typename R, bool FAST> __device__ __forceinline__ R gate_sigmoid(R v) { return sigmoid_r<R>(v); }
// forward gate: EX2 + RCP (relative error ~1e-6) rather than tanh.approx (2^-11): sigma is accumulated into the fp32
// residual stream e over all layers, and the tf32 mode is held to 2e-3 end to end
template <> __device__ __forceinline__ float gate_sigmoid<float, true>(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
template <typename R, bool FAST> __device__ __forceinline__ R gate_cutoff(float d, float upper) { return cutoff_r<R>(d, upper); }
template <> __device__ __forceinline__ float gate_cutoff<float, true>(float d, float upper) { return cutoff_sel<true>(d, upper); }

// D/4 threads per destination node walk its CSR row in order (deterministic sum); the loads of U consecutive edges
// are issued together so that each thread keeps U x 40 bytes in flight.
template <typename T, typename R, bool FAST>
__global__ void __launch_bounds__(256, FAST ? 2 : 1)
edge_gate_aggregate_kernel(const T* __restrict__ g, const T* __restrict__ s, const float* __restrict__ e,
                           const float* __restrict__ dist, const int32_t* __restrict__ row_ptr, int num_nodes, int D,
                           const float* mean, const float* var, const float* w, const float* bias, float eps,
                           float radius, int use_env, float* __restrict__ e_out, T* __restrict__ e_out_t,
                           T* __restrict__ gn_t, float* __restrict__ m) {
    constexpr int U = 4;
    const int tpr = D >> 2, npb = 256 / tpr;
    const int node = blockIdx.x * npb + threadIdx.x / tpr;
    const int col = (threadIdx.x % tpr) * 4;
    if (node >= num_nodes) return;
    R mu[4], rs[4], sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        mu[j] = mean ? (R)mean[col + j] : (R)0;              // null mean: g arrives centred on the statistics' mean (eval mode)
        rs[j] = (R)1 / sqrt((R)var[col + j] + (R)eps);
        sc[j] = (w ? (R)w[col + j] : (R)1) * rs[j];
        sh[j] = bias ? (R)bias[col + j] : (R)0;
    }
    R acc[4] = {0, 0, 0, 0};
    const int k0 = row_ptr[node], k1 = row_ptr[node + 1];
    auto edge = [&](int k, const typename Raw4<T>::type& gr, const typename Raw4<T>::type& sr, const float4& e4, float dk) {
        const int64_t o = (int64_t)k * D + col;
        const float4 g4 = cvt_raw4(gr), s4 = cvt_raw4(sr);
        const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w}, ee[4] = {e4.x, e4.y, e4.z, e4.w};
        const R env = use_env ? gate_cutoff<R, FAST>(dk, radius) : (R)1;
        float eo[4], gn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const R gc = (R)gg[j] - mu[j];
            gn[j] = (float)(gc * rs[j]);                            // normalised pre-activation, saved for backward
            const R gh = gc * sc[j] + sh[j];
            const float sig = (float)(env * gate_sigmoid<R, FAST>(gh));   // the reference materialises sigma_ij in fp32
            eo[j] = ee[j] + sig;                                    // cartnet.py:225
            acc[j] += (R)sig * (R)ss[j];                            // cartnet.py:259
        }
        const float4 eo4 = make_float4(eo[0], eo[1], eo[2], eo[3]);
        *reinterpret_cast<float4*>(e_out + o) = eo4;
        if (e_out_t) store4<T>(e_out_t + o, eo4);
        if (gn_t) store4<T>(gn_t + o, make_float4(gn[0], gn[1], gn[2], gn[3]));
    };
    int k = k0;
    for (; k + U <= k1; k += U) {
        typename Raw4<T>::type g4[U], s4[U];                 // unconverted (8 bytes for bf16) until they are used
        float4 e4[U];
        float dk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t o = (int64_t)(k + u) * D + col;
            g4[u] = ld_raw4<T>(g + o);
            s4[u] = ld_raw4<T>(s + o);
            e4[u] = __ldg(reinterpret_cast<const float4*>(e + o));
            dk[u] = use_env ? __ldg(dist + k + u) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) edge(k + u, g4[u], s4[u], e4[u], dk[u]);
    }
    for (; k < k1; ++k) {
        const int64_t o = (int64_t)k * D + col;
        edge(k, ld_raw4<T>(g + o), ld_raw4<T>(s + o), __ldg(reinterpret_cast<const float4*>(e + o)), use_env ? __ldg(dist + k) : 0.f);
    }
    *reinterpret_cast<float4*>(m + (int64_t)node * D + col) = make_float4((float)acc[0], (float)acc[1], (float)acc[2], (float)acc[3]);
}

template <typename T, typename R>
__global__ void node_update_kernel(const float* __restrict__ m, const float* __restrict__ x, int64_t total4, int D,
                                   const float* mean, const float* var, const float* w, const float* bias, float eps,
                                   float* __restrict__ x_out, T* __restrict__ x_out_t) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int64_t o = i * 4;
    const int col = (int)(o % D);
    const float4 m4 = *reinterpret_cast<const float4*>(m + o);
    const float4 x4 = *reinterpret_cast<const float4*>(x + o);
    const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, xx[4] = {x4.x, x4.y, x4.z, x4.w};
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const R rstd = (R)1 / sqrt((R)var[col + j] + (R)eps);
        const R y = ((R)mm[j] - (R)mean[col + j]) * ((w ? (R)w[col + j] : (R)1) * rstd) + (bias ? (R)bias[col + j] : (R)0);
        r[j] = (float)(y * sigmoid_r<R>(y)) + xx[j];                // cartnet.py:223
    }
    const float4 r4 = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(x_out + o) = r4;
    if (x_out_t) store4<T>(x_out_t + o, r4);
}

__global__ void node_bwd_apply_kernel(const float* __restrict__ dx, const float* __restrict__ m, int64_t total4, int D,
                                      int num_nodes, const float* mean, const float* var, const float* w,
                                      const float* bias, float eps, const float* __restrict__ sums, int training,
                                      float* __restrict__ dm) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int64_t o = i * 4;
    const int col = (int)(o % D);
    const BnCoef c = bn_coef(mean, var, w, bias, eps, col);
    const float4 mm = *reinterpret_cast<const float4*>(m + o);
    const float4 g = *reinterpret_cast<const float4*>(dx + o);
    const float4 y = bn_apply(c, mm), h = bn_hat(c, mm);
    float dy[4] = {g.x * dsiluf_(y.x), g.y * dsiluf_(y.y), g.z * dsiluf_(y.z), g.w * dsiluf_(y.w)};
    const float hh[4] = {h.x, h.y, h.z, h.w};
    const float sc[4] = {c.scale.x, c.scale.y, c.scale.z, c.scale.w};
    float out[4];
    const float inv_n = 1.0f / (float)num_nodes;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float corr = training ? (sums[col + j] * inv_n + hh[j] * sums[D + col + j] * inv_n) : 0.f;
        out[j] = sc[j] * (dy[j] - corr);
    }
    *reinterpret_cast<float4*>(dm + o) = make_float4(out[0], out[1], out[2], out[3]);
}

// 16-byte vectors of the operand type: 8 bf16 / 4 fp32 words per thread and instruction
template <typename T> struct Vec16 {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const T* p, float* v) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
    static __device__ __forceinline__ void store(T* p, const float* v) { store4<T>(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct Vec16<bf16p_t> {      // 8 elements: 16 bytes of high parts + 16 bytes of low parts
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const bf16p_t* p, float* v) {
        const char* h = pair_hi_addr(p);
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(h)), lo = __ldg(reinterpret_cast<const uint4*>(h + 128));
        const float4 a = join4_bf16(make_uint2(hi.x, hi.y), make_uint2(lo.x, lo.y)), b = join4_bf16(make_uint2(hi.z, hi.w), make_uint2(lo.z, lo.w));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(bf16p_t* p, const float* v) {
        uint2 h0, l0, h1, l1;
        split4_bf16(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
        split4_bf16(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
        char* h = pair_hi_addr(p);
        *reinterpret_cast<uint4*>(h) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(h + 128) = make_uint4(l0.x, l0.y, l1.x, l1.y);
    }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// dg = w * rstd * (dghat - [train](sum1/E + gn * sum2/E)). Each thread owns one 16-byte column group (coefficients
// hoisted); a block walks a CONTIGUOUS range of rows, U row groups (U x 256 threads x 16 B per tensor) in flight, so that
// every request is a full line and consecutive requests stay in one DRAM page.
template <typename T>
__global__ void __launch_bounds__(256)
edge_bwd_apply_kernel(const T* __restrict__ gn, const T* __restrict__ dghat, int64_t rows, int D, const float* var,
                      const float* w, float eps, const float* __restrict__ sums, int training, T* __restrict__ dg_t,
                      int64_t rows_per_block, const float* gmean, int input_is_g) {
    constexpr int U = 4, V = Vec16<T>::N;
    const int tpr = D / V, lanes = 256 / tpr;
    const int rl = threadIdx.x / tpr, col = (threadIdx.x % tpr) * V;
    float sc[V], c1[V], c2[V];
    const float inv_n = 1.0f / (float)rows;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const float rstd = 1.0f / sqrtf(var[col + j] + eps);
        sc[j] = (w ? w[col + j] : 1.0f) * rstd;
        c1[j] = training ? sums[col + j] * inv_n : 0.f;
        c2[j] = training ? sums[D + col + j] * inv_n : 0.f;
        if (input_is_g) {      // the tensor holds the stored (centred) g: gn * c2 = (g - mean) * rstd * c2
            c1[j] -= (gmean ? gmean[col + j] : 0.f) * rstd * c2[j];
            c2[j] *= rstd;
        }
    }
    const int64_t r0 = blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
    int64_t r = r0 + rl;
    for (; r + (int64_t)(U - 1) * lanes < r1; r += (int64_t)U * lanes) {
        float h[U][V], d[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t o = (r + (int64_t)u * lanes) * D + col;
            Vec16<T>::load(gn + o, h[u]);
            Vec16<T>::load(dghat + o, d[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float o4[V];
#pragma unroll
            for (int j = 0; j < V; ++j) o4[j] = sc[j] * (d[u][j] - (c1[j] + h[u][j] * c2[j]));
            Vec16<T>::store(dg_t + (r + (int64_t)u * lanes) * D + col, o4);
        }
    }
    for (; r < r1; r += lanes) {
        float h[V], d[V], o4[V];
        Vec16<T>::load(gn + r * D + col, h);
        Vec16<T>::load(dghat + r * D + col, d);
#pragma unroll
        for (int j = 0; j < V; ++j) o4[j] = sc[j] * (d[j] - (c1[j] + h[j] * c2[j]));
        Vec16<T>::store(dg_t + r * D + col, o4);
    }
}

// out[n, :] = sum_{k in CSR row n} x[perm ? perm[k] : k, :]
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
segment_sum_kernel(const T* __restrict__ x, int64_t ldx, const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm,
                   int num_nodes, int C, TO* __restrict__ out, int64_t ldo) {
    const int tpr = C >> 2, npb = 256 / tpr;
    const int node = blockIdx.x * npb + threadIdx.x / tpr;
    const int col = (threadIdx.x % tpr) * 4;
    if (node >= num_nodes) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k0 = ptr[node], k1 = ptr[node + 1];
    for (int k = k0; k < k1; ++k) {
        const int64_t r = perm ? (int64_t)perm[k] : (int64_t)k;
        const float4 v = load4<T>(x + r * ldx + col);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    store4<TO>(out + (int64_t)node * ldo + col, acc);
}

// Few, long segments (per-crystal sums over nodes: 64 segments of ~200 rows): one block per (segment, 128-column chunk),
// 8 row lanes that each add every 8th row in order, combined in a fixed order -- instead of one thread column walking
// a whole segment. Deterministic; identity row order only.
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
segment_sum_wide_kernel(const T* __restrict__ x, int64_t ldx, const int32_t* __restrict__ ptr, int C, TO* __restrict__ out,
                        int64_t ldo) {
    __shared__ float4 sm[8][32];
    const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int seg = blockIdx.x, col = blockIdx.y * 128 + lane * 4;
    const int k0 = ptr[seg], k1 = ptr[seg + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < C) {
        for (int k = k0 + rl; k < k1; k += 8) {
            const float4 v = load4<T>(x + (int64_t)k * ldx + col);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    sm[rl][lane] = acc;
    __syncthreads();
    if (rl == 0 && col < C) {
#pragma unroll
        for (int q = 1; q < 8; ++q) {
            const float4 v = sm[q][lane];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        store4<TO>(out + (int64_t)seg * ldo + col, acc);
    }
}

// Both transposed lifts in one pass: out[n, 0:C] = sum over the dst-CSR row of n of x[k, :], out[n, C:2C] = sum over the
// src-CSR row of n of x[perm[k], :]. Every row of x is wanted twice, once through each CSR; the edges that leave n end
// at n's neighbours, i.e. inside the same crystal, so while a block of nodes is being processed the second read of a
// row finds it in L2 (a crystal's rows are ~10 MB) instead of HBM, which two separate launches over [E, C] cannot.
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
segment_sum_pair_kernel(const T* __restrict__ x, int64_t ldx, const int32_t* __restrict__ row_ptr,
                        const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ perm, int num_nodes, int C,
                        TO* __restrict__ out, int64_t ldo) {
    constexpr int U = 4;
    const int tpr = C >> 2, npb = 256 / tpr;
    const int node = blockIdx.x * npb + threadIdx.x / tpr;
    const int col = (threadIdx.x % tpr) * 4;
    if (node >= num_nodes) return;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int32_t* ptr = pass ? col_ptr : row_ptr;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k0 = ptr[node], k1 = ptr[node + 1];
        int k = k0;
        for (; k + U <= k1; k += U) {
            typename Raw4<T>::type v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t r = pass ? (int64_t)__ldg(perm + k + u) : (int64_t)(k + u);
                v[u] = ld_raw4<T>(x + r * ldx + col);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float4 f = cvt_raw4(v[u]);
                acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
            }
        }
        for (; k < k1; ++k) {
            const int64_t r = pass ? (int64_t)__ldg(perm + k) : (int64_t)k;
            const float4 f = cvt_raw4(ld_raw4<T>(x + r * ldx + col));
            acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
        }
        store4<TO>(out + (int64_t)node * ldo + pass * C + col, acc);
    }
}

// Same result set as segment_sum_pair_kernel, laid out for L2 reuse. Every row of x is wanted twice (once through each CSR)
// and both uses belong to nodes of the same crystal, but with one thread column per (node, 4 columns) ~2 400 nodes are in
// flight at once: ~12 ADP crystals x 43 MB of rows (4-byte operands, C = 512) against 126 MB of L2, and the second read of
// a row went back to HBM (ncu: 2.66 GB read per launch for a 1.41 GB tensor). Here RL row lanes share a node (lane q adds
// rows q, q + RL, ... of the segment in order, the RL partial sums are combined in lane order: still one fixed order), so
// the same number of loads is in flight with RL x fewer nodes -- ~300 nodes = 1.5 crystals -- and the second read hits L2:
// ncu 1.44 GB of DRAM reads per launch instead of 2.75. The launch itself is only 5 % shorter (397 vs 419 us): L2 -> SM
// delivery (~7 TB/s here) is no faster than HBM, so on this part L2 reuse saves DRAM traffic, not time.
template <typename T, typename TO>
__global__ void __launch_bounds__(1024, 2)
segment_sum_pair_lanes_kernel(const T* __restrict__ x, int64_t ldx, const int32_t* __restrict__ row_ptr,
                              const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ perm, int num_nodes, int C,
                              int RL, TO* __restrict__ out, int64_t ldo) {
    constexpr int U = 2;                                     // rows of one lane in flight (4: 431 us, 2: 397 us at ADP-64)
    extern __shared__ float4 seg_sm[];                       // [RL][thread columns of the block]
    const int tpr = C >> 2;
    const int cpb = blockDim.x / RL;                         // thread columns per block = tpr * (nodes per block)
    const int tc = threadIdx.x % cpb, rl = threadIdx.x / cpb;
    const int node = blockIdx.x * (cpb / tpr) + tc / tpr;
    const int col = (tc % tpr) * 4;
    const bool live = node < num_nodes;
    const T* xc = x + col;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int32_t* ptr = pass ? col_ptr : row_ptr;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            const int k1 = ptr[node + 1];
            int k = ptr[node] + rl;
            for (; k + (U - 1) * RL < k1; k += U * RL) {
                int32_t r[U];
#pragma unroll
                for (int u = 0; u < U; ++u) r[u] = pass ? __ldg(perm + k + u * RL) : k + u * RL;
                typename Raw4<T>::type v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = ld_raw4<T>(xc + (int64_t)r[u] * ldx);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float4 f = cvt_raw4(v[u]);
                    acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
                }
            }
            for (; k < k1; k += RL) {
                const int32_t r0 = pass ? __ldg(perm + k) : k;
                const float4 f = cvt_raw4(ld_raw4<T>(xc + (int64_t)r0 * ldx));
                acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
            }
        }
        if (pass) __syncthreads();                           // lane 0 has finished reading the first pass's partial sums
        seg_sm[rl * cpb + tc] = acc;
        __syncthreads();
        if (rl == 0 && live) {
            for (int q = 1; q < RL; ++q) {
                const float4 v = seg_sm[q * cpb + tc];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            store4<TO>(out + (int64_t)node * ldo + pass * C + col, acc);
        }
    }
}

template <typename T>
__global__ void dsilu_mul_kernel(const float* __restrict__ dy, int64_t ld_dy, const typename ZOf<T>::type* __restrict__ z, int64_t ldz,
                                 T* __restrict__ y, int64_t ldy, int64_t rows, int C) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int c4 = C >> 2;
    if (i >= rows * c4) return;
    const int64_t r = i / c4;
    const int col = (int)(i % c4) * 4;
    const float4 d = *reinterpret_cast<const float4*>(dy + r * ld_dy + col);
    const float4 zz = load4<typename ZOf<T>::type>(z + r * ldz + col);
    store4<T>(y + r * ldy + col, make_float4(d.x * dsiluf_(zz.x), d.y * dsiluf_(zz.y), d.z * dsiluf_(zz.z), d.w * dsiluf_(zz.w)));
}

template <typename T>
__global__ void cast_rows_kernel(const float* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t ldd,
                                 int64_t rows, int C) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int c4 = C >> 2;
    if (i >= rows * c4) return;
    const int64_t r = i / c4;
    const int col = (int)(i % c4) * 4;
    store4<T>(dst + r * ldd + col, *reinterpret_cast<const float4*>(src + r * lds + col));
}

template <typename T>
__global__ void uncast_rows_kernel(const T* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd,
                                   int64_t rows, int C) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int c4 = C >> 2;
    if (i >= rows * c4) return;
    const int64_t r = i / c4;
    const int col = (int)(i % c4) * 4;
    *reinterpret_cast<float4*>(dst + r * ldd + col) = load4<T>(src + r * lds + col);
}

// feat[e, :] = [ cut(d) exp(-beta_k (exp(-alpha d) - mu_k)^2) (k < R) ; cart_dir (3, unless invariant) ; 1 ; 0 ... ]
// The first padding column (if ld leaves one) holds 1: the matching column of the zero-padded weight is 0, so the forward
// is unchanged, while in the backward the bias gradient sum_e dz falls out of the weight-gradient GEMM dz^T feat as that
// column -- no separate reduction pass over [E, 2D].
template <typename T, typename R>
__global__ void edge_features_kernel(const float* __restrict__ cart_dist, const float* __restrict__ cart_dir,
                                     const float* __restrict__ means, const float* __restrict__ betas, int nrbf,
                                     float upper, int invariant, int64_t E, T* __restrict__ feat, int ld) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int c4 = ld >> 2;
    if (i >= E * c4) return;
    const int64_t e = i / c4;
    const int col = (int)(i % c4) * 4;
    const float d = cart_dist[e];
    const float alpha = 5.0f / upper;                 // models/utils.py:26 (cutoff_lower = 0)
    const R ex = exp((R)(alpha * (-d)));              // models/utils.py:60
    const R cut = cutoff_r<R>(d, upper);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = col + j;
        if (k < nrbf) {
            const R t = ex - (R)means[k];
            v[j] = (float)(cut * exp(-(R)betas[k] * (t * t)));
        } else if (!invariant && k < nrbf + 3) {
            v[j] = cart_dir[e * 3 + (k - nrbf)];
        } else {
            v[j] = (k == nrbf + (invariant ? 0 : 3)) ? 1.f : 0.f;
        }
    }
    store4<T>(feat + e * ld + col, make_float4(v[0], v[1], v[2], v[3]));
}

// Centre of the gate pre-activation g = H_g G2^T + bg2, so that g - center can be stored in T without losing the
// bits BatchNorm needs (|mean(g)| is ~15 std at initialisation). Any shift within ~1 std of the true mean will do --
// BatchNorm removes a per-column shift exactly -- so the batch mean of a row SAMPLE of H_g is used (sums over `rows`
// sampled rows): center = G2 mean_s(H_g) + bg2, bias_c = bg2 - center. Eval mode: center = running_mean.
__global__ void __launch_bounds__(256)
gate_center_kernel(const float* __restrict__ hsum, int rows, const float* __restrict__ G2, const float* __restrict__ bg2,
                   const float* __restrict__ running_mean, int training, int D, float* __restrict__ bias_c,
                   float* __restrict__ center) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);       // one warp per output channel
    const int lane = threadIdx.x & 31;
    if (j >= D) return;
    if (!training) {
        if (lane == 0) { center[j] = running_mean[j]; bias_c[j] = bg2[j] - running_mean[j]; }
        return;
    }
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) acc = fmaf(G2[(int64_t)j * D + k], hsum[k], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const float mu = acc / (float)rows;
        bias_c[j] = -mu;
        center[j] = bg2[j] + mu;
    }
}

static inline bool row_shape_ok(int D) {
    int tpr = D / 4;
    return D % 4 == 0 && tpr >= 1 && tpr <= 256 && (256 % tpr) == 0;
}

}  // namespace cartnet

using namespace cartnet;

extern "C" {

int64_t cartnet_colstats_workspace(int32_t C) { return (int64_t)kRedBlocksMax * 3 * C * (int64_t)sizeof(double); }

int cartnet_colstats(const void* x, int32_t x_is_t, int32_t prec, int64_t rows, int32_t C, int64_t ld, const float* shift,
                     float* mean, float* var, float* running_mean, float* running_var, float momentum, double* partial,
                     cartnet_stream_t stream) {
    CN_CHECK_ARG(x && mean && var && partial && rows > 0, "colstats: bad arguments");
    CN_CHECK_ARG(colreduce_shape_ok(C) && ld % 4 == 0, "colstats: C/4 must be a power of two <= 256 (C=%d)", C);
    CN_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "colstats: running stats must come in pairs");
    cudaStream_t st = (cudaStream_t)stream;
    if (x_is_t && prec == CARTNET_PREC_BF16) {
        StatsF<__nv_bfloat16> f{(const __nv_bfloat16*)x, ld};
        return run_colreduce(f, rows, C, partial, FIN_STATS, mean, var, running_mean, running_var, momentum, 0, st, shift);
    }
    if (x_is_t && prec == CARTNET_PREC_BF16X3) {
        StatsF<bf16p_t> f{(const bf16p_t*)x, ld};
        return run_colreduce(f, rows, C, partial, FIN_STATS, mean, var, running_mean, running_var, momentum, 0, st, shift);
    }
    StatsF<float> f{(const float*)x, ld};
    return run_colreduce(f, rows, C, partial, FIN_STATS, mean, var, running_mean, running_var, momentum, 0, st, shift);
}

int cartnet_gate_center(const void* H_g, int64_t ldh, int64_t num_edges, int32_t D, const float* G2, const float* bg2,
                        const float* running_mean, int32_t training, int32_t prec, float* bias_c, float* center,
                        float* hsum, double* partial, cartnet_stream_t stream) {
    CN_CHECK_ARG(G2 && bg2 && bias_c && center && D > 0, "gate_center: null pointer");
    CN_CHECK_ARG(training || running_mean, "gate_center: eval mode needs running_mean");
    cudaStream_t st = (cudaStream_t)stream;
    int rows = 0;
    if (training) {
        CN_CHECK_ARG(H_g && hsum && partial && num_edges > 0, "gate_center: training mode needs H_g, hsum, partial and edges");
        // <= 4096 rows at a fixed stride: the sample mean is within std/64 of the batch mean
        const int64_t step = num_edges > 4096 ? num_edges / 4096 : 1;
        rows = (int)((num_edges + step - 1) / step);
        if (int rc = cartnet_colsum(H_g, 1, prec, rows, D, ldh * step, hsum, partial, stream)) return rc;
    }
    gate_center_kernel<<<ceil_div(D, 8), 256, 0, st>>>(hsum, rows, G2, bg2, running_mean, training, D, bias_c, center);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_colsum(const void* x, int32_t x_is_t, int32_t prec, int64_t rows, int32_t C, int64_t ld, float* out,
                   double* partial, cartnet_stream_t stream) {
    CN_CHECK_ARG(x && out && partial && rows >= 0, "colsum: bad arguments");
    CN_CHECK_ARG(colreduce_shape_ok(C) && ld % 4 == 0, "colsum: C/4 must be a power of two <= 256 (C=%d)", C);
    cudaStream_t st = (cudaStream_t)stream;
    if (x_is_t && prec == CARTNET_PREC_BF16) {
        SumF<__nv_bfloat16> f{(const __nv_bfloat16*)x, ld};
        return run_colreduce(f, rows, C, partial, FIN_SUMS, out, nullptr, nullptr, nullptr, 0.f, C, st);
    }
    if (x_is_t && prec == CARTNET_PREC_BF16X3) {
        SumF<bf16p_t> f{(const bf16p_t*)x, ld};
        return run_colreduce(f, rows, C, partial, FIN_SUMS, out, nullptr, nullptr, nullptr, 0.f, C, st);
    }
    SumF<float> f{(const float*)x, ld};
    return run_colreduce(f, rows, C, partial, FIN_SUMS, out, nullptr, nullptr, nullptr, 0.f, C, st);
}

int cartnet_edge_features(const float* cart_dist, const float* cart_dir, const float* means, const float* betas,
                          int32_t num_rbf, float cutoff_upper, int32_t invariant, int64_t num_edges, void* feat,
                          int32_t ld, int32_t prec, cartnet_stream_t stream) {
    if (num_edges <= 0) return 0;      // empty graph: zero-sized tensors carry null data pointers
    CN_CHECK_ARG(cart_dist && means && betas && feat, "edge_features: null pointer");
    CN_CHECK_ARG(invariant || cart_dir, "edge_features: cart_dir required unless invariant");
    CN_CHECK_ARG(ld % 4 == 0 && ld >= num_rbf + (invariant ? 0 : 3), "edge_features: ld=%d too small", ld);
    if (num_edges <= 0) return 0;
    const int64_t total = num_edges * (ld / 4);
    if (prec == CARTNET_PREC_FP32) {
        edge_features_kernel<float, double><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
            cart_dist, cart_dir, means, betas, num_rbf, cutoff_upper, invariant, num_edges, (float*)feat, ld);
    } else {
        CN_DISPATCH_PREC(prec, {
            edge_features_kernel<T, float><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
                cart_dist, cart_dir, means, betas, num_rbf, cutoff_upper, invariant, num_edges, (T*)feat, ld);
        });
    }
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_edge_gate_aggregate(const void* g_t, const void* s_t, const float* e, const float* dist,
                                const int32_t* row_ptr, int32_t num_nodes, int64_t num_edges, int32_t D,
                                const float* bn_mean, const float* bn_var, const float* bn_weight, const float* bn_bias,
                                float eps, float radius, int32_t use_envelope, float* e_out, void* e_out_t, void* gn_t,
                                int32_t prec, float* m, cartnet_stream_t stream) {
    CN_CHECK_ARG(row_ptr && bn_var && m, "edge_gate_aggregate: null pointer");
    CN_CHECK_ARG(num_edges == 0 || (g_t && s_t && e && dist && e_out), "edge_gate_aggregate: null edge tensor");
    CN_CHECK_ARG(row_shape_ok(D), "edge_gate_aggregate: unsupported D=%d", D);
    if (num_nodes <= 0) return 0;
    const int npb = 256 / (D / 4);
    if (prec == CARTNET_PREC_FP32) {
        edge_gate_aggregate_kernel<float, double, false><<<ceil_div(num_nodes, npb), 256, 0, (cudaStream_t)stream>>>(
            (const float*)g_t, (const float*)s_t, e, dist, row_ptr, num_nodes, D, bn_mean, bn_var, bn_weight, bn_bias, eps, radius,
            use_envelope, e_out, (float*)e_out_t, (float*)gn_t, m);
    } else {
        CN_DISPATCH_PREC(prec, {
            edge_gate_aggregate_kernel<T, float, true><<<ceil_div(num_nodes, npb), 256, 0, (cudaStream_t)stream>>>(
                (const T*)g_t, (const T*)s_t, e, dist, row_ptr, num_nodes, D, bn_mean, bn_var, bn_weight, bn_bias, eps, radius,
                use_envelope, e_out, (T*)e_out_t, (T*)gn_t, m);
        });
    }
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_node_update(const float* m, const float* x, int32_t num_nodes, int32_t D, const float* bn_mean,
                        const float* bn_var, const float* bn_weight, const float* bn_bias, float eps, float* x_out,
                        void* x_out_t, int32_t prec, cartnet_stream_t stream) {
    CN_CHECK_ARG(m && x && bn_mean && bn_var && x_out && D % 4 == 0, "node_update: bad arguments");
    if (num_nodes <= 0) return 0;
    const int64_t total4 = (int64_t)num_nodes * D / 4;
    if (prec == CARTNET_PREC_FP32) {
        node_update_kernel<float, double><<<(unsigned)ceil_div64(total4, 256), 256, 0, (cudaStream_t)stream>>>(
            m, x, total4, D, bn_mean, bn_var, bn_weight, bn_bias, eps, x_out, (float*)x_out_t);
    } else {
        CN_DISPATCH_PREC(prec, {
            node_update_kernel<T, float><<<(unsigned)ceil_div64(total4, 256), 256, 0, (cudaStream_t)stream>>>(
                m, x, total4, D, bn_mean, bn_var, bn_weight, bn_bias, eps, x_out, (T*)x_out_t);
        });
    }
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_node_update_bwd_reduce(const float* dx_out, const float* m, int32_t num_nodes, int32_t D,
                                   const float* bn_mean, const float* bn_var, const float* bn_weight,
                                   const float* bn_bias, float eps, float* sums, double* partial,
                                   cartnet_stream_t stream) {
    CN_CHECK_ARG(dx_out && m && bn_mean && bn_var && sums && partial && num_nodes > 0, "node_update_bwd_reduce: bad arguments");
    CN_CHECK_ARG(colreduce_shape_ok(D), "node_update_bwd_reduce: unsupported D=%d", D);
    NodeBwdF f{dx_out, m, D, bn_mean, bn_var, bn_weight, bn_bias, eps};
    return run_colreduce(f, (int64_t)num_nodes, D, partial, FIN_SUMS, sums, nullptr, nullptr, nullptr, 0.f, 2 * D,
                         (cudaStream_t)stream);
}

int cartnet_node_update_bwd_apply(const float* dx_out, const float* m, int32_t num_nodes, int32_t D,
                                  const float* bn_mean, const float* bn_var, const float* bn_weight,
                                  const float* bn_bias, float eps, const float* sums, int32_t training, float* dm,
                                  cartnet_stream_t stream) {
    CN_CHECK_ARG(dx_out && m && bn_mean && bn_var && dm && D % 4 == 0, "node_update_bwd_apply: bad arguments");
    CN_CHECK_ARG(!training || sums, "node_update_bwd_apply: sums required in training mode");
    if (num_nodes <= 0) return 0;
    const int64_t total4 = (int64_t)num_nodes * D / 4;
    node_bwd_apply_kernel<<<(unsigned)ceil_div64(total4, 256), 256, 0, (cudaStream_t)stream>>>(
        dx_out, m, total4, D, num_nodes, bn_mean, bn_var, bn_weight, bn_bias, eps, sums, training, dm);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_edge_gate_bwd_reduce(const void* gn_t, const void* s_t, const float* dist, const int32_t* dst32,
                                 const float* de_out, const float* dm, int64_t num_edges, int32_t D,
                                 const float* bn_weight, const float* bn_bias, float radius, int32_t use_envelope,
                                 void* ds_t, void* dghat_t, int32_t prec, float* sums, double* partial,
                                 const float* g_mean, const float* g_var, float eps, cartnet_stream_t stream) {
    CN_CHECK_ARG(gn_t && s_t && dist && dst32 && dm && ds_t && dghat_t && sums && partial && num_edges > 0,
                 "edge_gate_bwd_reduce: bad arguments");
    CN_CHECK_ARG(colreduce_shape_ok(D), "edge_gate_bwd_reduce: unsupported D=%d", D);
    cudaStream_t st = (cudaStream_t)stream;
    if (prec == CARTNET_PREC_FP32) {
        EdgeBwdF<float, false> f{(const float*)gn_t, (const float*)s_t, dist, dst32, de_out, dm, D, bn_weight, bn_bias, radius,
                                 use_envelope, (float*)ds_t, (float*)dghat_t, g_mean, g_var, eps};
        return run_colreduce(f, num_edges, D, partial, FIN_SUMS, sums, nullptr, nullptr, nullptr, 0.f, 3 * D, st);
    }
    CN_DISPATCH_PREC(prec, {
        EdgeBwdF<T, true> f{(const T*)gn_t, (const T*)s_t, dist, dst32, de_out, dm, D, bn_weight, bn_bias, radius, use_envelope,
                            (T*)ds_t, (T*)dghat_t, g_mean, g_var, eps};
        return run_colreduce(f, num_edges, D, partial, FIN_SUMS, sums, nullptr, nullptr, nullptr, 0.f, 3 * D, st);
    });
    return 0;
}

int cartnet_edge_gate_bwd_apply(const void* gn_t, const void* dghat_t, int64_t num_edges, int32_t D, const float* bn_var,
                                const float* bn_weight, float eps, const float* sums, int32_t training, void* dg_t,
                                int32_t prec, const float* g_mean, int32_t input_is_g, cartnet_stream_t stream) {
    CN_CHECK_ARG(gn_t && dghat_t && bn_var && dg_t && row_shape_ok(D), "edge_gate_bwd_apply: bad arguments");
    CN_CHECK_ARG(!training || sums, "edge_gate_bwd_apply: sums required in training mode");
    if (num_edges <= 0) return 0;
    CN_DISPATCH_PREC(prec, {
        constexpr int V = Vec16<T>::N;
        CN_CHECK_ARG(D % V == 0 && 256 % (D / V) == 0, "edge_gate_bwd_apply: unsupported D=%d", D);
        const int lanes = 256 / (D / V);
        // contiguous row ranges, a multiple of one unrolled batch (4 x lanes rows); 3 blocks per SM = one resident wave at 66 registers
        int64_t per = ceil_div64(ceil_div64(num_edges, (int64_t)3 * kNumSMs), (int64_t)4 * lanes) * 4 * lanes;
        const int64_t blocks = ceil_div64(num_edges, per);
        edge_bwd_apply_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            (const T*)gn_t, (const T*)dghat_t, num_edges, D, bn_var, bn_weight, eps, sums, training, (T*)dg_t, per, g_mean, input_is_g);
    });
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_segment_sum(const void* x, int64_t ldx, const int32_t* ptr, const int32_t* perm, int32_t num_nodes,
                        int32_t C, void* out, int64_t ldo, int32_t out_is_t, int32_t prec, cartnet_stream_t stream) {
    CN_CHECK_ARG(ptr && out && (x || num_nodes == 0), "segment_sum: null pointer");
    CN_CHECK_ARG(row_shape_ok(C) && ldx % 4 == 0 && ldo % 4 == 0, "segment_sum: unsupported C=%d", C);
    if (num_nodes <= 0) return 0;
    const int npb = 256 / (C / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (!perm && (int64_t)num_nodes * (C / 4) <= 16384) {      // too few thread columns to fill the GPU: split the rows instead
        const dim3 grid((unsigned)num_nodes, (unsigned)ceil_div(C, 128));
        CN_DISPATCH_PREC(prec, {
            if (out_is_t)
                segment_sum_wide_kernel<T, T><<<grid, 256, 0, st>>>((const T*)x, ldx, ptr, C, (T*)out, ldo);
            else
                segment_sum_wide_kernel<T, float><<<grid, 256, 0, st>>>((const T*)x, ldx, ptr, C, (float*)out, ldo);
        });
        CN_LAUNCH_CHECK();
        return 0;
    }
    CN_DISPATCH_PREC(prec, {
        if (out_is_t)
            segment_sum_kernel<T, T><<<ceil_div(num_nodes, npb), 256, 0, st>>>((const T*)x, ldx, ptr, perm, num_nodes, C, (T*)out, ldo);
        else
            segment_sum_kernel<T, float><<<ceil_div(num_nodes, npb), 256, 0, st>>>((const T*)x, ldx, ptr, perm, num_nodes, C, (float*)out, ldo);
    });
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_segment_sum_pair(const void* x, int64_t ldx, const int32_t* row_ptr, const int32_t* col_ptr,
                             const int32_t* perm_src, int32_t num_nodes, int32_t C, void* out, int64_t ldo,
                             int32_t out_is_t, int32_t prec, cartnet_stream_t stream) {
    CN_CHECK_ARG(row_ptr && col_ptr && perm_src && out && (x || num_nodes == 0), "segment_sum_pair: null pointer");
    CN_CHECK_ARG(row_shape_ok(C) && ldx % 4 == 0 && ldo % 4 == 0 && ldo >= 2 * (int64_t)C, "segment_sum_pair: unsupported C=%d", C);
    if (num_nodes <= 0) return 0;
    const int npb = 256 / (C / 4);
    cudaStream_t st = (cudaStream_t)stream;
    static const int lanes_on = getenv("CARTNET_SEGSUM_LANES") ? atoi(getenv("CARTNET_SEGSUM_LANES")) : 1;      // 0: one thread column per node (A/B)
    if (lanes_on) {
        const int tpr = C / 4;
        const int RL = tpr * 8 <= 1024 ? 8 : 1024 / tpr;                 // row lanes per node
        const int nodes_pb = 1024 / (tpr * RL);                          // 1024-thread blocks
        const size_t smem = 1024 * sizeof(float4);
        CN_DISPATCH_PREC(prec, {
            if (out_is_t)
                segment_sum_pair_lanes_kernel<T, T><<<ceil_div(num_nodes, nodes_pb), 1024, smem, st>>>((const T*)x, ldx, row_ptr, col_ptr, perm_src, num_nodes, C, RL, (T*)out, ldo);
            else
                segment_sum_pair_lanes_kernel<T, float><<<ceil_div(num_nodes, nodes_pb), 1024, smem, st>>>((const T*)x, ldx, row_ptr, col_ptr, perm_src, num_nodes, C, RL, (float*)out, ldo);
        });
        CN_LAUNCH_CHECK();
        return 0;
    }
    CN_DISPATCH_PREC(prec, {
        if (out_is_t)
            segment_sum_pair_kernel<T, T><<<ceil_div(num_nodes, npb), 256, 0, st>>>((const T*)x, ldx, row_ptr, col_ptr, perm_src, num_nodes, C, (T*)out, ldo);
        else
            segment_sum_pair_kernel<T, float><<<ceil_div(num_nodes, npb), 256, 0, st>>>((const T*)x, ldx, row_ptr, col_ptr, perm_src, num_nodes, C, (float*)out, ldo);
    });
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_dsilu_mul(const float* dy, int64_t ld_dy, const void* z, int64_t ldz, void* y, int64_t ldy, int64_t rows,
                      int32_t C, int32_t prec, float* colsum, double* partial, cartnet_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (rows <= 0) {
        if (colsum) CN_CUDA(cudaMemsetAsync(colsum, 0, (size_t)C * sizeof(float), st));
        return 0;
    }
    CN_CHECK_ARG(dy && z && y && C % 4 == 0 && ld_dy % 4 == 0 && ldz % 4 == 0 && ldy % 4 == 0, "dsilu_mul: bad arguments");
    if (colsum) {      // one pass: the column sums of y ride along (fp64 partials, fixed order)
        CN_CHECK_ARG(partial && colreduce_shape_ok(C), "dsilu_mul: column sums need a workspace and C/4 a power of two <= 256 (C=%d)", C);
        CN_DISPATCH_PREC(prec, {
            DsiluMulF<T> f{dy, ld_dy, (const typename ZOf<T>::type*)z, ldz, (T*)y, ldy};
            return run_colreduce(f, rows, C, partial, FIN_SUMS, colsum, nullptr, nullptr, nullptr, 0.f, C, st);
        });
    }
    const int64_t total = rows * (C / 4);
    CN_DISPATCH_PREC(prec, {
        dsilu_mul_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(dy, ld_dy, (const typename ZOf<T>::type*)z, ldz, (T*)y, ldy, rows, C);
    });
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_cast_rows(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int32_t C, int32_t prec,
                      cartnet_stream_t stream) {
    if (rows <= 0) return 0;
    CN_CHECK_ARG(src && dst && C % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "cast_rows: bad arguments");
    if (rows <= 0) return 0;
    const int64_t total = rows * (C / 4);
    CN_DISPATCH_PREC(prec, {
        cast_rows_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(src, lds, (T*)dst, ldd, rows, C);
    });
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_uncast_rows(const void* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int32_t C, int32_t prec,
                        cartnet_stream_t stream) {
    if (rows <= 0) return 0;
    CN_CHECK_ARG(src && dst && C % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "uncast_rows: bad arguments");
    const int64_t total = rows * (C / 4);
    CN_DISPATCH_PREC(prec, {
        uncast_rows_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)src, lds, dst, ldd, rows, C);
    });
    CN_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
