// Cholesky head tail: the second Linear of the head MLP (Dh -> 6), the softplus diagonal, the
// upper-triangular assembly and U = L^T L of /root/reference/models/cartnet.py:293-305 in ONE launch per direction,
// instead of ~12 eager launches forward and ~25 backward. Node-side and tiny (Dh = 128, ~7k non-H atoms at ADP-64), so
// the design goal is launch count and determinism, not bandwidth: one warp per atom, lanes over the Dh hidden channels,
// the [6, Dh] weight in registers; the weight-gradient reduction over atoms is fixed-order (per warp -> per block in
// shared memory -> fp64 over blocks in a second launch), no atomics.
#include "common.cuh"

namespace cartnet {

constexpr int HEAD_THREADS = 256;           // 8 warps = 8 atoms in flight per block
constexpr int HEAD_MAX_CHUNKS = 4;          // Dh <= 4 * 128

__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus defaults
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// p = W1 h + b1 ; d = softplus(p[0:3]) ; L = [[d0,p3,p4],[0,d1,p5],[0,0,d2]] ; U = L^T L          (cartnet.py:294-303)
template <int CH>
__global__ void __launch_bounds__(HEAD_THREADS)
cholesky_head_fwd_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ W1, const float* __restrict__ b1,
                         int n, int Dh, float* __restrict__ p6, float* __restrict__ U) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (HEAD_THREADS / 32) + (threadIdx.x >> 5), nwarps = gridDim.x * (HEAD_THREADS / 32);
    float4 w[6][CH];
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int c = j * 128 + lane * 4;
            w[k][j] = c < Dh ? *reinterpret_cast<const float4*>(W1 + (int64_t)k * Dh + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    for (int r = warp; r < n; r += nwarps) {
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int c = j * 128 + lane * 4;
            if (c < Dh) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(h + (int64_t)r * ldh + c));
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    acc[k] = fmaf(x.x, w[k][j].x, fmaf(x.y, w[k][j].y, fmaf(x.z, w[k][j].z, fmaf(x.w, w[k][j].w, acc[k]))));
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
            float p[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) { p[k] = acc[k] + b1[k]; p6[(int64_t)r * 6 + k] = p[k]; }
            const float d0 = softplusf_(p[0]), d1 = softplusf_(p[1]), d2 = softplusf_(p[2]);
            const float L[3][3] = {{d0, p[3], p[4]}, {0.f, d1, p[5]}, {0.f, 0.f, d2}};
            float* u = U + (int64_t)r * 9;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) u[i * 3 + j] = L[0][i] * L[0][j] + L[1][i] * L[1][j] + L[2][i] * L[2][j];
        }
    }
}

// dL = L (dU + dU^T) ; d(p3,p4,p5) = dL[0][1], dL[0][2], dL[1][2] ; d(p_i) = dL[i][i] sigmoid(p_i) (i < 3)
// dh = W1^T dp ; dW1 = sum_atoms dp (x) h ; db1 = sum_atoms dp. Block partials: [gridDim.x][6 * Dh + 8] fp32.
template <int CH>
__global__ void __launch_bounds__(HEAD_THREADS)
cholesky_head_bwd_kernel(const float* __restrict__ dU, const float* __restrict__ h, int64_t ldh, const float* __restrict__ p6,
                         const float* __restrict__ W1, int n, int Dh, float* __restrict__ dh, int64_t lddh,
                         float* __restrict__ partial) {
    extern __shared__ float sm[];                       // [8 warps][6 * Dh + 8]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * (HEAD_THREADS / 32) + wib, nwarps = gridDim.x * (HEAD_THREADS / 32);
    const int stride = 6 * Dh + 8;
    float4 w[6][CH], gw[6][CH];
    float gb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int c = j * 128 + lane * 4;
            w[k][j] = c < Dh ? *reinterpret_cast<const float4*>(W1 + (int64_t)k * Dh + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            gw[k][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    for (int r = warp; r < n; r += nwarps) {
        // every lane recomputes the 6 scalars (a few dozen flops) instead of broadcasting them
        float p[6], g[9];
#pragma unroll
        for (int k = 0; k < 6; ++k) p[k] = __ldg(p6 + (int64_t)r * 6 + k);
#pragma unroll
        for (int k = 0; k < 9; ++k) g[k] = __ldg(dU + (int64_t)r * 9 + k);
        const float d0 = softplusf_(p[0]), d1 = softplusf_(p[1]), d2 = softplusf_(p[2]);
        const float L[3][3] = {{d0, p[3], p[4]}, {0.f, d1, p[5]}, {0.f, 0.f, d2}};
        float S[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) S[i][j] = g[i * 3 + j] + g[j * 3 + i];
        float dL[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) dL[i][j] = L[i][0] * S[0][j] + L[i][1] * S[1][j] + L[i][2] * S[2][j];
        float dp[6];
#pragma unroll
        for (int i = 0; i < 3; ++i) dp[i] = dL[i][i] * (p[i] > 20.f ? 1.f : 1.f / (1.f + expf(-p[i])));
        dp[3] = dL[0][1]; dp[4] = dL[0][2]; dp[5] = dL[1][2];
#pragma unroll
        for (int k = 0; k < 6; ++k) gb[k] += dp[k];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int c = j * 128 + lane * 4;
            if (c < Dh) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(h + (int64_t)r * ldh + c));
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    o.x = fmaf(dp[k], w[k][j].x, o.x); o.y = fmaf(dp[k], w[k][j].y, o.y);
                    o.z = fmaf(dp[k], w[k][j].z, o.z); o.w = fmaf(dp[k], w[k][j].w, o.w);
                    gw[k][j].x = fmaf(dp[k], x.x, gw[k][j].x); gw[k][j].y = fmaf(dp[k], x.y, gw[k][j].y);
                    gw[k][j].z = fmaf(dp[k], x.z, gw[k][j].z); gw[k][j].w = fmaf(dp[k], x.w, gw[k][j].w);
                }
                *reinterpret_cast<float4*>(dh + (int64_t)r * lddh + c) = o;
            }
        }
    }
    // per-warp partials -> shared memory -> fixed-order sum over the 8 warps -> one partial row per block
    float* mine = sm + wib * stride;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int c = j * 128 + lane * 4;
            if (c < Dh) *reinterpret_cast<float4*>(mine + k * Dh + c) = gw[k][j];
        }
        if (lane == 0) mine[6 * Dh + k] = gb[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 6 * Dh + 6; i += HEAD_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < HEAD_THREADS / 32; ++q) s += sm[q * stride + i];
        partial[(int64_t)blockIdx.x * stride + i] = s;
    }
}

__global__ void cholesky_head_final_kernel(const float* __restrict__ partial, int blocks, int Dh, float* __restrict__ dW1,
                                           float* __restrict__ db1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = 6 * Dh + 8;
    if (i >= 6 * Dh + 6) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += (double)partial[(int64_t)b * stride + i];
    if (i < 6 * Dh) dW1[i] = (float)s;
    else db1[i - 6 * Dh] = (float)s;
}

static int head_blocks(int n) {
    int b = ceil_div(n, HEAD_THREADS / 32);      // few block partials: the final pass walks them serially per output
    return b < 1 ? 1 : (b > kNumSMs / 2 ? kNumSMs / 2 : b);
}

// ---------------------------------------------------------------------------------------------------------------
// L1 / MSE loss pair of /root/reference/train/metrics.py:15-28 (nn.L1Loss + nn.MSELoss, reduction = "mean") in ONE
// launch: a single block walks the elements in a fixed order (n ~ 60 k values at ADP-64: node-side, latency-bound), fp64
// running sums, both means written to out[0:2]. The backward is one elementwise launch:
// dpred = dMAE * sign(pred - true) / n + dMSE * 2 (pred - true) / n   (what autograd derives for the two modules).
constexpr int LOSS_THREADS = 1024;
__global__ void __launch_bounds__(LOSS_THREADS)
loss_l1_mse_kernel(const float* __restrict__ pred, const float* __restrict__ truth, int64_t n, float* __restrict__ out) {
    __shared__ double sm[2][LOSS_THREADS / 32];
    double a = 0.0, q = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += LOSS_THREADS) {
        const float d = pred[i] - truth[i];
        a += (double)fabsf(d);
        q += (double)d * (double)d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = a; sm[1][threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sq = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) { sa += sm[0][w]; sq += sm[1][w]; }
        out[0] = (float)(sa / (double)n);
        out[1] = (float)(sq / (double)n);
    }
}

__global__ void loss_l1_mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ truth, int64_t n,
                                       const float* __restrict__ dmae, const float* __restrict__ dmse, float* __restrict__ dpred) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = pred[i] - truth[i];
    const float ga = dmae ? dmae[0] : 0.f, gq = dmse ? dmse[0] : 0.f;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    dpred[i] = (ga * sgn + gq * 2.0f * d) / (float)n;
}

}  // namespace cartnet

using namespace cartnet;

extern "C" {

int64_t cartnet_cholesky_head_workspace(int32_t n, int32_t Dh) {
    return (int64_t)head_blocks(n) * (6 * (int64_t)Dh + 8) * (int64_t)sizeof(float);
}

int cartnet_cholesky_head_fwd(const float* h, int64_t ldh, const float* W1, const float* b1, int32_t n, int32_t Dh, float* p6,
                              float* U, cartnet_stream_t stream) {
    if (n <= 0) return 0;
    CN_CHECK_ARG(h && W1 && b1 && p6 && U, "cholesky_head_fwd: null pointer");
    CN_CHECK_ARG(Dh % 4 == 0 && Dh > 0 && Dh <= 128 * HEAD_MAX_CHUNKS && ldh % 4 == 0, "cholesky_head_fwd: unsupported Dh=%d", Dh);
    const int ch = ceil_div(Dh, 128), blocks = head_blocks(n);
    cudaStream_t st = (cudaStream_t)stream;
#define CN_HEAD_FWD(C_) \
    case C_: cholesky_head_fwd_kernel<C_><<<blocks, HEAD_THREADS, 0, st>>>(h, ldh, W1, b1, n, Dh, p6, U); break;
    switch (ch) {
        CN_HEAD_FWD(1) CN_HEAD_FWD(2) CN_HEAD_FWD(3) CN_HEAD_FWD(4)
        default: set_error("cholesky_head_fwd: Dh=%d > 512 not instantiated", Dh); return 2;
    }
#undef CN_HEAD_FWD
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_cholesky_head_bwd(const float* dU, const float* h, int64_t ldh, const float* p6, const float* W1, int32_t n, int32_t Dh,
                              float* dh, int64_t lddh, float* dW1, float* db1, float* partial, cartnet_stream_t stream) {
    CN_CHECK_ARG(dW1 && db1, "cholesky_head_bwd: null gradient output");
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) {
        CN_CUDA(cudaMemsetAsync(dW1, 0, 6 * (size_t)Dh * sizeof(float), st));
        CN_CUDA(cudaMemsetAsync(db1, 0, 6 * sizeof(float), st));
        return 0;
    }
    CN_CHECK_ARG(dU && h && p6 && W1 && dh && partial, "cholesky_head_bwd: null pointer");
    CN_CHECK_ARG(Dh % 4 == 0 && Dh > 0 && Dh <= 128 * HEAD_MAX_CHUNKS && ldh % 4 == 0 && lddh % 4 == 0, "cholesky_head_bwd: unsupported Dh=%d", Dh);
    const int ch = ceil_div(Dh, 128), blocks = head_blocks(n);
    const size_t smem = (size_t)(HEAD_THREADS / 32) * (6 * (size_t)Dh + 8) * sizeof(float);
#define CN_HEAD_BWD(C_)                                                                                                         \
    case C_:                                                                                                                    \
        if (smem > 48 * 1024) CN_CUDA(cudaFuncSetAttribute(cholesky_head_bwd_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cholesky_head_bwd_kernel<C_><<<blocks, HEAD_THREADS, smem, st>>>(dU, h, ldh, p6, W1, n, Dh, dh, lddh, partial);         \
        break;
    switch (ch) {
        CN_HEAD_BWD(1) CN_HEAD_BWD(2) CN_HEAD_BWD(3) CN_HEAD_BWD(4)
        default: set_error("cholesky_head_bwd: Dh=%d > 512 not instantiated", Dh); return 2;
    }
#undef CN_HEAD_BWD
    CN_LAUNCH_CHECK();
    cholesky_head_final_kernel<<<ceil_div(6 * Dh + 6, 128), 128, 0, st>>>(partial, blocks, Dh, dW1, db1);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_loss_l1_mse(const float* pred, const float* truth, int64_t n, float* out2, cartnet_stream_t stream) {
    CN_CHECK_ARG(pred && truth && out2 && n > 0, "loss_l1_mse: bad arguments");
    loss_l1_mse_kernel<<<1, LOSS_THREADS, 0, (cudaStream_t)stream>>>(pred, truth, n, out2);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_loss_l1_mse_bwd(const float* pred, const float* truth, int64_t n, const float* dmae, const float* dmse,
                            float* dpred, cartnet_stream_t stream) {
    CN_CHECK_ARG(pred && truth && dpred && n > 0, "loss_l1_mse_bwd: bad arguments");
    loss_l1_mse_bwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(pred, truth, n, dmae, dmse, dpred);
    CN_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
