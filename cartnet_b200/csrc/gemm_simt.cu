// fp32 SIMT GEMMs of the 1e-5 parity path (CARTNET_PREC_FP32) and the C-ABI dispatch for both
// precisions. 128x128x16 tiles, 256 threads, 8x8 register micro-tile, register-staged prefetch.
//   MODE 0 ("NT"): C[M,N] = A[M,K] * B[N,K]^T, fused epilogue        (forward / dgrad)
//   MODE 1 ("TN"): C[M,N] = sum_k A[k,M]^T B[k,N], split-K partials   (wgrad, K = edges)
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace cartnet {

int gemm_tc_nt(const cartnet_gemm_t& d, cudaStream_t st, double* stats = nullptr, int* stats_blocks = nullptr);   // bf16 or tf32 by d.prec                                  // gemm_tc.cu
int gemm_tc_tn(int prec, int M, int N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb, const TnDst& C,
               int64_t ldc, float* ws, int64_t ws_bytes, cudaStream_t st);                 // gemm_tc.cu
int64_t gemm_tc_tn_workspace(int prec, int M, int N, int64_t K);

constexpr int BM = 128, BN = 128, BK = 16, PADM = 4;

__device__ __forceinline__ float4 ld4_guard(const float* p, bool ok) {
    return ok ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int MODE>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
             int64_t ldb, EpiParams<float> epi, float* __restrict__ partial, int n_tiles, int64_t k_chunk) {
    __shared__ __align__(16) float As[BK][BM + PADM];
    __shared__ __align__(16) float Bs[BK][BN + PADM];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
    const int64_t kb = (int64_t)blockIdx.y * k_chunk;
    const int64_t ke = (kb + k_chunk < K) ? kb + k_chunk : K;
    const int ty = tid >> 4, tx = tid & 15;

    // Parity path: every 16-wide K block is accumulated in fp32 (chain length 16) and the running total is
    // carried in fp64, so the rounding error does not grow with K (the reference's MKL sgemm also keeps many
    // short partial sums). Costs ~15% and registers; this is the 1e-5 path, not the fast one.
    double dacc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dacc[i][j] = 0.0;

    float4 ra[2], rb[2];
    auto load_tile = [&](int64_t k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = tid + 256 * h;
            if (MODE == 0) {   // k-contiguous operands: float4 along k
                const int row = idx >> 2, kq = (idx & 3) * 4;
                const int64_t k = k0 + kq;
                ra[h] = ld4_guard(A + (int64_t)(m0 + row) * lda + k, (m0 + row) < M && k < ke);
                rb[h] = ld4_guard(B + (int64_t)(n0 + row) * ldb + k, (n0 + row) < N && k < ke);
            } else {           // row index = k, columns contiguous: float4 along m / n
                const int kk = idx >> 5, c4 = (idx & 31) * 4;
                const int64_t k = k0 + kk;
                ra[h] = ld4_guard(A + k * lda + m0 + c4, k < ke && (m0 + c4) < M);
                rb[h] = ld4_guard(B + k * ldb + n0 + c4, k < ke && (n0 + c4) < N);
            }
        }
    };
    auto store_tile = [&]() {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = tid + 256 * h;
            if (MODE == 0) {
                const int row = idx >> 2, kq = (idx & 3) * 4;
                As[kq + 0][row] = ra[h].x; As[kq + 1][row] = ra[h].y; As[kq + 2][row] = ra[h].z; As[kq + 3][row] = ra[h].w;
                Bs[kq + 0][row] = rb[h].x; Bs[kq + 1][row] = rb[h].y; Bs[kq + 2][row] = rb[h].z; Bs[kq + 3][row] = rb[h].w;
            } else {
                const int kk = idx >> 5, c4 = (idx & 31) * 4;
                *reinterpret_cast<float4*>(&As[kk][c4]) = ra[h];
                *reinterpret_cast<float4*>(&Bs[kk][c4]) = rb[h];
            }
        }
    };

    if (kb < ke) load_tile(kb);
    for (int64_t k0 = kb; k0 < ke; k0 += BK) {
        store_tile();
        __syncthreads();
        if (k0 + BK < ke) load_tile(k0 + BK);
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        // unrolled by 4, not fully: 16 x (64 FFMA + 4 LDS) = 18 KB of code per block iteration made the 8 warps of the SM
        // stall on instruction fetch (ncu: stall_no_instruction 0.9-1.5 cycles per issue)
#pragma unroll 4
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dacc[i][j] += (double)acc[i][j];
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= M) continue;
        EpiRow<float> er;
        if (MODE == 0) er = epi_row(epi, row);
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int col = n0 + jh * 64 + tx * 4;
            if (col >= N) continue;   // N % 4 == 0 is required, so a float4 never straddles N
            if (MODE == 0) {
                epi_apply4_f64(epi, er, (int64_t)row, col, &dacc[i][jh * 4]);
            } else {
                const float4 v = make_float4((float)dacc[i][jh * 4 + 0], (float)dacc[i][jh * 4 + 1], (float)dacc[i][jh * 4 + 2], (float)dacc[i][jh * 4 + 3]);
                *reinterpret_cast<float4*>(partial + ((int64_t)blockIdx.y * M + row) * N + col) = v;
            }
        }
    }
}

// C_b[m', n] = sum_z partial[z][m][n]  (fixed order -> deterministic); rows are split into equal blocks b = m / rows_per_blk
// with their own destination base (the wgrad of a row-packed weight goes straight into the reference's layout).
// A block owns 32 consecutive float4 outputs; its 8 warps each sum a contiguous range of the splits (loads of different
// splits are independent, 148 of them in sequence per thread made this latency-bound), then warp 0 adds the 8 range
// sums in a fixed order.
constexpr int SKR_GROUPS = 8;
__global__ void __launch_bounds__(32 * SKR_GROUPS)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, TnDst dst, int64_t ldc) {
    __shared__ float4 sm[SKR_GROUPS][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int64_t i = blockIdx.x * (int64_t)32 + lane;                   // over M*N/4
    const int64_t total4 = (int64_t)M * N / 4;
    const int64_t lin = i * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total4) {
        const int per = (splits + SKR_GROUPS - 1) / SKR_GROUPS;
        const int z0 = grp * per, z1 = (z0 + per) < splits ? (z0 + per) : splits;
        const float* p = partial + lin;
        const int64_t zs = (int64_t)M * N;
        int z = z0;
        for (; z + 4 <= z1; z += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p + (int64_t)z * zs));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p + (int64_t)(z + 1) * zs));
            const float4 c = __ldg(reinterpret_cast<const float4*>(p + (int64_t)(z + 2) * zs));
            const float4 d = __ldg(reinterpret_cast<const float4*>(p + (int64_t)(z + 3) * zs));
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
            s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
            s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
        }
        for (; z < z1; ++z) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p + (int64_t)z * zs));
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    sm[grp][lane] = s;
    __syncthreads();
    if (grp == 0 && i < total4) {
#pragma unroll
        for (int g = 1; g < SKR_GROUPS; ++g) {
            const float4 v = sm[g][lane];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        const int m = (int)(lin / N), n = (int)(lin % N);
        const int blk = m / dst.rows_per_blk;
        *reinterpret_cast<float4*>(dst.c[blk] + (int64_t)(m - blk * dst.rows_per_blk) * ldc + n) = s;
    }
}

static int simt_tn_splits(int M, int N, int64_t K) {
    const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
    int64_t by_sm = (2 * kNumSMs + tiles - 1) / tiles;
    int64_t by_k = ceil_div64(K, 8 * BK);
    int64_t s = by_sm < by_k ? by_sm : by_k;
    return (int)(s < 1 ? 1 : s);
}

int launch_splitk_reduce(const float* partial, int splits, int M, int N, const TnDst& dst, int64_t ldc, cudaStream_t st) {
    const int64_t total4 = (int64_t)M * N / 4;
    splitk_reduce_kernel<<<(unsigned)ceil_div64(total4, 32), 32 * SKR_GROUPS, 0, st>>>(partial, splits, M, N, dst, ldc);
    CN_LAUNCH_CHECK();
    return 0;
}

}  // namespace cartnet

using namespace cartnet;

extern "C" {

int cartnet_gemm(const cartnet_gemm_t* d, cartnet_stream_t stream) {
    CN_CHECK_ARG(d, "gemm: null descriptor");
    CN_CHECK_ARG(d->M >= 0 && d->N > 0 && d->K > 0, "gemm: bad shape M=%d N=%d K=%d", d->M, d->N, d->K);
    if (d->M == 0) return 0;           // empty graph: zero-sized tensors carry null data pointers
    CN_CHECK_ARG(d->A && d->B, "gemm: null operand");
    CN_CHECK_ARG(d->N % 4 == 0 && d->K % 4 == 0 && d->lda % 4 == 0 && d->ldb % 4 == 0, "gemm: N,K,lda,ldb must be multiples of 4");
    CN_CHECK_ARG(!(d->act == CARTNET_ACT_MUL_DSILU) || d->z_in, "gemm: ACT_MUL_DSILU needs z_in");
    CN_CHECK_ARG((!d->gather0 || d->gidx0) && (!d->gather1 || d->gidx1), "gemm: gather without index");
    CN_CHECK_ARG(d->out_f32 || d->out_t || d->z_out, "gemm: no output");
    if (d->M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->prec != CARTNET_PREC_FP32 && d->prec >= 0 && d->prec <= CARTNET_PREC_BF16X3) return gemm_tc_nt(*d, st);
    CN_CHECK_ARG(d->prec == CARTNET_PREC_FP32, "gemm: unknown prec %d", d->prec);
    const int n_tiles = ceil_div(d->N, BN);
    const int64_t tiles = (int64_t)ceil_div(d->M, BM) * n_tiles;
    sgemm_kernel<0><<<dim3((unsigned)tiles, 1), 256, 0, st>>>(d->M, d->N, d->K, (const float*)d->A, d->lda,
                                                             (const float*)d->B, d->ldb, make_epi<float>(*d), nullptr,
                                                             n_tiles, (int64_t)d->K);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_gemm_colstats(const cartnet_gemm_t* d, const float* shift, float* mean, float* var, float* running_mean,
                          float* running_var, float momentum, double* partial, cartnet_stream_t stream) {
    CN_CHECK_ARG(d && d->out_t && !d->out_f32 && !d->z_out && d->act == CARTNET_ACT_NONE && !d->gather0 && !d->gather1 && !d->resid,
                 "gemm_colstats: only the bias + T-output epilogue is supported");
    CN_CHECK_ARG(mean && var && partial && d->M > 0, "gemm_colstats: bad arguments");
    CN_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "gemm_colstats: running stats must come in pairs");
    cudaStream_t st = (cudaStream_t)stream;
    if (d->prec != CARTNET_PREC_FP32 && d->prec >= 0 && d->prec <= CARTNET_PREC_BF16X3) {
        CN_CHECK_ARG(d->A && d->B && d->bias && d->N > 0 && d->K > 0, "gemm_colstats: bad GEMM arguments");
        int blocks = 0;
        if (int rc = gemm_tc_nt(*d, st, partial, &blocks)) return rc;       // sums ride in the epilogue
        return launch_colstats_final(partial, blocks, d->N, d->M, shift, mean, var, running_mean, running_var, momentum, st);
    }
    // fp32 parity mode: SIMT GEMM, then the fp64 statistics pass over its output
    if (int rc = cartnet_gemm(d, stream)) return rc;
    return cartnet_colstats(d->out_t, 1, d->prec, d->M, d->N, d->ldt, shift, mean, var, running_mean, running_var, momentum, partial, stream);
}

int64_t cartnet_gemm_tn_workspace(int32_t prec, int32_t M, int32_t N, int64_t K) {
    if (prec != CARTNET_PREC_FP32 && prec >= 0 && prec <= CARTNET_PREC_BF16X3) return gemm_tc_tn_workspace(prec, M, N, K);
    return (int64_t)simt_tn_splits(M, N, K) * M * N * (int64_t)sizeof(float);
}

int cartnet_gemm_tn_blocks(int32_t prec, int32_t M, int32_t N, int64_t K, const void* A, int64_t lda, const void* B,
                           int64_t ldb, float* const* C_blocks, int32_t num_blocks, int64_t ldc, float* workspace,
                           int64_t workspace_bytes, cartnet_stream_t stream) {
    CN_CHECK_ARG(M > 0 && N > 0 && K >= 0 && A && B && C_blocks, "gemm_tn: bad arguments");
    CN_CHECK_ARG(num_blocks >= 1 && num_blocks <= 4 && M % num_blocks == 0, "gemm_tn: 1..4 equal row blocks (M=%d, blocks=%d)", M, num_blocks);
    CN_CHECK_ARG(M % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "gemm_tn: M,N,ld* must be multiples of 4");
    CN_CHECK_ARG(workspace && workspace_bytes >= cartnet_gemm_tn_workspace(prec, M, N, K), "gemm_tn: workspace too small");
    TnDst dst = {};
    dst.rows_per_blk = M / num_blocks;
    for (int b = 0; b < num_blocks; ++b) {
        CN_CHECK_ARG(C_blocks[b], "gemm_tn: null output block %d", b);
        dst.c[b] = C_blocks[b];
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (prec != CARTNET_PREC_FP32 && prec >= 0 && prec <= CARTNET_PREC_BF16X3)
        return gemm_tc_tn(prec, M, N, K, A, lda, B, ldb, dst, ldc, workspace, workspace_bytes, st);
    CN_CHECK_ARG(prec == CARTNET_PREC_FP32, "gemm_tn: unknown prec %d", prec);
    if (K <= 0) {
        for (int b = 0; b < num_blocks; ++b)
            CN_CUDA(cudaMemset2DAsync(dst.c[b], ldc * sizeof(float), 0, (size_t)N * sizeof(float), dst.rows_per_blk, st));
        return 0;
    }
    const int splits = simt_tn_splits(M, N, K);
    const int n_tiles = ceil_div(N, BN);
    const int tiles = ceil_div(M, BM) * n_tiles;
    int64_t k_chunk = ceil_div64(ceil_div64(K > 0 ? K : 1, splits), BK) * BK;
    EpiParams<float> none = {};
    sgemm_kernel<1><<<dim3((unsigned)tiles, (unsigned)splits), 256, 0, st>>>(M, N, K, (const float*)A, lda,
                                                                             (const float*)B, ldb, none, workspace,
                                                                             n_tiles, k_chunk);
    CN_LAUNCH_CHECK();
    return launch_splitk_reduce(workspace, splits, M, N, dst, ldc, st);
}

int cartnet_gemm_tn(int32_t prec, int32_t M, int32_t N, int64_t K, const void* A, int64_t lda, const void* B,
                    int64_t ldb, float* C, int64_t ldc, float* workspace, int64_t workspace_bytes,
                    cartnet_stream_t stream) {
    CN_CHECK_ARG(C, "gemm_tn: null output");
    float* blocks[1] = {C};
    return cartnet_gemm_tn_blocks(prec, M, N, K, A, lda, B, ldb, blocks, 1, ldc, workspace, workspace_bytes, stream);
}

}  // extern "C"
