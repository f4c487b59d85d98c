// placeholder until the tcgen05 kernels land
#include "common.cuh"
namespace cartnet {
int gemm_tc_nt(const cartnet_gemm_t& d, cudaStream_t st) { set_error("bf16 tcgen05 GEMM not built yet"); return 3; }
int gemm_tc_tn(int prec, int M, int N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, float* ws, int64_t ws_bytes, cudaStream_t st) { set_error("bf16 tcgen05 GEMM not built yet"); return 3; }
int64_t gemm_tc_tn_workspace(int prec, int M, int N, int64_t K) { return 16; }
}
