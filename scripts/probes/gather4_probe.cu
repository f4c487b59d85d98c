// Probe (not product code): semantics and throughput of TMA tile::gather4 on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/build/gather4_probe scripts/probes/gather4_probe.cu -lcuda
// Table P [R, C] bf16 row-major; each gather4 fetches 4 rows x 64 columns (128 B) into shared memory (SWIZZLE_128B).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void gather4(void* dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

// correctness: one CTA, gathers 4 rows, dumps the 512 bytes that landed
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int r0, int r1, int r2, int r3, uint16_t* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < 2048 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xFFFF;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 512);
        gather4(smem, &tm, c0, r0, r1, r2, r3, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 2048 / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

// throughput: every CTA gathers `iters` x 32 gather4 (= 128 rows x 128 B = 16 KB per batch) with random rows
__global__ void __launch_bounds__(128, 1) bw_kernel(const __grid_constant__ CUtensorMap tm, const int* __restrict__ rows, int iters, int ncolblk, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[4];
    if (threadIdx.x == 0) { for (int s = 0; s < 4; ++s) mbar_init(&bar[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int* myrows = rows + (size_t)blockIdx.x * iters * 128;
    if (threadIdx.x < 32) {
        // lane l issues gather4 #l of each batch; 4 batches in flight
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            const int s = it & 3;
            if (it >= 4) { mbar_wait(&bar[s], phase); if (s == 3) phase ^= 1; }
            if (threadIdx.x == 0) mbar_expect_tx(&bar[s], 16384);
            __syncwarp();
            const int* r = myrows + it * 128 + threadIdx.x * 4;
            gather4(smem + s * 16384 + threadIdx.x * 512, &tm, (it % ncolblk) * 64, r[0], r[1], r[2], r[3], &bar[s]);
        }
        for (int s = 0; s < 4; ++s) mbar_wait(&bar[s], (uint32_t)(((iters + 3 - s) / 4 - 1) & 1));      // last use of each stage
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = reinterpret_cast<float*>(smem)[5];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int R = 50000, C = 256;      // 25.6 MB: L2 resident
    std::vector<uint16_t> h((size_t)R * C);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = (uint16_t)((r * 7 + c) & 0xFFFF);
    uint16_t* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    uint16_t* out; CK(cudaMalloc(&out, 2048));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
    for (int boxrows : {1}) {   // boxDim[1] = 4 encodes but the instruction then faults (illegal instruction): the box row count must be 1
        for (int swz = 0; swz < 2; ++swz) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
            cuuint64_t strides[1] = {(cuuint64_t)C * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            printf("boxrows=%d swizzle=%d encode rc=%d\n", boxrows, swz, (int)r);
            if (r != CUDA_SUCCESS) continue;
            CK(cudaMemset(out, 0, 2048));
            const int rows[4] = {5, 40001, 17, 12345};
            probe_kernel<<<1, 128, 8192>>>(tm, 64, rows[0], rows[1], rows[2], rows[3], out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<uint16_t> o(1024); CK(cudaMemcpy(o.data(), out, 2048, cudaMemcpyDeviceToHost));
            // expected: smem row i (128 B) = table row rows[i], cols 64..127; swizzle: 16-byte chunk j of smem row i at chunk j ^ (i & 7)
            int ok = 1;
            for (int i = 0; i < 4 && ok; ++i) for (int c = 0; c < 64; ++c) {
                const int chunk = c / 8, phys = swz ? (chunk ^ (i & 7)) : chunk;
                const uint16_t got = o[i * 64 + phys * 8 + (c & 7)], want = (uint16_t)((rows[i] * 7 + 64 + c) & 0xFFFF);
                if (got != want) { printf("  MISMATCH smem row %d col %d: got %u want %u\n", i, c, got, want); ok = 0; break; }
            }
            printf("  gather4 rows land as 4 consecutive 128-byte smem rows%s: %s; bytes beyond 512 untouched: %s\n", swz ? " (128B swizzle)" : "", ok ? "YES" : "NO",
                   o[256] == 0xFFFF ? "yes" : "no");
        }
    }
    // throughput with the variant box {64,1} + SWIZZLE_128B (if it encoded)
    {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, 1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) {
            const int iters = 2000, grid = 148;
            std::vector<int> rows((size_t)grid * iters * 128);
            srand(1);
            for (auto& v : rows) v = rand() % R;
            int* drows; CK(cudaMalloc(&drows, rows.size() * 4)); CK(cudaMemcpy(drows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice));
            float* sink; CK(cudaMalloc(&sink, grid * 4));
            CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 1024));
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                bw_kernel<<<grid, 128, 4 * 16384 + 1024>>>(tm, drows, iters, C / 64, sink);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("bw kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = (double)grid * iters * 16384;
                printf("gather4 throughput: %.3f ms, %.1f GB/s, %.1f M gather4/s per SM\n", ms, bytes / ms * 1e-6, (double)iters * 32 / ms * 1e-3);
            }
        }
    }
    return 0;
}
