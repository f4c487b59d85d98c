// tcgen05 / TMEM / TMA GEMMs for sm_100a (CARTNET_PREC_BF16: kind::f16 on bf16 operands,
// CARTNET_PREC_TF32: kind::tf32 on fp32 words that were rounded to tf32 where they were produced). Hand-written PTX; descriptor bit
// layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
//
// CARTNET_PREC_BF16X3: kind::f16 on hi|lo bf16 pairs, three MMAs per product (common.cuh::bf16p_t).
//
// NT kernel  C[M,N] = A[M,K] B[N,K]^T  (+ fused epilogue, gemm_epilogue.cuh)
//   * persistent, one CTA per SM; the CTA's <= 128 KB weight slice B[n0:n0+BN, 0:K] is loaded ONCE by TMA and
//     stays resident in shared memory (K-major, 128B swizzle), so per 128-row tile only A streams through a
//     2..6-stage TMA/mbarrier ring -- weights never re-cross L2->SM per tile;
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one thread, tcgen05.mma, M=128, N=BN),
//     warps 2..9 = epilogue: tcgen05.ld 32 columns at a time from one of TWO TMEM accumulators, so the
//     epilogue of tile i overlaps the MMAs of tile i+1; gathers / bias / SiLU / stores are fused there;
//   * edge-sized GEMMs of the 4-byte operand modes run as CTA PAIRS (cta_group::2, M = 256): each SM keeps half of the
//     pair's weight slice, so a pair covers twice the columns and A is streamed half as often (see "CTA pairs").
//
// TN kernel  C[M,N] = sum_k A[k,M]^T B[k,N]  (weight gradients; K = edges or nodes, split over CTAs)
//   * both operands are MN-major as they lie in HBM ([k, mn] row-major), fed by TMA boxes of
//     [kblock rows x 128 bytes] straight into the canonical MN-major 128B-swizzle layout -- no transposes;
//   * each CTA owns a 256 x Nblk output block (two M=128 accumulators = up to 512 TMEM columns) and a
//     contiguous K range; partial blocks go to a workspace and are summed in a fixed order (deterministic).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace cartnet {

int launch_splitk_reduce(const float* partial, int splits, int M, int N, const TnDst& dst, int64_t ldc, cudaStream_t st);

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait with a suspend-time hint: the warp sleeps in hardware (up to ~the hint, in ns) instead of burning
    // issue slots of the scheduler it shares with the epilogue warps
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// pulls a box into L2 ahead of the demand load (no shared memory, no barrier): the A stream's first touch of a tile is a
// DRAM access; with only 2-3 stages of shared memory per CTA the demand loads alone cannot keep enough bytes in flight
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// ---- CTA pairs (cta_group::2). Two CTAs of a 2-CTA cluster (the two SMs of a TPC) run ONE M = 256 MMA: each CTA stages
// its own 128 rows of A and HALF of the weight slice (BN/2 rows of B) in its own shared memory, the even ("leader") CTA's
// single MMA thread issues tcgen05.mma.cta_group::2 which reads both CTAs' shared memory and writes rows 0..127 of the
// accumulator into the leader's TMEM and rows 128..255 into the peer's. With the same 128 KB of weights per SM the pair
// covers twice the output columns, so A crosses the L2 -> SM fabric half as often -- the fabric (~7.5 TB/s measured on
// these GEMMs), not HBM, is what the 1-CTA kernel saturates when N / BN > 1. Protocol (as cute's SM100 2-SM atoms):
// TMA loads of both CTAs signal the LEADER's full barrier (address with the peer bit cleared), the leader's commits are
// multicast to both CTAs' empty / tmem_full barriers, both CTAs' epilogue warps arrive on the leader's tmem_empty.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;        // cute::Sm100MmaPeerBitMask: same offset in the pair's even CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* slot, uint32_t ncols) {
    if (CG == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        tmem_alloc(slot, ncols);
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else tmem_dealloc(taddr, ncols);
}
template <int CG>
__device__ __forceinline__ void tc_commit_cg(uint64_t* bar) {
    if (CG == 2) {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
    } else {
        tc_commit(bar);
    }
}
template <int CG>
__device__ __forceinline__ void tma_load_2d_cg(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    if (CG == 2) {
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                smem_u32(smem_dst)),
            "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar) & kPeerBitMask)
            : "memory");
    } else {
        tma_load_2d(smem_dst, map, c0, c1, bar);
    }
}
// arrive on the pair leader's copy of a barrier (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
template <bool TF32, int CG>
__device__ __forceinline__ void tc_mma_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 2 && TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else if (CG == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        tc_mma<TF32>(tmem_d, adesc, bdesc, idesc, accumulate);
    }
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane = output row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32=1 [4,6), a_format [7,10), b_format [10,13)
// (BF16 = 1, TF32 = 2), a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ inline uint32_t instr_desc(int fmt, int a_mn_major, int b_mn_major, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <typename T>
struct TcTraits;
template <>
struct TcTraits<__nv_bfloat16> {
    static constexpr int KB = 64;        // elements per 128-byte swizzle row
    static constexpr int UMMA_K = 16;
    static constexpr int FMT = 1;
    static constexpr bool TF32 = false;
    static constexpr int NP = 1;         // operand parts per element (2 = hi | lo bf16 pair)
    static constexpr int TMA_ES = 2;     // bytes per element as TMA sees the tensor
    static constexpr int TN_KROWS = 64;  // k rows per TN pipeline stage
    static constexpr CUtensorMapDataType DT = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    // MN-major operands (TN kernel): canonical SWIZZLE_128B atom = 8 k rows x 128 B
    static constexpr uint32_t MN_LAYOUT = 2, MN_SBO = 1024;
    static constexpr CUtensorMapSwizzle MN_SWIZZLE = CU_TENSOR_MAP_SWIZZLE_128B;
};
// Split precision (CARTNET_PREC_BF16X3): the tensor is a bf16 matrix of twice the logical width in which every 64-element
// chunk is [64 hi | 64 lo] (common.cuh::bf16p_t); a K block (or an MN box) of 64 elements is two 128-byte TMA boxes, and
// a product costs three kind::f16 MMAs into the same fp32 accumulator: hi*hi + hi*lo + lo*hi (lo*lo ~ 2^-18 is dropped).
template <>
struct TcTraits<bf16p_t> {
    static constexpr int KB = 64;
    static constexpr int UMMA_K = 16;
    static constexpr int FMT = 1;
    static constexpr bool TF32 = false;
    static constexpr int NP = 2;
    static constexpr int TMA_ES = 2;
    static constexpr int TN_KROWS = 32;
    static constexpr CUtensorMapDataType DT = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    static constexpr uint32_t MN_LAYOUT = 2, MN_SBO = 1024;
    static constexpr CUtensorMapSwizzle MN_SWIZZLE = CU_TENSOR_MAP_SWIZZLE_128B;
};
template <>
struct TcTraits<tf32_t> {
    static constexpr int KB = 32;
    static constexpr int UMMA_K = 8;
    static constexpr int FMT = 2;
    static constexpr bool TF32 = true;
    static constexpr int NP = 1;
    static constexpr int TMA_ES = 4;
    static constexpr int TN_KROWS = 32;
    static constexpr CUtensorMapDataType DT = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    // 32-bit MN-major operands only exist as SWIZZLE_128B_BASE32B (Swizzle<2,5,2>): atom = 4 k rows x 128 B,
    // 32-byte chunks XORed with the row index; TMA writes it with SWIZZLE_128B_ATOM_32B
    static constexpr uint32_t MN_LAYOUT = 1, MN_SBO = 512;
    static constexpr CUtensorMapSwizzle MN_SWIZZLE = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
};

// ------------------------------------------------------------------------------------------ specialised epilogue
// The epilogue is the critical path of these GEMMs (K is only 256..1024), so it is specialised at compile time
// per use-site: EPI is a bit mask of the steps of cartnet_gemm_t that are present; EPI_GENERIC keeps the
// runtime-checked version for any other combination.
enum : int { EB_BIAS = 1, EB_GATHER = 2, EB_ZOUT = 4, EB_SILU = 8, EB_DSILU = 16, EB_RESID = 32, EB_OUTF = 64, EB_OUTT = 128,
              EB_STATS = 256 };   // EB_STATS: per-column sum / sum of squares of the output rides in the epilogue (BatchNorm statistics)
constexpr int EPI_GENERIC = -1;

// sigmoid(v) = 0.5 tanh(v/2) + 0.5: ONE MUFU op (tanh.approx, ~2^-11 rel. error -- below the bf16 / tf32 operand
// rounding of these modes) instead of EX2 + RCP; MUFU is quarter-rate and was ~25% of the first Linear's epilogue.
__device__ __forceinline__ float tanh_approx(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float silu_fast(float v) {
    const float h = 0.5f * v;
    return fmaf(h, tanh_approx(h), h);                      // v * sigmoid(v)
}
__device__ __forceinline__ float dsilu_fast(float z) {
    const float sg = fmaf(0.5f, tanh_approx(0.5f * z), 0.5f);
    return sg * fmaf(z, 1.0f - sg, 1.0f);                   // s (1 + z (1 - s))
}

template <int EPI>
__host__ __device__ constexpr bool epi_has(int bit, bool runtime) { return EPI == EPI_GENERIC ? runtime : ((EPI & bit) != 0); }

// 8 consecutive elements of a T-typed row, unconverted: 16 bytes of bf16, 32 bytes of tf32 words, or the 16 + 16 bytes of a
// bf16 pair run (8 elements never straddle a 64-element chunk when the column is a multiple of 8)
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { uint4 a; };
template <> struct Raw8<tf32_t> { float4 a, b; };
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<bf16p_t> { uint4 hi, lo; };
template <> struct Raw8<__half> { uint4 a; };
__device__ __forceinline__ Raw8<__half> ld_raw8(const __half* p) { Raw8<__half> r; r.a = __ldg(reinterpret_cast<const uint4*>(p)); return r; }
__device__ __forceinline__ void cvt_raw8(const Raw8<__half>& r, float4& v0, float4& v1) {
    half4raw lo, hi;
    lo.v = make_uint2(r.a.x, r.a.y); hi.v = make_uint2(r.a.z, r.a.w);
    v0 = cvt_raw4(lo); v1 = cvt_raw4(hi);
}
__device__ __forceinline__ void store8(__half* p, const float4& v0, const float4& v1) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_half2_sat(v0.x, v0.y), pack_half2_sat(v0.z, v0.w), pack_half2_sat(v1.x, v1.y), pack_half2_sat(v1.z, v1.w));
}
__device__ __forceinline__ Raw8<__nv_bfloat16> ld_raw8(const __nv_bfloat16* p) { Raw8<__nv_bfloat16> r; r.a = __ldg(reinterpret_cast<const uint4*>(p)); return r; }
__device__ __forceinline__ Raw8<tf32_t> ld_raw8(const tf32_t* p) {
    Raw8<tf32_t> r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    return r;
}
__device__ __forceinline__ Raw8<float> ld_raw8(const float* p) {
    Raw8<float> r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    return r;
}
__device__ __forceinline__ Raw8<bf16p_t> ld_raw8(const bf16p_t* p) {
    const char* h = pair_hi_addr(p);
    Raw8<bf16p_t> r;
    r.hi = __ldg(reinterpret_cast<const uint4*>(h));
    r.lo = __ldg(reinterpret_cast<const uint4*>(h + 128));
    return r;
}
__device__ __forceinline__ void cvt_raw8(const Raw8<__nv_bfloat16>& r, float4& v0, float4& v1) {
    const float2 a = bf162_to_float2(r.a.x), b = bf162_to_float2(r.a.y), c = bf162_to_float2(r.a.z), d = bf162_to_float2(r.a.w);
    v0 = make_float4(a.x, a.y, b.x, b.y);
    v1 = make_float4(c.x, c.y, d.x, d.y);
}
__device__ __forceinline__ void cvt_raw8(const Raw8<tf32_t>& r, float4& v0, float4& v1) { v0 = r.a; v1 = r.b; }
__device__ __forceinline__ void cvt_raw8(const Raw8<float>& r, float4& v0, float4& v1) { v0 = r.a; v1 = r.b; }
__device__ __forceinline__ void cvt_raw8(const Raw8<bf16p_t>& r, float4& v0, float4& v1) {
    v0 = join4_bf16(make_uint2(r.hi.x, r.hi.y), make_uint2(r.lo.x, r.lo.y));
    v1 = join4_bf16(make_uint2(r.hi.z, r.hi.w), make_uint2(r.lo.z, r.lo.w));
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float4& v0, const float4& v1) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v0.x, v0.y), b = __floats2bfloat162_rn(v0.z, v0.w);
    const __nv_bfloat162 c = __floats2bfloat162_rn(v1.x, v1.y), d = __floats2bfloat162_rn(v1.z, v1.w);
    *reinterpret_cast<uint4*>(p) = make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                                              *reinterpret_cast<const uint32_t*>(&c), *reinterpret_cast<const uint32_t*>(&d));
}
__device__ __forceinline__ void store8(tf32_t* p, const float4& v0, const float4& v1) {
    reinterpret_cast<float4*>(p)[0] = make_float4(round_tf32(v0.x), round_tf32(v0.y), round_tf32(v0.z), round_tf32(v0.w));
    reinterpret_cast<float4*>(p)[1] = make_float4(round_tf32(v1.x), round_tf32(v1.y), round_tf32(v1.z), round_tf32(v1.w));
}
__device__ __forceinline__ void store8(float* p, const float4& v0, const float4& v1) {
    reinterpret_cast<float4*>(p)[0] = v0;
    reinterpret_cast<float4*>(p)[1] = v1;
}
__device__ __forceinline__ void store8(bf16p_t* p, const float4& v0, const float4& v1) {
    uint2 h0, l0, h1, l1;
    split4_bf16(v0, h0, l0);
    split4_bf16(v1, h1, l1);
    char* h = pair_hi_addr(p);
    *reinterpret_cast<uint4*>(h) = make_uint4(h0.x, h0.y, h1.x, h1.y);
    *reinterpret_cast<uint4*>(h + 128) = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

// CPT consecutive columns of one row (CPT = 4 or 8) as NV = CPT / 4 float4s, through the 4- or 8-element raw accessors.
// Which CPT is best depends on the element size: a 32-column row segment of 4-byte elements is a full 128-byte line when 8
// lanes own 4 columns each (one 16-byte access per lane), while 2-byte elements want 8 columns per lane.
template <typename T, int CPT> struct RawV;
template <typename T> struct RawV<T, 4> { typename Raw4<T>::type r; };
template <typename T> struct RawV<T, 8> { Raw8<T> r; };
template <typename T> __device__ __forceinline__ void ldv(RawV<T, 4>& o, const T* p) { o.r = ld_raw4<T>(p); }
template <typename T> __device__ __forceinline__ void ldv(RawV<T, 8>& o, const T* p) { o.r = ld_raw8(p); }
template <typename T> __device__ __forceinline__ void cvtv(const RawV<T, 4>& r, float4* v) { v[0] = cvt_raw4(r.r); }
template <typename T> __device__ __forceinline__ void cvtv(const RawV<T, 8>& r, float4* v) { cvt_raw8(r.r, v[0], v[1]); }
template <int CPT, typename T> __device__ __forceinline__ void storev(T* p, const float4* v) {
    if (CPT == 4) store4<T>(p, v[0]); else store8(p, v[0], v[1]);
}

// the arithmetic of the epilogue on 4 consecutive columns of one row (no stores): v = acc + bias + gathers; zval = v;
// v = act(v, z_in); v += resid
template <typename T, int EPI>
__device__ __forceinline__ void epi_math4(const EpiParams<T>& p, float4& v, float4& zval, const float4& bias4, bool has_g0, bool has_g1,
                                          const float4& ga, const float4& gb, const float4& z, const float4& rs) {
    if (epi_has<EPI>(EB_BIAS, p.bias != nullptr)) { v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w; }
    if (epi_has<EPI>(EB_GATHER, has_g0)) { v.x += ga.x; v.y += ga.y; v.z += ga.z; v.w += ga.w; }
    if (epi_has<EPI>(EB_GATHER, has_g1)) { v.x += gb.x; v.y += gb.y; v.z += gb.z; v.w += gb.w; }
    zval = v;
    if (epi_has<EPI>(EB_SILU, p.act == CARTNET_ACT_SILU)) {
        v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w);
    }
    if (epi_has<EPI>(EB_DSILU, p.act == CARTNET_ACT_MUL_DSILU)) {
        v.x *= dsilu_fast(z.x); v.y *= dsilu_fast(z.y); v.z *= dsilu_fast(z.z); v.w *= dsilu_fast(z.w);
    }
    if (epi_has<EPI>(EB_RESID, p.resid != nullptr)) { v.x += rs.x; v.y += rs.y; v.z += rs.z; v.w += rs.w; }
}

// The 32 / LPR row groups of a warp hold partial column sums of the same CPT columns: combine them with a fixed butterfly
// (lanes that differ only in the row-group bits), then the owner lanes (row group 0) add the 32-row block sums to the warp's
// running totals in shared memory.
template <int CPT>
__device__ __forceinline__ void stats_flushv(float* wstat, int cidx, int lane, float4* ss, float4* sq) {
    constexpr int NV = CPT / 4, LPR = 32 / CPT;
#pragma unroll
    for (int h = 0; h < NV; ++h) {
#pragma unroll
        for (int o = LPR; o <= 16; o <<= 1) {
            ss[h].x += __shfl_xor_sync(0xffffffffu, ss[h].x, o); ss[h].y += __shfl_xor_sync(0xffffffffu, ss[h].y, o);
            ss[h].z += __shfl_xor_sync(0xffffffffu, ss[h].z, o); ss[h].w += __shfl_xor_sync(0xffffffffu, ss[h].w, o);
            sq[h].x += __shfl_xor_sync(0xffffffffu, sq[h].x, o); sq[h].y += __shfl_xor_sync(0xffffffffu, sq[h].y, o);
            sq[h].z += __shfl_xor_sync(0xffffffffu, sq[h].z, o); sq[h].w += __shfl_xor_sync(0xffffffffu, sq[h].w, o);
        }
        if (lane < LPR) {
            float4* a = reinterpret_cast<float4*>(wstat + cidx + 4 * h);
            float4* b = reinterpret_cast<float4*>(wstat + 128 + cidx + 4 * h);
            float4 va = *a, vb = *b;
            va.x += ss[h].x; va.y += ss[h].y; va.z += ss[h].z; va.w += ss[h].w;
            vb.x += sq[h].x; vb.y += sq[h].y; vb.z += sq[h].z; vb.w += sq[h].w;
            *a = va; *b = vb;
        }
    }
}

// ------------------------------------------------------------------------------------------ NT kernel
constexpr int NT_STAGES = 6;                     // at most; the host picks the deepest ring that fits next to the weight slice
constexpr int NT_A_PART_BYTES = 128 * 128;       // 128 rows x 128 B (one operand part of one K block)
constexpr int NT_EPI_WARPS = 8;
constexpr int NT_THREADS = 64 + 32 * NT_EPI_WARPS;
constexpr int NT_STG_PITCH = 32;                                  // floats; 16-byte chunks XOR-swizzled with the row: conflict-free float4 accesses both ways
constexpr int NT_STG_BYTES = NT_EPI_WARPS * 32 * NT_STG_PITCH * 4;   // per-warp 32x32 fp32 transpose buffers

struct NtBars {
    uint64_t a_full[NT_STAGES], a_empty[NT_STAGES], b_full, tmem_full[2], tmem_empty[2];
    uint32_t tmem_slot;
};
constexpr int NT_STAT_BYTES = NT_EPI_WARPS * 2 * 128 * 4;          // per-warp [sum | sumsq][<= 128 columns] (EB_STATS only)

// CG = 1: one CTA per tile of 128 rows. CG = 2: a CTA pair per tile of 256 rows (see "CTA pairs" above; 4-byte operands);
// m_tiles counts tiles of 128 * CG rows, BN is the pair's column count (each CTA keeps BN / CG rows of B).
template <typename T, int EPI, int CG = 1>
__global__ void __launch_bounds__(NT_THREADS, 1)
tc_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
             int BN, int n_tiles, int m_tiles, int stages, EpiParams<T> epi, double* __restrict__ stats) {
    using TR = TcTraits<T>;
    constexpr int NP = TR::NP;
    const bool prefetch = (stages & 16) != 0;        // host flags folded into `stages`
    const bool uni_path = (stages & 32) != 0;        // gather0 rows that are equal over a warp's 32 rows are loaded once
    const bool gpf = (stages & 192) != 0;            // L1 prefetch of the next block's gathered rows
    const int gpf_dist = (stages & 128) ? 2 : 1;
    stages &= 15;
    constexpr int KBOX = 128 / TR::TMA_ES;           // TMA elements per 128-byte box row
    constexpr int A_STAGE_BYTES = NP * NT_A_PART_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int kblks = K / TR::KB;
    const int rank = CG == 2 ? (int)cluster_ctarank() : 0;       // 0 = the CTA (of a pair) that issues the MMAs
    const int b_part_bytes = (BN / CG) * 128;
    const int b_kb_bytes = NP * b_part_bytes;
    uint8_t* smemB = smem;
    uint8_t* smemA = smem + (size_t)kblks * b_kb_bytes;
    float* smemStg = reinterpret_cast<float*>(smemA + stages * A_STAGE_BYTES);
    NtBars* bars = reinterpret_cast<NtBars*>(reinterpret_cast<uint8_t*>(smemStg) + NT_STG_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = (int)blockIdx.x / CG, units = (int)gridDim.x / CG;      // CTA (pair) index
    const int n_tile = unit % n_tiles;
    const int m_first = unit / n_tiles, m_stride = units / n_tiles;
    const int n0 = n_tile * BN;
    const uint32_t tmem_cols = (uint32_t)(2 * BN < 32 ? 32 : 2 * BN);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NT_STAGES; ++s) { mbar_init(&bars->a_full[s], 1); mbar_init(&bars->a_empty[s], 1); }
        mbar_init(&bars->b_full, 1);
        // BN <= 32: only the first column group (4 warps) has columns; idle warps must not arrive (they would run ahead
        // of the MMA warp and complete phases it has not waited for yet)
        for (int a = 0; a < 2; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], CG * (BN > 32 ? NT_EPI_WARPS : NT_EPI_WARPS / 2)); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_cg<CG>(&bars->tmem_slot, tmem_cols);
    tc_fence_before();
    if (CG == 2) cluster_sync_all();      // the peer's barriers exist before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            if (rank == 0) mbar_expect_tx(&bars->b_full, (uint32_t)(CG * kblks * b_kb_bytes));      // both CTAs' halves
            for (int kb = 0; kb < kblks; ++kb)
#pragma unroll
                for (int pt = 0; pt < NP; ++pt)
                    tma_load_2d_cg<CG>(smemB + (size_t)kb * b_kb_bytes + pt * b_part_bytes, &tmB, (kb * NP + pt) * KBOX, n0 + rank * (BN / CG), &bars->b_full);
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = m_first; mt < m_tiles; mt += m_stride) {
                for (int kb = 0; kb < kblks; ++kb) {
                    // L2 prefetch of the same K block of this CTA's NEXT tile (one of the CTAs that share the tile does it)
                    if (prefetch && n_tile == 0 && mt + m_stride < m_tiles) {
#pragma unroll
                        for (int pt = 0; pt < NP; ++pt) tma_prefetch_2d(&tmA, (kb * NP + pt) * KBOX, ((mt + m_stride) * CG + rank) * 128);
                    }
                    mbar_wait(&bars->a_empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&bars->a_full[stage], CG * A_STAGE_BYTES);
#pragma unroll
                    for (int pt = 0; pt < NP; ++pt)
                        tma_load_2d_cg<CG>(smemA + stage * A_STAGE_BYTES + pt * NT_A_PART_BYTES, &tmA, (kb * NP + pt) * KBOX, (mt * CG + rank) * 128, &bars->a_full[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
            if (CG == 2) {
                // the leader's last commits still arrive on this CTA's barriers: do not leave before every stage was released
                for (int i = 0; i < stages; ++i) {
                    mbar_wait(&bars->a_empty[stage], phase ^ 1);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (pairs: the leader CTA's only)
        static_assert(CG == 1 || sizeof(T) == 4, "CTA pairs are built for the 4-byte operand modes (bf16 pairs, tf32)");
        if (rank == 0) {
        const uint32_t idesc = instr_desc(TR::FMT, 0, 0, 128 * CG, BN);
        mbar_wait(&bars->b_full, 0);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int mt = m_first; mt < m_tiles; mt += m_stride) {
            mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < kblks; ++kb) {
                mbar_wait(&bars->a_full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smemA + stage * A_STAGE_BYTES);
                    const uint32_t b_addr = smem_u32(smemB + (size_t)kb * b_kb_bytes);
#pragma unroll
                    for (int j = 0; j < TR::KB / TR::UMMA_K; ++j) {   // 4 MMAs of K = 32 bytes inside the swizzle row
                        const uint64_t ad = smem_desc(a_addr + j * 32, 16, 1024);
                        const uint64_t bd = smem_desc(b_addr + j * 32, 16, 1024);
                        if (NP == 2) {      // split precision: the two cross terms first, then hi*hi, all into one fp32 accumulator
                            const uint64_t al = smem_desc(a_addr + NT_A_PART_BYTES + j * 32, 16, 1024);
                            const uint64_t bl = smem_desc(b_addr + b_part_bytes + j * 32, 16, 1024);
                            tc_mma_cg<false, CG>(tmem_base + (uint32_t)(acc * BN), al, bd, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                            tc_mma_cg<false, CG>(tmem_base + (uint32_t)(acc * BN), ad, bl, idesc, 1u);
                            tc_mma_cg<false, CG>(tmem_base + (uint32_t)(acc * BN), ad, bd, idesc, 1u);
                        } else {
                            tc_mma_cg<TR::TF32, CG>(tmem_base + (uint32_t)(acc * BN), ad, bd, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        }
                    }
                    tc_commit_cg<CG>(&bars->a_empty[stage]);           // frees the A stage (in both CTAs of a pair) when the MMAs retire
                    if (kb == kblks - 1) tc_commit_cg<CG>(&bars->tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        }
    } else {
        // ------------------------------------------------ epilogue: TMEM -> registers -> smem transpose -> fused epilogue -> HBM
        // tcgen05.ld hands each thread one ROW (32 columns). Going to HBM from that layout would make every warp
        // access touch 32 different 128-byte lines with 16 bytes each, so each 32x32 block is first transposed
        // through a per-warp shared-memory buffer: afterwards 8 lanes cover one row's 32 columns (one full line),
        // 4 rows per instruction, and the gathers / residual reads / stores of the epilogue are all line-coalesced.
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;              // column group (0/1)
        const int cols_per_grp = (BN / 2) < 32 ? 32 : (BN / 2);
        const int c_begin = grp * cols_per_grp;
        const int c_end = (c_begin + cols_per_grp) < BN ? (c_begin + cols_per_grp) : BN;
        float* stg = smemStg + (warp - 2) * 32 * NT_STG_PITCH;
        // columns per thread. Measured on the ADP-64 shapes (same box, scripts/gemm_microbench.py): 8 columns win wherever the
        // epilogue reads row-gathered or pre-activation tensors (GEMM1 pair mode 0.925 vs 1.07 ms: one full 32-byte sector per
        // lane and half the load instructions) and for 2-byte operands; 4 columns (a full 128-byte line per row and
        // instruction) win by ~3 % for the store-only / residual epilogues of the 4-byte modes.
#ifdef CN_NT_CPT_FORCE
        constexpr int CPT = CN_NT_CPT_FORCE;        // A/B builds only
#else
        constexpr int CPT = (sizeof(T) == 2 || EPI < 0 || (EPI & (EB_GATHER | EB_DSILU)) != 0) ? 8 : 4;
#endif
        constexpr int NV = CPT / 4;                 // float4s per thread and row
        constexpr int LPR = 32 / CPT;               // lanes per 32-column row segment
        constexpr int RPI = 32 / LPR;               // rows per instruction
        constexpr int NIT = 32 / RPI;               // row iterations per 32-row block
        const int sub_r = lane / LPR, sub_c = (lane % LPR) * CPT;
        // EB_STATS: this warp's running column sums over all of its tiles (fp32; the caller centres the output so that
        // |mean| <~ std), one owner lane per column -> fixed order, no atomics; written out as fp64 partials at the end
        float* wstat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + ((sizeof(NtBars) + 15) & ~(size_t)15)) + (warp - 2) * 256;
        constexpr bool kStats = EPI != EPI_GENERIC && (EPI & EB_STATS) != 0;
        // pair operands: the weight slice leaves no shared memory for per-warp statistics and BN <= 128 (at most two
        // 32-column blocks per warp), so each thread keeps the running sums of its 8 columns in registers over ALL its
        // tiles; the 8 row groups are combined once, at the end, by the same fixed butterfly
        constexpr bool kStatsRegs = kStats && TR::NP == 2;
        constexpr int NBW = 2 * CG;                 // 32-column blocks per warp and tile (BN <= 128 * CG)
        float4 rs_s[NBW][NV], rs_q[NBW][NV];
#pragma unroll
        for (int a = 0; a < NBW; ++a)
#pragma unroll
            for (int h = 0; h < NV; ++h) rs_s[a][h] = rs_q[a][h] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kStats && !kStatsRegs) {
            for (int i = lane; i < 256; i += 32) wstat[i] = 0.f;
            __syncwarp();
        }
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int mt = m_first; mt < m_tiles && c_begin < c_end; mt += m_stride) {
            const int64_t row0 = ((int64_t)mt * CG + rank) * 128 + q * 32;
            // gather-row indices of the warp's 32 output rows, one row per lane (hoisted out of the column loop); the thread
            // that finishes row it * RPI + sub_r fetches its index with a shuffle
            int32_t i0v = 0, i1v = 0;
            if (epi_has<EPI>(EB_GATHER, epi.gather0 != nullptr)) {
                const int64_t row = row0 + lane;
                i0v = row < M ? epi.gidx0[row] : 0;
                i1v = (row < M && epi.gather1 != nullptr) ? epi.gidx1[row] : 0;
            }
            const bool full = row0 + 32 <= (int64_t)M;
            // gather0 is indexed by the CSR-sorted destination: ~70 % of the 32-row blocks have ONE destination, whose row is
            // then loaded once per thread instead of once per row (-3 KB of the 22 KB a block moves through the LSU)
            const bool uni = uni_path && epi_has<EPI>(EB_GATHER, epi.gather0 != nullptr) &&
                             __all_sync(0xffffffffu, i0v == __shfl_sync(0xffffffffu, i0v, 0));
            bool first = true;
            for (int c = c_begin; c < c_end; c += 32) {
                const int col = n0 + c + sub_c;
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                // L1 prefetch of the NEXT column block's gathered row segments (one 128-byte line per row: lane = row): the
                // first use of the gathered data is where these warps wait (ncu: 22 % of all stall samples on that one FADD)
                if (gpf && epi_has<EPI>(EB_GATHER, epi.gather0 != nullptr) && row0 + lane < M) {
                    const int cn = c + 32 * gpf_dist;
                    if (cn < c_end) {
                        if (!uni) prefetch_l1(epi.gather0 + (int64_t)i0v * epi.ldg + n0 + cn);
                        prefetch_l1(epi.gather1 + (int64_t)i1v * epi.ldg + n0 + cn);
                    }
                    if (c == c_begin && gpf_dist == 2 && c + 32 < c_end) {
                        if (!uni) prefetch_l1(epi.gather0 + (int64_t)i0v * epi.ldg + n0 + c + 32);
                        prefetch_l1(epi.gather1 + (int64_t)i1v * epi.ldg + n0 + c + 32);
                    }
                }
                float4 bias[NV];
#pragma unroll
                for (int h = 0; h < NV; ++h)
                    bias[h] = epi_has<EPI>(EB_BIAS, epi.bias != nullptr) ? *reinterpret_cast<const float4*>(epi.bias + col + 4 * h) : zero;
                const bool has_g0 = epi_has<EPI>(EB_GATHER, epi.gather0 != nullptr);
                const bool has_g1 = epi_has<EPI>(EB_GATHER, epi.gather1 != nullptr);
                const bool has_z = epi_has<EPI>(EB_DSILU, epi.act == CARTNET_ACT_MUL_DSILU);
                const bool has_r = epi_has<EPI>(EB_RESID, epi.resid != nullptr);
                // phase 1: every global read of this 32x32 block is issued up front -- before the accumulator is even
                // fetched from TMEM, so that their L2 / HBM round trip overlaps the tcgen05.ld + shared-memory transpose
                // (and, for a tile's first block, the wait for its MMAs) -- and before any store, whose possible aliasing
                // would otherwise serialise the round trips. A thread owns CPT consecutive columns of NIT rows; every
                // global access is a 16-byte (or, for pair halves at CPT = 4, 8-byte) vector.
                // FULL blocks (all 32 rows < M, i.e. every block but the last few) carry no per-row predicates so the
                // compiler can interleave the independent rows and hide the MUFU / FMA latencies.
                using G = typename GatherOf<T>::type;
                RawV<G, CPT> ra[NIT], rb[NIT];
                RawV<typename ZOf<T>::type, CPT> rz[NIT];
                float4 rr[NIT][NV];
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int64_t row = row0 + it * RPI + sub_r;
                    const int32_t i0 = has_g0 ? __shfl_sync(0xffffffffu, i0v, it * RPI + sub_r) : 0;
                    const int32_t i1 = has_g1 ? __shfl_sync(0xffffffffu, i1v, it * RPI + sub_r) : 0;
                    if (full || row < M) {
                        if (has_g0 && (!uni || it == 0)) ldv(ra[it], epi.gather0 + (int64_t)i0 * epi.ldg + col);
                        if (has_g1) ldv(rb[it], epi.gather1 + (int64_t)i1 * epi.ldg + col);
                        if (has_z) ldv(rz[it], epi.z_in + row * epi.ldzin + col);
                        if (has_r) {
#pragma unroll
                            for (int h = 0; h < NV; ++h) rr[it][h] = __ldg(reinterpret_cast<const float4*>(epi.resid + row * epi.ldr + col + 4 * h));
                        }
                    }
                }
                if (first) {
                    mbar_wait(&bars->tmem_full[acc], acc_phase);
                    tc_fence_after();
                    first = false;
                }
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stg + lane * NT_STG_PITCH + 4 * (j ^ (lane & 7))) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                float4 tv[NIT][NV];
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int srow = it * RPI + sub_r;
#pragma unroll
                    for (int h = 0; h < NV; ++h)
                        tv[it][h] = *reinterpret_cast<const float4*>(stg + srow * NT_STG_PITCH + 4 * (((sub_c >> 2) + h) ^ (srow & 7)));
                }
                __syncwarp();
                float4 ss[NV], sq[NV];
#pragma unroll
                for (int h = 0; h < NV; ++h) ss[h] = sq[h] = zero;
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int64_t row = row0 + it * RPI + sub_r;
                    if (full || row < M) {
                        float4 ga[NV], gb[NV], zi[NV], zv[NV], o[NV];
#pragma unroll
                        for (int h = 0; h < NV; ++h) ga[h] = gb[h] = zi[h] = zero;
                        if (has_g0) {
                            if (uni) cvtv(ra[0], ga);
                            else cvtv(ra[it], ga);
                        }
                        if (has_g1) cvtv(rb[it], gb);
                        if (has_z) cvtv(rz[it], zi);
#pragma unroll
                        for (int h = 0; h < NV; ++h) {
                            o[h] = tv[it][h];
                            epi_math4<T, EPI>(epi, o[h], zv[h], bias[h], has_g0, has_g1, ga[h], gb[h], zi[h], has_r ? rr[it][h] : zero);
                        }
                        if (epi_has<EPI>(EB_ZOUT, epi.z_out != nullptr)) storev<CPT>(epi.z_out + row * epi.ldz + col, zv);
                        if (epi_has<EPI>(EB_OUTF, epi.out_f32 != nullptr)) {
#pragma unroll
                            for (int h = 0; h < NV; ++h) *reinterpret_cast<float4*>(epi.out_f32 + row * epi.ldo + col + 4 * h) = o[h];
                        }
                        if (epi_has<EPI>(EB_OUTT, epi.out_t != nullptr)) storev<CPT>(epi.out_t + row * epi.ldt + col, o);
                        if (kStats) {
#pragma unroll
                            for (int h = 0; h < NV; ++h) {
                                ss[h].x += o[h].x; ss[h].y += o[h].y; ss[h].z += o[h].z; ss[h].w += o[h].w;
                                sq[h].x = fmaf(o[h].x, o[h].x, sq[h].x); sq[h].y = fmaf(o[h].y, o[h].y, sq[h].y);
                                sq[h].z = fmaf(o[h].z, o[h].z, sq[h].z); sq[h].w = fmaf(o[h].w, o[h].w, sq[h].w);
                            }
                        }
                    }
                }
                if (kStatsRegs) {
                    const int cb = (c - c_begin) >> 5;          // 0 .. NBW-1
#pragma unroll
                    for (int a = 0; a < NBW; ++a)
                        if (a == cb) {
#pragma unroll
                            for (int h = 0; h < NV; ++h) {
                                rs_s[a][h].x += ss[h].x; rs_s[a][h].y += ss[h].y; rs_s[a][h].z += ss[h].z; rs_s[a][h].w += ss[h].w;
                                rs_q[a][h].x += sq[h].x; rs_q[a][h].y += sq[h].y; rs_q[a][h].z += sq[h].z; rs_q[a][h].w += sq[h].w;
                            }
                        }
                } else if (kStats) {
                    stats_flushv<CPT>(wstat, c - c_begin + sub_c, lane, ss, sq);      // rows >= M contribute nothing
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_leader(&bars->tmem_empty[acc]);
                else mbar_arrive(&bars->tmem_empty[acc]);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (kStatsRegs) {
            const int64_t blk = (int64_t)((unit / n_tiles) * CG + rank) * 4 + q;
#pragma unroll
            for (int a = 0; a < NBW; ++a) {
                if (c_begin + 32 * a >= c_end) break;
#pragma unroll
                for (int h = 0; h < NV; ++h) {
#pragma unroll
                    for (int o = LPR; o <= 16; o <<= 1) {
                        rs_s[a][h].x += __shfl_xor_sync(0xffffffffu, rs_s[a][h].x, o); rs_s[a][h].y += __shfl_xor_sync(0xffffffffu, rs_s[a][h].y, o);
                        rs_s[a][h].z += __shfl_xor_sync(0xffffffffu, rs_s[a][h].z, o); rs_s[a][h].w += __shfl_xor_sync(0xffffffffu, rs_s[a][h].w, o);
                        rs_q[a][h].x += __shfl_xor_sync(0xffffffffu, rs_q[a][h].x, o); rs_q[a][h].y += __shfl_xor_sync(0xffffffffu, rs_q[a][h].y, o);
                        rs_q[a][h].z += __shfl_xor_sync(0xffffffffu, rs_q[a][h].z, o); rs_q[a][h].w += __shfl_xor_sync(0xffffffffu, rs_q[a][h].w, o);
                    }
                    if (lane < LPR) {
                        const int64_t cg = n0 + c_begin + 32 * a + sub_c + 4 * h;
                        double* ps = stats + (blk * 2 + 0) * N + cg;
                        double* pq = stats + (blk * 2 + 1) * N + cg;
                        ps[0] = (double)rs_s[a][h].x; ps[1] = (double)rs_s[a][h].y; ps[2] = (double)rs_s[a][h].z; ps[3] = (double)rs_s[a][h].w;
                        pq[0] = (double)rs_q[a][h].x; pq[1] = (double)rs_q[a][h].y; pq[2] = (double)rs_q[a][h].z; pq[3] = (double)rs_q[a][h].w;
                    }
                }
            }
        } else if (kStats) {
            // one partial row per (m-CTA, TMEM lane quarter): [blk][sum | sumsq][N] fp64, summed in a fixed order afterwards
            __syncwarp();
            const int64_t blk = (int64_t)((unit / n_tiles) * CG + rank) * 4 + q;
            for (int i = lane; i < c_end - c_begin; i += 32) {
                stats[(blk * 2 + 0) * N + n0 + c_begin + i] = (double)wstat[i];
                stats[(blk * 2 + 1) * N + n0 + c_begin + i] = (double)wstat[128 + i];
            }
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();      // the leader's MMAs read the peer's shared memory until its last commit
    else __syncthreads();
    if (warp == 2) tmem_dealloc_cg<CG>(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------ TN kernel
constexpr int TN_STAGES = 3;
constexpr int TN_THREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int TN_MBLK = 256;

struct TnBars {
    uint64_t full[TN_STAGES], empty[TN_STAGES], tmem_full;
    uint32_t tmem_slot;
};

template <typename T>
__global__ void __launch_bounds__(TN_THREADS, 1)
tc_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int64_t K, int M, int N,
             int Nblk, int n_blocks, int kblks_total, int kblks_per_split, float* __restrict__ partial) {
    using TR = TcTraits<T>;
    constexpr int NP = TR::NP;
    constexpr int BOXW = TR::KB;                    // mn elements per 128-byte row
    constexpr int KROWS = TR::TN_KROWS;             // k rows per stage: 64 (bf16) / 32 (tf32, bf16 pairs)
    constexpr int BOX_BYTES = KROWS * 128;
    constexpr int A_BOXES = TN_MBLK / BOXW;
    constexpr int A_PART_BYTES = A_BOXES * BOX_BYTES;
    constexpr int A_BYTES = NP * A_PART_BYTES;      // 32 KB
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_boxes = Nblk / BOXW;
    const int b_part_bytes = b_boxes * BOX_BYTES;
    const int stage_bytes = A_BYTES + NP * b_part_bytes;
    TnBars* bars = reinterpret_cast<TnBars*>(smem + (size_t)TN_STAGES * stage_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int mb = blockIdx.y / n_blocks, nb = blockIdx.y % n_blocks;
    const int kb0 = split * kblks_per_split;
    const int kb1 = (kb0 + kblks_per_split) < kblks_total ? (kb0 + kblks_per_split) : kblks_total;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(2 * Nblk)) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TN_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bars->tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&bars->empty[stage], phase ^ 1);
                mbar_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                uint8_t* sb = sa + A_BYTES;
                // pairs: element column c (a multiple of 64) = bf16 columns [2c, 2c+64) (hi) and [2c+64, 2c+128) (lo)
#pragma unroll
                for (int pt = 0; pt < NP; ++pt) {
                    for (int i = 0; i < A_BOXES; ++i)
                        tma_load_2d(sa + pt * A_PART_BYTES + i * BOX_BYTES, &tmA, NP * (mb * TN_MBLK + i * BOXW) + pt * BOXW, kb * KROWS, &bars->full[stage]);
                    for (int i = 0; i < b_boxes; ++i)
                        tma_load_2d(sb + pt * b_part_bytes + i * BOX_BYTES, &tmB, NP * (nb * Nblk + i * BOXW) + pt * BOXW, kb * KROWS, &bars->full[stage]);
                }
                if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = instr_desc(TR::FMT, 1, 1, 128, Nblk);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                for (int j = 0; j < KROWS / TR::UMMA_K; ++j) {
                    // MN-major canonical layout: LBO = distance between 128-byte-wide MN chunks (one TMA box),
                    // SBO = distance between swizzle atoms along k (8 rows bf16 / 4 rows tf32); UMMA_K rows = UMMA_K * 128 bytes
                    const uint32_t koff = (uint32_t)(j * TR::UMMA_K * 128);
                    const uint64_t bd = smem_desc(b_addr + koff, BOX_BYTES, TR::MN_SBO, TR::MN_LAYOUT);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint64_t ad = smem_desc(a_addr + h * (A_PART_BYTES / 2) + koff, BOX_BYTES, TR::MN_SBO, TR::MN_LAYOUT);
                        const uint32_t first = (kb > kb0 || j > 0) ? 1u : 0u;
                        if (NP == 2) {
                            const uint64_t al = smem_desc(a_addr + A_PART_BYTES + h * (A_PART_BYTES / 2) + koff, BOX_BYTES, TR::MN_SBO, TR::MN_LAYOUT);
                            const uint64_t bl = smem_desc(b_addr + b_part_bytes + koff, BOX_BYTES, TR::MN_SBO, TR::MN_LAYOUT);
                            tc_mma<false>(tmem_base + (uint32_t)(h * Nblk), al, bd, idesc, first);
                            tc_mma<false>(tmem_base + (uint32_t)(h * Nblk), ad, bl, idesc, 1u);
                            tc_mma<false>(tmem_base + (uint32_t)(h * Nblk), ad, bd, idesc, 1u);
                        } else {
                            tc_mma<TR::TF32>(tmem_base + (uint32_t)(h * Nblk), ad, bd, idesc, first);
                        }
                    }
                }
                tc_commit(&bars->empty[stage]);
                if (kb == kb1 - 1) tc_commit(&bars->tmem_full);
            }
            __syncwarp();
            if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        const int q = warp & 3;
        mbar_wait(&bars->tmem_full, 0);
        tc_fence_after();
        for (int h = 0; h < 2; ++h) {
            const int row = mb * TN_MBLK + h * 128 + q * 32 + lane;
            if (mb * TN_MBLK + h * 128 >= M) break;        // M = odd multiple of 128: the last block's second half is TMA zero fill
            float* dst = partial + ((int64_t)split * M + row) * N + (int64_t)nb * Nblk;
            for (int c = 0; c < Nblk; c += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * Nblk + c), v);
#pragma unroll
                for (int gq = 0; gq < 8; ++gq)
                    *reinterpret_cast<float4*>(dst + c + 4 * gq) = make_float4(v[4 * gq], v[4 * gq + 1], v[4 * gq + 2], v[4 * gq + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------ host
// cudaFuncSetAttribute is per device: remember it per (kernel, device), not per process
template <typename K>
static int ensure_big_smem(K kernel, bool* done /* [64] */) {
    int dev = 0;
    CN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (!done[dev]) {
        CN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        done[dev] = true;
    }
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D row-major tensor [rows, cols] with row pitch ld (elements); box = [box_rows, box_cols], 128B swizzle
static int make_map(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* base, int64_t rows, int64_t cols,
                    int64_t ld, int box_cols, int box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * esize) & 15)) {
        set_error("tcgen05 GEMM operand must be 16-byte aligned with a 16-byte-multiple row pitch");
        return 2;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * esize)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld box=%dx%d)", (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols); return 1; }
    return 0;
}

// launches a CG = 2 instantiation as clusters of two CTAs
template <typename K, typename E>
static int launch_pair(K kernel, int grid, size_t smem, cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int Kd,
                       int BN, int n_tiles, int m_tiles, int stages, const E& epi, double* stats) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CN_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, M, N, Kd, BN, n_tiles, m_tiles, stages, epi, stats));
    return 0;
}

template <typename T>
static int run_nt(const cartnet_gemm_t& d, cudaStream_t st, double* stats, int* stats_blocks, bool allow_pair = true) {
    using TR = TcTraits<T>;
    constexpr int NP = TR::NP;
    const int slot = (int)sizeof(T);          // shared-memory bytes per operand element (pairs: hi + lo)
    CN_CHECK_ARG(d.K % TR::KB == 0, "tcgen05 gemm: K=%d must be a multiple of %d", d.K, TR::KB);
    if (NP == 2) {
        CN_CHECK_ARG((reinterpret_cast<uintptr_t>(d.A) & 255) == 0 && (reinterpret_cast<uintptr_t>(d.B) & 255) == 0 && d.lda % 64 == 0 && d.ldb % 64 == 0,
                     "tcgen05 gemm: bf16x3 operands need 256-byte aligned bases and row pitches that are multiples of 64");
    }
    // resident weight slice: largest BN with BN*K*slot <= 128 KB that divides N and leaves room for >= 2 A stages
    int BN = 0, stages = 0;
    size_t smem = 0;
    static const int bn_cap = getenv("CARTNET_NT_BN_CAP") ? atoi(getenv("CARTNET_NT_BN_CAP")) : 256;   // tuning knob (experiments)
    const bool stats_in_regs = stats && NP == 2;            // pairs: per-thread running sums, at most two column blocks per warp
    for (int cand : {256, 128, 64, 32}) {
        if (cand > bn_cap || (int64_t)cand * d.K * slot > 131072 || d.N % cand != 0 || (stats_in_regs && cand > 128)) continue;
        for (int ns : {6, 5, 4, 3, 2}) {
            const size_t need = 1024 + (size_t)cand * d.K * slot + (size_t)ns * NP * NT_A_PART_BYTES + NT_STG_BYTES + sizeof(NtBars) + 64 +
                                ((stats && !stats_in_regs) ? NT_STAT_BYTES : 0);
            if (need <= (size_t)227 * 1024) { BN = cand; stages = ns; smem = need; break; }
        }
        if (BN) break;
    }
    CN_CHECK_ARG(BN > 0, "tcgen05 gemm: no resident tile for N=%d K=%d", d.N, d.K);
    // CTA pairs (cta_group::2): each CTA keeps half of the pair's weight slice, so the pair covers twice the columns with
    // the same shared memory and A crosses the L2 -> SM fabric half as often. Edge-sized GEMMs of the 4-byte operand modes.
    const char* cg_str = getenv("CARTNET_NT_CG");                 // "1": never pair (A/B runs and the pair-vs-single tests; read per call)
    const int cg_env = cg_str ? atoi(cg_str) : 2;
    int CG = 1;
    if (slot == 4 && cg_env == 2 && allow_pair && d.M >= CARTNET_NT_PAIR_MIN_ROWS && BN < d.N) {
        for (int cand : {256, 128, 64}) {
            if (cand <= BN || cand > 2 * bn_cap || (int64_t)(cand / 2) * d.K * slot > 131072 || d.N % cand != 0) continue;
            for (int ns : {6, 5, 4, 3, 2}) {
                const size_t need = 1024 + (size_t)(cand / 2) * d.K * slot + (size_t)ns * NP * NT_A_PART_BYTES + NT_STG_BYTES + sizeof(NtBars) + 64 +
                                    ((stats && !stats_in_regs) ? NT_STAT_BYTES : 0);
                if (need <= (size_t)227 * 1024) { BN = cand; stages = ns; smem = need; CG = 2; break; }
            }
            if (CG == 2) break;
        }
    }
    const int n_tiles = d.N / BN, m_tiles = ceil_div(d.M, 128 * CG);
    CN_CHECK_ARG(n_tiles <= kNumSMs / CG, "tcgen05 gemm: too many N tiles (%d)", n_tiles);
    int units = (kNumSMs / CG / n_tiles) * n_tiles;                // CTAs, or CTA pairs
    if ((int64_t)units > (int64_t)m_tiles * n_tiles) units = m_tiles * n_tiles;
    const int grid = units * CG;
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, TR::DT, TR::TMA_ES, d.A, d.M, (int64_t)NP * d.K, NP * d.lda, 128 / TR::TMA_ES, 128);
    if (rc) return rc;
    rc = make_map(&tmB, TR::DT, TR::TMA_ES, d.B, d.N, (int64_t)NP * d.K, NP * d.ldb, 128 / TR::TMA_ES, BN / CG);
    if (rc) return rc;
    if (stats_blocks) *stats_blocks = (units / n_tiles) * CG * 4;
    static const int pf_env = getenv("CARTNET_NT_PREFETCH") ? atoi(getenv("CARTNET_NT_PREFETCH")) : 1;     // tuning knob (experiments)
    // CTA pairs read every A tile once, from one SM: the prefetch only adds DRAM reads there (ncu: +0.1..0.3 GB per launch,
    // step 28.1 -> 27.8 ms without it) except for the two-stage K >= 512 ring, which still gains 2 %
    if (pf_env && m_tiles > units / n_tiles && (CG == 1 || d.K >= 512)) stages |= 16;
    const char* uni_str = getenv("CARTNET_NT_UNIFORM");             // "0": always one gather0 load per row (A/B)
    if (!uni_str || atoi(uni_str) != 0) stages |= 32;
    const char* gpf_str = getenv("CARTNET_NT_GATHER_PREFETCH");     // L1 prefetch of the gathered rows, blocks ahead: "0" off, "1" (default), "2"
    const int gpf_mode = gpf_str ? atoi(gpf_str) : 1;               // first edge Linear at ADP-64: 0.812 / 0.774 / 0.781 ms
    if (gpf_mode == 1) stages |= 64;
    if (gpf_mode == 2) stages |= 128;
    int mask = 0;
    if (stats) mask |= EB_STATS;
    if (d.bias) mask |= EB_BIAS;
    if (d.gather0 && d.gather1) mask |= EB_GATHER;
    if (d.z_out) mask |= EB_ZOUT;
    if (d.act == CARTNET_ACT_SILU) mask |= EB_SILU;
    if (d.act == CARTNET_ACT_MUL_DSILU) mask |= EB_DSILU;
    if (d.resid) mask |= EB_RESID;
    if (d.out_f32) mask |= EB_OUTF;
    if (d.out_t) mask |= EB_OUTT;
    if ((d.gather0 != nullptr) != (d.gather1 != nullptr)) mask = -2;   // single gather: generic path
    const EpiParams<T> epi = make_epi<T>(d);
#define CN_NT_CASE(M_)                                                                                               \
    if (mask == (M_)) {                                                                                              \
        if constexpr (sizeof(T) == 4) {                                                                              \
            if (CG == 2) {                                                                                           \
                static bool attr2_done[64] = {};                                                                     \
                if (int rc_ = ensure_big_smem(tc_nt_kernel<T, (M_), 2>, attr2_done)) return rc_;                     \
                return launch_pair(tc_nt_kernel<T, (M_), 2>, grid, smem, st, tmA, tmB, d.M, d.N, d.K, BN, n_tiles, m_tiles, stages, epi, stats); \
            }                                                                                                        \
        }                                                                                                            \
        static bool attr_done[64] = {};                                                                              \
        if (int rc_ = ensure_big_smem(tc_nt_kernel<T, (M_)>, attr_done)) return rc_;                                 \
        tc_nt_kernel<T, (M_)><<<grid, NT_THREADS, smem, st>>>(tmA, tmB, d.M, d.N, d.K, BN, n_tiles, m_tiles, stages, epi, stats); \
        CN_LAUNCH_CHECK();                                                                                           \
        return 0;                                                                                                    \
    }
    CN_NT_CASE(EB_OUTT)                                                   // node projections P
    CN_NT_CASE(EB_BIAS | EB_GATHER | EB_ZOUT | EB_SILU | EB_OUTT)         // first Linear of both MLPs (per edge)
    CN_NT_CASE(EB_BIAS | EB_OUTF)                                         // second Linear of MLP_gate -> g (fp32: BatchNorm input)
    CN_NT_CASE(EB_BIAS | EB_OUTT)                                         // second Linear of MLP_aggr -> s (T)
    CN_NT_CASE(EB_BIAS | EB_OUTT | EB_STATS)                              // second Linear of MLP_gate -> centred g (T) + BatchNorm sums
    CN_NT_CASE(EB_DSILU | EB_OUTT)                                        // dgrad through the second Linears
    CN_NT_CASE(EB_RESID | EB_OUTF)                                        // dgrad to e / x with the residual
    CN_NT_CASE(EB_OUTF)                                                   // dgrad to e when no gradient enters e_out (last layer)
    CN_NT_CASE(EB_BIAS | EB_GATHER | EB_SILU | EB_OUTT)                   // inference (no backward): the pre-activations are not stored
    CN_NT_CASE(EB_BIAS | EB_SILU | EB_OUTT)
    CN_NT_CASE(EB_BIAS | EB_SILU | EB_OUTF | EB_OUTT)
    CN_NT_CASE(EB_BIAS | EB_SILU | EB_OUTF)
    CN_NT_CASE(EB_BIAS | EB_ZOUT | EB_SILU | EB_OUTT)                     // edge encoder, first Linear
    CN_NT_CASE(EB_BIAS | EB_ZOUT | EB_SILU | EB_OUTF | EB_OUTT)           // edge encoder, second Linear (bf16)
    CN_NT_CASE(EB_BIAS | EB_ZOUT | EB_SILU | EB_OUTF)                     // edge encoder, second Linear (tf32)
#undef CN_NT_CASE
    CN_CHECK_ARG(!stats, "tcgen05 gemm: fused column statistics need the bias + T-output epilogue");
    if (CG == 2) return run_nt<T>(d, st, stats, stats_blocks, false);      // CTA pairs exist for the specialised epilogues only
    {
        static bool attr_done[64] = {};
        if (int rc_ = ensure_big_smem(tc_nt_kernel<T, EPI_GENERIC>, attr_done)) return rc_;
        tc_nt_kernel<T, EPI_GENERIC><<<grid, NT_THREADS, smem, st>>>(tmA, tmB, d.M, d.N, d.K, BN, n_tiles, m_tiles, stages, epi, nullptr);
        CN_LAUNCH_CHECK();
    }
    return 0;
}

// stats: optional fp64 partial buffer [stats_blocks][2][N] (sum | sum of squares of the output columns)
int gemm_tc_nt(const cartnet_gemm_t& d, cudaStream_t st, double* stats, int* stats_blocks) {
    if (d.prec == CARTNET_PREC_BF16) return run_nt<__nv_bfloat16>(d, st, stats, stats_blocks);
    if (d.prec == CARTNET_PREC_BF16X3) return run_nt<bf16p_t>(d, st, stats, stats_blocks);
    return run_nt<tf32_t>(d, st, stats, stats_blocks);
}

struct TnPlan {
    int Nblk, n_blocks, m_blocks, kblks_total, kblks_per_split, splits;
};
static bool tn_plan(int prec, int M, int N, int64_t K, TnPlan* p) {
    const int boxw = prec == CARTNET_PREC_TF32 ? 32 : 64;           // mn elements per TMA box
    const int krows = prec == CARTNET_PREC_BF16 ? 64 : 32;          // k rows per pipeline stage (TcTraits::TN_KROWS)
    if (M % 128 != 0 || N % boxw != 0) return false;      // rows beyond M inside the last 256-row block are zero-filled by TMA
    p->Nblk = N <= 256 ? N : 256;
    if (N % p->Nblk != 0 || p->Nblk % 32 != 0 || p->Nblk % 16 != 0) return false;
    p->n_blocks = N / p->Nblk;
    p->m_blocks = ceil_div(M, TN_MBLK);
    p->kblks_total = (int)ceil_div64(K > 0 ? K : 1, krows);
    const int problems = p->n_blocks * p->m_blocks;
    int s = kNumSMs / problems;
    const int by_k = p->kblks_total / (1024 / krows);   // at least 1024 k rows per split: partials must not outweigh the operands
    if (s > by_k) s = by_k;
    if (s < 1) s = 1;
    p->kblks_per_split = ceil_div(p->kblks_total, s);
    p->splits = ceil_div(p->kblks_total, p->kblks_per_split);
    return true;
}

int64_t gemm_tc_tn_workspace(int prec, int M, int N, int64_t K) {
    TnPlan p;
    if (!tn_plan(prec, M, N, K, &p)) return 16;
    return (int64_t)p.splits * M * N * (int64_t)sizeof(float);
}

template <typename T>
static int run_tn(int prec, int M, int N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb, const TnDst& C,
                  int64_t ldc, float* ws, cudaStream_t st) {
    using TR = TcTraits<T>;
    constexpr int NP = TR::NP;
    TnPlan p;
    CN_CHECK_ARG(tn_plan(prec, M, N, K, &p), "tcgen05 gemm_tn: unsupported shape M=%d N=%d (need M %% 128 == 0, N %% %d == 0)", M, N, TR::KB);
    if (NP == 2) {
        CN_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 255) == 0 && (reinterpret_cast<uintptr_t>(B) & 255) == 0 && lda % 64 == 0 && ldb % 64 == 0,
                     "tcgen05 gemm_tn: bf16x3 operands need 256-byte aligned bases and row pitches that are multiples of 64");
    }
    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, TR::DT, TR::TMA_ES, A, K, (int64_t)NP * M, NP * lda, TR::KB, TR::TN_KROWS, TR::MN_SWIZZLE);
    if (rc) return rc;
    rc = make_map(&tmB, TR::DT, TR::TMA_ES, B, K, (int64_t)NP * N, NP * ldb, TR::KB, TR::TN_KROWS, TR::MN_SWIZZLE);
    if (rc) return rc;
    const int box_bytes = TR::TN_KROWS * 128;
    const size_t stage_bytes = (size_t)NP * ((size_t)(TN_MBLK / TR::KB) * box_bytes + (size_t)(p.Nblk / TR::KB) * box_bytes);
    const size_t smem = 1024 + TN_STAGES * stage_bytes + sizeof(TnBars) + 64;
    static bool attr_done[64] = {};
    if (int rc_ = ensure_big_smem(tc_tn_kernel<T>, attr_done)) return rc_;
    tc_tn_kernel<T><<<dim3(p.splits, p.m_blocks * p.n_blocks), TN_THREADS, smem, st>>>(
        tmA, tmB, K, M, N, p.Nblk, p.n_blocks, p.kblks_total, p.kblks_per_split, ws);
    CN_LAUNCH_CHECK();
    return launch_splitk_reduce(ws, p.splits, M, N, C, ldc, st);
}

int gemm_tc_tn(int prec, int M, int N, int64_t K, const void* A, int64_t lda, const void* B, int64_t ldb, const TnDst& C,
               int64_t ldc, float* ws, int64_t ws_bytes, cudaStream_t st) {
    if (K <= 0) {
        for (int b = 0; b < M / C.rows_per_blk; ++b)
            CN_CUDA(cudaMemset2DAsync(C.c[b], ldc * sizeof(float), 0, (size_t)N * sizeof(float), C.rows_per_blk, st));
        return 0;
    }
    if (prec == CARTNET_PREC_BF16) return run_tn<__nv_bfloat16>(prec, M, N, K, A, lda, B, ldb, C, ldc, ws, st);
    if (prec == CARTNET_PREC_BF16X3) return run_tn<bf16p_t>(prec, M, N, K, A, lda, B, ldb, C, ldc, ws, st);
    return run_tn<tf32_t>(prec, M, N, K, A, lda, B, ldb, C, ldc, ws, st);
}

}  // namespace cartnet
