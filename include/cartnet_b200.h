/* cartnet_b200 -- C ABI of the B200-native CartNet hot path (libcartnet_b200.so).
 *
 * Everything the Python host code (cartnet_b200/) calls on the device goes through the
 * entry points below: plain pointers, sizes and PODs, no torch types, no C++ exceptions.
 * All pointers are DEVICE pointers unless marked "host". Every tensor is allocated and
 * owned by the caller; the library never allocates or frees device memory. Calls are
 * stateless and stream-ordered on `stream` (a cudaStream_t), so one process per GPU is safe.
 * Return value: 0 on success, non-zero on error; cartnet_last_error() (thread-local)
 * describes the last failure.
 *
 * The reference (imatge-upc/CartNet) is pure Python and has no FFI; each group cites the
 * reference lines (relative to /root/reference) whose eager-op chain it replaces.
 *
 * "T" below is the GEMM operand type selected by `prec`:
 *   CARTNET_PREC_FP32 : T = float,         GEMMs on the fp32 SIMT pipe  (1e-5 parity path)
 *   CARTNET_PREC_BF16 : T = __nv_bfloat16, GEMMs on tcgen05/TMEM fed by TMA (2e-3 path)
 *   CARTNET_PREC_TF32 : T = float,         GEMMs on tcgen05 kind::tf32 reading the fp32 tensors directly
 *   CARTNET_PREC_BF16X3 : T = 4-byte opaque slot; split-precision tensor-core mode. Every run of 64 consecutive elements
 *                       (256 bytes, 256-byte aligned) stores 64 bf16 high parts followed by 64 bf16 low parts
 *                       (value = hi + lo, ~16 mantissa bits); GEMMs are three tcgen05 kind::f16 MMAs per product
 *                       (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM). Bases must be 256-byte aligned, row pitches
 *                       and column offsets multiples of 64 elements. This is the tensor-core mode that holds the 2e-3
 *                       tolerance in TRAINING mode (edge BatchNorm amplifies operand rounding ~15x).
 *                       Tensors that are never contracted are NOT pair-typed in this mode: gathered addends are plain
 *                       fp32, stored pre-activations (z_out / z_in / dsilu_mul's z) are fp16 with saturating stores.
 */
#ifndef CARTNET_B200_H
#define CARTNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cartnet_stream_t; /* cudaStream_t */

enum { CARTNET_PREC_FP32 = 0, CARTNET_PREC_BF16 = 1, CARTNET_PREC_TF32 = 2, CARTNET_PREC_BF16X3 = 3 };
/* BF16X3 GEMMs with at least this many rows run as CTA pairs (tcgen05 cta_group::2: one M = 256 MMA per two SMs, each SM
 * keeping half of the pair's weight slice); the layer picks its launch shapes accordingly. Results do not depend on it. */
#define CARTNET_NT_PAIR_MIN_ROWS 32768
enum { CARTNET_ACT_NONE = 0, CARTNET_ACT_SILU = 1, CARTNET_ACT_MUL_DSILU = 2 };

int cartnet_version(void);
const char* cartnet_last_error(void);
/* 1 if the visible device is sm_100 (B200); the bf16 path needs it. */
int cartnet_device_ok(int device);
/* Number of kernels this library has launched in this process so far (bench.py's gpu_launches). */
int64_t cartnet_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Periodic radius graph -- replaces dataset/utils.py:57-237 (radius_graph_pbc) and the
 * callers' post-processing dataset/figshare_dataset.py:67-68.
 * Edge set AND order are bit-exact: (dst, src, cell) lexicographic, 1e-4 < d^2 <= r^2,
 * fp32 arithmetic without FMA contraction in the reference's operation order.
 * ------------------------------------------------------------------------------------- */

/* rep_k = ceil(radius * ||(a_i x a_j)/V||) per crystal (utils.py:135-156).
 * reps_out[B,3] int32. If reps_max (int32[3], pre-zeroed) is non-null it receives the max over the
 * batch (utils.py:163). pbc_mask bit k set = axis k periodic. */
int cartnet_nlist_reps(const float* cell, int32_t num_crystals, float radius, int32_t pbc_mask,
                       int32_t* reps_out, int32_t* reps_max, cartnet_stream_t stream);

/* Pass 1: row_count[n] = number of edges whose destination is atom n.
 * crystal_ptr[B+1]: first atom of each crystal; node_crystal[N]: crystal of each atom.
 * reps[B,3] per crystal when reps_stride = 3, or one shared int32[3] (batch max) when reps_stride = 0.
 * radius_sq = (float)(radius*radius) rounded from the double product like torch.le does (utils.py:202). */
int cartnet_nlist_count(const float* pos, const float* cell, const int32_t* crystal_ptr,
                        const int32_t* node_crystal, int32_t num_nodes, float radius, float radius_sq,
                        const int32_t* reps, int32_t reps_stride, int32_t* row_count,
                        cartnet_stream_t stream);

/* out[0] = 0, out[i+1] = sum_{j<=i} in[j]  (n+1 outputs). Single-launch, deterministic. */
int cartnet_exclusive_scan_i32(const int32_t* in, int32_t n, int32_t* out, cartnet_stream_t stream);

/* Pass 2: writes the edges of row n at [row_ptr[n], row_ptr[n+1]).
 * edge_index[2,E] int64 (row 0 = src j, row 1 = dst i; utils.py:235), unit_cell[E,3] f32,
 * dist[E] = sqrt(d^2), direction[E,3] = pos[dst] - (pos[src] + offset)  (utils.py:193-198,237).
 * Optional (may be null): cart_dist[E] = ||direction||, cart_dir[E,3] = direction/max(||.||,1e-12)
 * (figshare_dataset.py:67-68); src32/dst32[E] int32 copies for the layer kernels. */
int cartnet_nlist_fill(const float* pos, const float* cell, const int32_t* crystal_ptr,
                       const int32_t* node_crystal, int32_t num_nodes, float radius, float radius_sq,
                       const int32_t* reps, int32_t reps_stride, const int32_t* row_ptr,
                       int64_t* edge_index, int64_t num_edges, float* unit_cell, float* dist,
                       float* direction, float* cart_dist, float* cart_dir, int32_t* src32,
                       int32_t* dst32, cartnet_stream_t stream);

/* Cell-list variant of the two passes above (same outputs, same order, bit for bit) for crystals that are large against
 * the radius: atoms are binned on a grid in wrapped fractional coordinates with bins at least r |b_k| wide, a warp per
 * destination visits the 27 surrounding bins, tests every candidate with the reference's exact fp32 sequence and restores
 * the reference's row order ((source, cell index) ascending, utils.py:116-123,166-170) by a rank sort of the row's keys.
 * Crystals with fewer than 64 bins (ADP / JARVIS / MP sized cells) keep the all-pairs scan inside the same launches, so a
 * mixed batch is one call. workspace: >= cartnet_nlist_cells_workspace(num_nodes, num_crystals) bytes, filled by _build and
 * read by _count / _fill. */
int64_t cartnet_nlist_cells_workspace(int32_t num_nodes, int32_t num_crystals);
int cartnet_nlist_cells_build(const float* pos, const float* cell, const int32_t* crystal_ptr,
                              const int32_t* node_crystal, int32_t num_nodes, int32_t num_crystals, float radius,
                              const int32_t* reps, int32_t reps_stride, void* workspace, cartnet_stream_t stream);
int cartnet_nlist_cells_count(const float* pos, const float* cell, const int32_t* crystal_ptr,
                              const int32_t* node_crystal, int32_t num_nodes, int32_t num_crystals, float radius,
                              float radius_sq, const int32_t* reps, int32_t reps_stride, const void* workspace,
                              int32_t* row_count, cartnet_stream_t stream);
int cartnet_nlist_cells_fill(const float* pos, const float* cell, const int32_t* crystal_ptr,
                             const int32_t* node_crystal, int32_t num_nodes, int32_t num_crystals, float radius,
                             float radius_sq, const int32_t* reps, int32_t reps_stride, const void* workspace,
                             const int32_t* row_ptr, int64_t* edge_index, int64_t num_edges, float* unit_cell,
                             float* dist, float* direction, float* cart_dist, float* cart_dir, int32_t* src32,
                             int32_t* dst32, cartnet_stream_t stream);

/* kNN neighbour cap -- replaces get_max_neighbors_mask (dataset/utils.py:240-360, call at :215-233).
 * For every destination row of a dst-sorted graph: keep[e] = 1 iff the row has <= threshold edges, or (non-strict)
 * d2[e] <= (threshold+1)-th smallest d2 of the row + tolerance (degenerate neighbours stay together), or (strict)
 * e is among the `threshold` smallest (ties by edge order). d2 is recomputed from `direction` with the reference's
 * rounding ((dx*dx+dy*dy)+dz*dz, no FMA). d2_scratch: float[E]; new_row_count[n] = kept edges of row n. */
int cartnet_nlist_knn_mask(const float* direction, const int32_t* row_ptr, int32_t num_nodes, int32_t threshold,
                           float tolerance, int32_t strict, float* d2_scratch, uint8_t* keep,
                           int32_t* new_row_count, cartnet_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Graph plan for the layer kernels -- replaces PyG MessagePassing's lift/scatter indexing
 * (call at models/cartnet.py:218-221) and torch_scatter.scatter (cartnet.py:259).
 * ------------------------------------------------------------------------------------- */

/* edge_index[2,E] int64 -> src32/dst32 int32; flags[0] = 1 if dst is non-decreasing, flags[1] = 1 if
 * any index is outside [0, num_nodes). */
int cartnet_graph_split(const int64_t* edge_index, int64_t num_edges, int32_t num_nodes,
                        int32_t* src32, int32_t* dst32, int32_t* flags, cartnet_stream_t stream);

/* Counting-sort CSR of `keys[E]` (values in [0,num_nodes)): ptr[num_nodes+1], perm[E] = edge ids
 * grouped by key, ascending edge id inside a group (deterministic). `cursor` is int32[num_nodes]
 * scratch. When keys are already sorted perm is the identity. */
int cartnet_graph_csr(const int32_t* keys, int64_t num_edges, int32_t num_nodes, int32_t* ptr,
                      int32_t* perm, int32_t* cursor, cartnet_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Edge featuriser -- replaces models/utils.py:56-61,87-91 and the cat at cartnet.py:159.
 * feat[E, ld] (T): cols 0..num_rbf-1 = cut(d)*exp(-beta_k (exp(-alpha d) - mu_k)^2), then (unless
 * invariant) 3 cols cart_dir, then (if ld leaves room) ONE column of ones -- the matching column of the
 * zero-padded weight is 0, so it only serves the backward: d(bias) = that column of dz^T feat -- and the
 * remaining cols up to ld zero.
 * ------------------------------------------------------------------------------------- */
int cartnet_edge_features(const float* cart_dist, const float* cart_dir, const float* means,
                          const float* betas, int32_t num_rbf, float cutoff_upper, int32_t invariant,
                          int64_t num_edges, void* feat, int32_t ld, int32_t prec,
                          cartnet_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Dense contractions -- replace the addmm/mm chains of cartnet.py:133-136,187-196 and their
 * autograd backward. The gathered epilogue implements the split of the first Linear
 * W1 = [W_i | W_j | W_e]: per-node projections are added per edge.
 * ------------------------------------------------------------------------------------- */
typedef struct cartnet_gemm {
    int32_t prec;            /* CARTNET_PREC_* */
    int32_t M, N, K;         /* C[M,N] = A[M,K] * B[N,K]^T ; bf16: K % 64 == 0, N % 64 == 0 */
    const void* A;           /* T, row-major, leading dimension lda (elements) */
    int64_t lda;
    const void* B;           /* T, [N,K] row-major (PyTorch Linear weight layout), ld = ldb */
    int64_t ldb;
    /* epilogue, applied in this order; null pointer = step skipped */
    const float* bias;       /* v += bias[col] */
    const void* gather0;     /* v += gather0[gidx0[row]*ldg + col]      (T; plain fp32 in the BF16X3 mode: only added, never contracted) */
    const int32_t* gidx0;
    const void* gather1;     /* v += gather1[gidx1[row]*ldg + col]      (T) */
    const int32_t* gidx1;
    int64_t ldg;
    void* z_out;             /* z_out[row*ldz + col] = v                (T; fp16 words, saturating, in the BF16X3 mode) */
    int64_t ldz;
    int32_t act;             /* CARTNET_ACT_*: none | v = silu(v) | v *= silu'(z_in[row*ldzin+col]) */
    int32_t _pad;
    const void* z_in;        /* T (fp16 in the BF16X3 mode) */
    int64_t ldzin;
    const float* resid;      /* v += resid[row*ldr + col] */
    int64_t ldr;
    float* out_f32;          /* out_f32[row*ldo + col] = v */
    int64_t ldo;
    void* out_t;             /* out_t[row*ldt + col] = (T) v */
    int64_t ldt;
} cartnet_gemm_t;

int cartnet_gemm(const cartnet_gemm_t* desc /* host */, cartnet_stream_t stream);

/* C[M,N] (fp32, ldc) = sum_k A[k, 0:M]^T * B[k, 0:N]  -- the weight-gradient contraction over
 * K = edges (or nodes). A: [K, lda] T, B: [K, ldb] T. Split-K with a fixed-order second pass, so the
 * result is deterministic. workspace: fp32, at least cartnet_gemm_tn_workspace(...) bytes. */
int64_t cartnet_gemm_tn_workspace(int32_t prec, int32_t M, int32_t N, int64_t K);
int cartnet_gemm_tn(int32_t prec, int32_t M, int32_t N, int64_t K, const void* A, int64_t lda,
                    const void* B, int64_t ldb, float* C, int64_t ldc, float* workspace,
                    int64_t workspace_bytes, cartnet_stream_t stream);

/* GEMM (bias + T output only) whose per-column output statistics come with it: mean/var of C's columns over the M rows
 * as cartnet_colstats would compute them from out_t (same shift / running-statistics semantics). In the tensor-core
 * modes the sums are accumulated in the epilogue (no extra pass over C, taken before the rounding to T); in fp32 mode
 * the statistics pass runs after the GEMM. partial: >= cartnet_colstats_workspace(N) bytes. */
int cartnet_gemm_colstats(const cartnet_gemm_t* d /* host */, const float* shift, float* mean, float* var,
                          float* running_mean, float* running_var, float momentum, double* partial,
                          cartnet_stream_t stream);

/* Same, with the M output rows split into num_blocks (1..4) equal blocks that are written to separate bases
 * C_blocks[b] (host array of device pointers, each [M/num_blocks, N] with pitch ldc): the gradient of a row-packed
 * weight ([G1_e;A1_e], [G1_i;A1_i;G1_j;A1_j]) lands directly in the reference's [D,3D] parameter layout. */
int cartnet_gemm_tn_blocks(int32_t prec, int32_t M, int32_t N, int64_t K, const void* A, int64_t lda,
                           const void* B, int64_t ldb, float* const* C_blocks, int32_t num_blocks, int64_t ldc,
                           float* workspace, int64_t workspace_bytes, cartnet_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Column statistics / BatchNorm pieces -- replace nn.BatchNorm1d over E rows (cartnet.py:198,238)
 * and over N rows (cartnet.py:199,269). Sums are accumulated in fp64 in a fixed order.
 * ------------------------------------------------------------------------------------- */

/* mean[C], var[C] (biased) of x[rows, C] (T when x_is_t, else fp32; pitch ld). If running_mean/var non-null they are
 * updated in place like nn.BatchNorm1d in train mode: r = (1-momentum) r + momentum * stat, with the UNBIASED
 * variance; `shift` (nullable, [C]) is added to the mean in that update only (x was stored centred: x = true - shift).
 * partial: fp64 scratch, >= cartnet_colstats_workspace(C) bytes. */
int64_t cartnet_colstats_workspace(int32_t C);
int cartnet_colstats(const void* x, int32_t x_is_t, int32_t prec, int64_t rows, int32_t C, int64_t ld,
                     const float* shift, float* mean, float* var, float* running_mean, float* running_var,
                     float momentum, double* partial, cartnet_stream_t stream);

/* Centre for the gate pre-activation g = H_g G2^T + bg2 (cartnet.py:195-196) so that g - center can be stored in T:
 * BatchNorm removes any per-column shift exactly, so an approximate mean is enough. Training: center = bg2 +
 * G2 mean_s(H_g) over <= 4096 rows of H_g (T, pitch ldh) sampled at a fixed stride; eval: center = running_mean.
 * Outputs bias_c = bg2 - center (the bias of the centred GEMM) and center, both [D]. hsum: [D] fp32 scratch. */
int cartnet_gate_center(const void* H_g, int64_t ldh, int64_t num_edges, int32_t D, const float* G2,
                        const float* bg2, const float* running_mean, int32_t training, int32_t prec,
                        float* bias_c, float* center, float* hsum, double* partial, cartnet_stream_t stream);

/* out[C] (fp32) = column sums of x[rows, C] (T or fp32 selected by x_is_t/prec). */
int cartnet_colsum(const void* x, int32_t x_is_t, int32_t prec, int64_t rows, int32_t C, int64_t ld,
                   float* out, double* partial, cartnet_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused edge / node passes of CartNet_layer (cartnet.py:230-274) and their backward.
 * D = dim_in (multiple of 4). Edges are dst-sorted; row_ptr is the dst CSR.
 * bn_*: mean/var are the batch statistics (train) or the running ones (eval), eps = 1e-5.
 * ------------------------------------------------------------------------------------- */

/* Forward edge pass:  ghat = BN(g); sig = env(dist) * sigmoid(ghat); e_out = e + sig;
 * m[i] = sum_{edges -> i, CSR order} sig * s   (deterministic, no atomics).
 * env(d) = 0.5 (cos(pi d / radius) + 1) (d < radius) when use_envelope else 1.
 * g_t and s_t are T; g_t may be stored centred (cartnet_gate_center) with bn_mean = the mean of the stored values
 * (null = 0: eval mode, centred on the running mean). e_out_t (T shadow for the next layer's GEMM operand) may be
 * null. gn_t (T, may be null) receives the normalised pre-activation (g - mean) * rstd: with s_t it is all the
 * backward pass needs, so g_t itself is scratch. */
int cartnet_edge_gate_aggregate(const void* g_t, const void* s_t, const float* e, const float* dist,
                                const int32_t* row_ptr, int32_t num_nodes, int64_t num_edges, int32_t D,
                                const float* bn_mean, const float* bn_var, const float* bn_weight,
                                const float* bn_bias, float eps, float radius, int32_t use_envelope,
                                float* e_out, void* e_out_t, void* gn_t, int32_t prec, float* m,
                                cartnet_stream_t stream);

/* Forward node pass: x_out = silu(BN2(m)) + x ; x_out_t optional T shadow. */
int cartnet_node_update(const float* m, const float* x, int32_t num_nodes, int32_t D,
                        const float* bn_mean, const float* bn_var, const float* bn_weight,
                        const float* bn_bias, float eps, float* x_out, void* x_out_t, int32_t prec,
                        cartnet_stream_t stream);

/* Backward node pass, step 1: dy = dx_out * silu'(BN2(m)); sums[0:D] = sum_n dy,
 * sums[D:2D] = sum_n dy * yhat   (yhat = (m-mean)*rstd). fp64 partial scratch as in colstats. */
int cartnet_node_update_bwd_reduce(const float* dx_out, const float* m, int32_t num_nodes, int32_t D,
                                   const float* bn_mean, const float* bn_var, const float* bn_weight,
                                   const float* bn_bias, float eps, float* sums, double* partial,
                                   cartnet_stream_t stream);
/* step 2: dm = weight*rstd*(dy - [train](sum_dy/N + yhat*sum_dy_yhat/N)); writes dm[N,D].
 * training = 0 gives the affine (eval-mode) backward. */
int cartnet_node_update_bwd_apply(const float* dx_out, const float* m, int32_t num_nodes, int32_t D,
                                  const float* bn_mean, const float* bn_var, const float* bn_weight,
                                  const float* bn_bias, float eps, const float* sums, int32_t training,
                                  float* dm, cartnet_stream_t stream);

/* Backward edge pass, step 1 (per edge, channel), from the saved gn_t / s_t (both T):
 * ghat = gn * weight + bias; sig = env * sigmoid(ghat); dmd = dm[dst];  ds = sig * dmd  -> ds_t (T);
 * dghat = (de_out + s * dmd) * env * sigmoid'(ghat) -> dghat_t (T);
 * sums[0:D] = sum_e dghat, sums[D:2D] = sum_e dghat * gn, sums[2D:3D] = sum_e ds (sums has 3D entries; the sums
 * are taken before the rounding to T).
 * de_out may be null (no gradient flows into e_out, e.g. the last layer: the heads read only x).
 * g_var != null: `gn_t` holds the stored, centred pre-activation g (what cartnet_gemm_colstats wrote) and
 * gn = (g - g_mean) * rsqrt(g_var + eps) is formed on the fly (g_mean null = 0: eval mode, g centred on the running
 * mean), so that the forward pass need not write a normalised copy. g_var null: `gn_t` is already normalised. */
int cartnet_edge_gate_bwd_reduce(const void* gn_t, const void* s_t, const float* dist, const int32_t* dst32,
                                 const float* de_out, const float* dm, int64_t num_edges, int32_t D,
                                 const float* bn_weight, const float* bn_bias, float radius,
                                 int32_t use_envelope, void* ds_t, void* dghat_t, int32_t prec, float* sums,
                                 double* partial, const float* g_mean, const float* g_var, float eps,
                                 cartnet_stream_t stream);
/* step 2: dg = weight*rstd*(dghat - [train](sum/E + gn*sum2/E)) -> dg_t (T). input_is_g != 0: `gn_t` holds the
 * stored g as above (normalised with g_mean, bn_var, eps). */
int cartnet_edge_gate_bwd_apply(const void* gn_t, const void* dghat_t, int64_t num_edges, int32_t D,
                                const float* bn_var, const float* bn_weight, float eps, const float* sums,
                                int32_t training, void* dg_t, int32_t prec, const float* g_mean,
                                int32_t input_is_g, cartnet_stream_t stream);

/* out[n, 0:C] = sum over CSR row n of x[perm[k], 0:C] (perm may be null = identity). x is T,
 * out is T (out_is_t=1) or fp32. Used for d(P_i) (dst CSR) and d(P_j) (src CSR) -- the transpose of the
 * lifts at cartnet.py:218-221. Fixed order => deterministic. */
int cartnet_segment_sum(const void* x, int64_t ldx, const int32_t* ptr, const int32_t* perm,
                        int32_t num_nodes, int32_t C, void* out, int64_t ldo, int32_t out_is_t,
                        int32_t prec, cartnet_stream_t stream);

/* Both transposed lifts in one launch: out[n, 0:C] = dst-CSR (row_ptr, identity order) sum, out[n, C:2C] = src-CSR
 * (col_ptr through perm_src) sum of the same x. The second read of each row of x is served by L2 (see the kernel). */
int cartnet_segment_sum_pair(const void* x, int64_t ldx, const int32_t* row_ptr, const int32_t* col_ptr,
                             const int32_t* perm_src, int32_t num_nodes, int32_t C, void* out, int64_t ldo,
                             int32_t out_is_t, int32_t prec, cartnet_stream_t stream);

/* y_t = (T)(dy * silu'(z)) elementwise over [rows, C]; dy fp32 (ld_dy), z T (ldz), y T (ldy). Optional colsum [C]:
 * the column sums of y (the bias gradient of the Linear in front of the SiLU) taken in the same pass
 * (partial >= cartnet_colstats_workspace(C) bytes, C/4 a power of two <= 256); null = plain elementwise pass. */
int cartnet_dsilu_mul(const float* dy, int64_t ld_dy, const void* z, int64_t ldz, void* y, int64_t ldy,
                      int64_t rows, int32_t C, int32_t prec, float* colsum, double* partial,
                      cartnet_stream_t stream);

/* dst_t = (T) src  over [rows, C]. */
int cartnet_cast_rows(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int32_t C,
                      int32_t prec, cartnet_stream_t stream);

/* dst (fp32) = value of src_t over [rows, C]: the inverse view of cartnet_cast_rows (exact; in the bf16x3 mode hi + lo). */
int cartnet_uncast_rows(const void* src_t, int64_t lds, float* dst, int64_t ldd, int64_t rows, int32_t C,
                        int32_t prec, cartnet_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Whole-layer entry points -- CartNet_layer.forward (models/cartnet.py:204-274) and its backward as ONE call
 * each: the host issues the same kernels as the primitives above in a fixed order, without returning to the
 * caller's language in between. All buffers (saved activations, scratch, outputs) are caller-allocated.
 * D = dim_in; T as above. Weight layout of the reference is kept: G1/A1 = MLP_gate[0]/MLP_aggr[0].weight [D,3D]
 * with columns [x_i | x_j | e] (cartnet.py:237), G2/A2 = MLP_*[2].weight [D,D].
 * ------------------------------------------------------------------------------------- */
typedef struct cartnet_layer {
    int32_t prec, training, use_envelope, D;
    int32_t num_nodes, _pad0;
    int64_t num_edges;
    float radius, eps, momentum1, momentum2;
    /* graph (dst-sorted edges) */
    const int32_t *src32, *dst32, *row_ptr, *col_ptr, *perm_src;
    const float* dist;                               /* [E] */
    /* layer inputs: fp32 residual streams and their T-typed operand copies (may alias when T = float) */
    const float *x, *e;
    const void *x_t, *e_t;
    /* parameters (fp32, reference layout) */
    const float *G1, *A1, *bg1, *ba1, *G2, *A2, *bg2, *ba2, *bn1_w, *bn1_b, *bn2_w, *bn2_b;
    float *bn1_rm, *bn1_rv, *bn2_rm, *bn2_rv;        /* running statistics, updated in place when training */
    /* T-typed packed weights, filled by cartnet_layer_pack_weights:
     * W1n [4D,D] = [G1_i;A1_i;G1_j;A1_j], W1e [2D,D] = [G1_e;A1_e], G2t/A2t [D,D] copies, and the transposes
     * W1nT [D,4D], W1eT [D,2D], G2T, A2T [D,D] used by the dgrad GEMMs; b1 [2D] = [bg1;ba1] fp32 */
    void *W1n_t, *W1e_t, *G2_t, *A2_t, *W1nT_t, *W1eT_t, *G2T_t, *A2T_t;
    float* b1;
    /* forward: saved activations and outputs */
    void *P, *Z, *H;                                 /* T: [N,4D], [E,2D], [E,2D]; BF16X3: P plain fp32, Z fp16 */
    void* g_t;                                       /* T [E,D]: centred gate pre-activation (scratch after the forward pass) */
    float *center, *bias_c, *hsum;                   /* [D] each: cartnet_gate_center outputs / scratch */
    float* m;                                        /* [N,D] */
    void *s_t, *gn_t;                                /* T [E,D]: MLP_aggr output (saved); gn_t: optional normalised copy of g_t -- null: g_t itself is kept for backward */
    float *mean1, *var1, *mean2, *var2;              /* [D] statistics used (batch or running) */
    float *x_out, *e_out;                            /* [N,D], [E,D] */
    void *x_out_t, *e_out_t;                         /* T copies for the next layer (null when T = float) */
    /* backward: inputs, scratch, outputs */
    const float *dx_out, *de_out;                    /* de_out may be null: treated as zero without being read */
    float* dm;                                       /* [N,D] */
    void *ds_t, *dg_t;                               /* T [E,D] */
    void* dghat_t;                                   /* T [E,D] */
    void *dZ, *dP;                                   /* T [E,2D], [N,4D] */
    float *sums1, *sums2;                            /* [3D], [2D] */
    float *dx_in, *de_in;                            /* [N,D], [E,D] */
    float *dG1, *dA1, *dbg1, *dba1, *dG2, *dA2, *dbg2, *dba2, *dbn1_w, *dbn1_b, *dbn2_w, *dbn2_b;
    /* workspaces */
    double* partial;                                 /* >= cartnet_colstats_workspace(2D) */
    float* splitk;                                   /* >= cartnet_layer_splitk_bytes(...) */
    int64_t splitk_bytes;
} cartnet_layer_t;

int64_t cartnet_layer_splitk_bytes(int32_t prec, int32_t D, int32_t num_nodes, int64_t num_edges);
int cartnet_layer_pack_weights(const cartnet_layer_t* L /* host */, cartnet_stream_t stream);
int cartnet_layer_fwd(const cartnet_layer_t* L /* host */, cartnet_stream_t stream);
int cartnet_layer_bwd(const cartnet_layer_t* L /* host */, cartnet_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Cholesky head tail -- replaces models/cartnet.py:293-305 after the head's first Linear + SiLU
 * (SURVEY.md 8(f)3). h [n, Dh] fp32 (row pitch ldh) is the SiLU output for the n non-H atoms.
 *   fwd: p6 [n,6] = h W1^T + b1 (saved for backward);  d = softplus(p6[:, 0:3]);
 *        L = [[d0,p3,p4],[0,d1,p5],[0,0,d2]] (upper triangular, cartnet.py:296-301);  U [n,3,3] = L^T L.
 *   bwd: dU [n,3,3] (any, not necessarily symmetric) -> dh [n, Dh] (row pitch lddh), dW1 [6, Dh], db1 [6].
 *        The reduction over atoms is fixed-order (deterministic); partial >= cartnet_cholesky_head_workspace bytes.
 * Dh % 4 == 0, Dh <= 512. fp32 arithmetic in every precision mode (node-side, negligible cost).
 * ------------------------------------------------------------------------------------- */
int64_t cartnet_cholesky_head_workspace(int32_t n, int32_t Dh);
int cartnet_cholesky_head_fwd(const float* h, int64_t ldh, const float* W1, const float* b1, int32_t n, int32_t Dh,
                              float* p6, float* U, cartnet_stream_t stream);
int cartnet_cholesky_head_bwd(const float* dU, const float* h, int64_t ldh, const float* p6, const float* W1, int32_t n,
                              int32_t Dh, float* dh, int64_t lddh, float* dW1, float* db1, float* partial,
                              cartnet_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Device-side batch assembly -- replaces PyG Batch.from_data_list / DataLoader collation on the host followed by
 * batch.to("cuda:0") (loader/loader.py:114-124, train/train.py:169); SURVEY.md 8(f)1.
 * The data set is resident in HBM as per-field blobs (each the concatenation of all crystals). A batch = `ids`
 * (crystal ids, device int32 [num_selected]). Every output field is a segmented gather of the selected crystals'
 * ranges; index-valued fields get the batch's cumulative node / edge offset added (what PyG does to edge_index).
 *   blob_ptr  [4][blob_ptr_pitch] int32 (device): first element of every crystal in the blobs, per kind
 *   batch_ptr [4][num_selected+1] int32 (device): first element of every selected crystal in the batch, per kind
 *   kinds: 0 = per node, 1 = per edge, 2 = per non-H atom, 3 = per graph;  totals[kind] = elements in the batch (host).
 * One launch assembles all fields. cartnet_collate_close_csr writes the closing entries ptr[num_nodes] = num_edges.
 * ------------------------------------------------------------------------------------- */
enum { CARTNET_COLLATE_PER_NODE = 0, CARTNET_COLLATE_PER_EDGE = 1, CARTNET_COLLATE_PER_NONH = 2, CARTNET_COLLATE_PER_GRAPH = 3 };
enum {
    CARTNET_COLLATE_COPY = 0,                 /* dst[i] = src[s]   (elem_bytes = 1 or a multiple of 4) */
    CARTNET_COLLATE_I32_PLUS_NODE = 1,        /* int32: dst[i] = src[s] + node offset of the crystal in the batch */
    CARTNET_COLLATE_I32_PLUS_EDGE = 2,        /* int32: dst[i] = src[s] + edge offset of the crystal in the batch */
    CARTNET_COLLATE_I32_TO_I64_PLUS_NODE = 3, /* int32 blob -> int64 output + node offset (edge_index rows, non_H_index) */
    CARTNET_COLLATE_SLOT_I64 = 4              /* int64: dst[i] = position of the crystal in the batch (the `batch` vector) */
};
typedef struct cartnet_collate_field {
    const void* src;
    void* dst;
    int32_t kind, op, elem_bytes, _pad;
} cartnet_collate_field_t;
int cartnet_collate(const cartnet_collate_field_t* fields /* host */, int32_t num_fields, const int32_t* ids,
                    int32_t num_selected, const int32_t* blob_ptr, int64_t blob_ptr_pitch, const int32_t* batch_ptr,
                    const int64_t* totals /* host [4] */, cartnet_stream_t stream);
int cartnet_collate_close_csr(int32_t* row_ptr, int32_t* col_ptr, int64_t num_nodes, int64_t num_edges,
                              cartnet_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Loss pair of train/metrics.py:15-28 (nn.L1Loss and nn.MSELoss, reduction "mean") -- SURVEY.md 8(f)3.
 *   fwd: out2[0] = mean |pred - true|, out2[1] = mean (pred - true)^2 over n contiguous fp32 values, one launch,
 *        fixed-order fp64 sums (deterministic).
 *   bwd: dpred = dmae[0] * sign(pred - true) / n + dmse[0] * 2 (pred - true) / n ; dmae / dmse are DEVICE scalars
 *        (nullable = 0), so the seed of the model's backward is produced without a host round trip.
 * ------------------------------------------------------------------------------------- */
int cartnet_loss_l1_mse(const float* pred, const float* truth, int64_t n, float* out2, cartnet_stream_t stream);
int cartnet_loss_l1_mse_bwd(const float* pred, const float* truth, int64_t n, const float* dmae, const float* dmse,
                            float* dpred, cartnet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CARTNET_B200_H */
