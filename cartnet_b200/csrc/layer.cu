// Whole-layer orchestration: CartNet_layer forward / backward as one C-ABI call each. The host code below only
// sequences the kernels of the primitive entry points (same launches, same order as cartnet_b200/functional.py
// used to issue from Python), so that a training step costs ~10 FFI calls instead of ~140.
#include "common.cuh"

namespace cartnet {

// One launch: pack / cast / transpose all weights of a layer into the T-typed operand buffers.
template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ G1, const float* __restrict__ A1, const float* __restrict__ G2,
                                    const float* __restrict__ A2, const float* __restrict__ bg1, const float* __restrict__ ba1,
                                    int D, T* __restrict__ W1n, T* __restrict__ W1e, T* __restrict__ G2t, T* __restrict__ A2t,
                                    T* __restrict__ W1nT, T* __restrict__ W1eT, T* __restrict__ G2T, T* __restrict__ A2T,
                                    float* __restrict__ b1) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t DD = (int64_t)D * D;
    if (i < 2 * (int64_t)D) b1[i] = i < D ? bg1[i] : ba1[i - D];
    if (i < 4 * DD) {            // W1n[r, c], r in [0,4D): blocks G1_i, A1_i, G1_j, A1_j
        const int r = (int)(i / D), c = (int)(i % D), blk = r / D, rr = r % D;
        const float* src = (blk & 1) ? A1 : G1;
        const float v = src[(int64_t)rr * 3 * D + (blk >> 1) * D + c];
        store1<T>(&W1n[i], v);
        store1<T>(&W1nT[(int64_t)c * 4 * D + r], v);
    } else if (i < 6 * DD) {     // W1e[r, c], r in [0,2D): blocks G1_e, A1_e
        const int64_t k = i - 4 * DD;
        const int r = (int)(k / D), c = (int)(k % D), blk = r / D, rr = r % D;
        const float v = (blk ? A1 : G1)[(int64_t)rr * 3 * D + 2 * D + c];
        store1<T>(&W1e[k], v);
        store1<T>(&W1eT[(int64_t)c * 2 * D + r], v);
    } else if (i < 7 * DD) {
        const int64_t k = i - 6 * DD;
        const int r = (int)(k / D), c = (int)(k % D);
        const float v = G2[k];
        store1<T>(&G2t[k], v);
        store1<T>(&G2T[(int64_t)c * D + r], v);
    } else if (i < 8 * DD) {
        const int64_t k = i - 7 * DD;
        const int r = (int)(k / D), c = (int)(k % D);
        const float v = A2[k];
        store1<T>(&A2t[k], v);
        store1<T>(&A2T[(int64_t)c * D + r], v);
    }
}

__global__ void bias_grad_kernel(const float* __restrict__ sums1, const float* __restrict__ sums2, const float* __restrict__ bn1_w,
                                 const float* __restrict__ var1, float eps, int training, int D, float* dba2, float* dbg2,
                                 float* dbn1_w, float* dbn1_b, float* dbn2_w, float* dbn2_b) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    dba2[c] = sums1[2 * D + c];
    // d(bg2) = sum_e dg: identically zero under batch statistics, gamma * rstd * sum_e dghat under running statistics
    dbg2[c] = training ? 0.f : (bn1_w ? bn1_w[c] : 1.f) * (1.0f / sqrtf(var1[c] + eps)) * sums1[c];
    dbn1_b[c] = sums1[c];
    dbn1_w[c] = sums1[D + c];
    dbn2_b[c] = sums2[c];
    dbn2_w[c] = sums2[D + c];
}

static inline size_t tsize(int prec) { return prec == CARTNET_PREC_BF16 ? 2 : 4; }
static inline size_t zsize(int prec) { return prec == CARTNET_PREC_BF16X3 ? 2 : tsize(prec); }   // pre-activations: fp16 in the pair mode
static inline const void* toff(const void* p, int prec, int64_t elems) { return (const char*)p + elems * (int64_t)tsize(prec); }
static inline void* toff(void* p, int prec, int64_t elems) { return (char*)p + elems * (int64_t)tsize(prec); }

static cartnet_gemm_t gemm_desc(int prec, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb) {
    cartnet_gemm_t d = {};
    d.prec = prec; d.M = M; d.N = N; d.K = K; d.A = A; d.lda = lda; d.B = B; d.ldb = ldb;
    return d;
}

#define CN_TRY(expr)            \
    do {                        \
        int _rc = (expr);       \
        if (_rc != 0) return _rc; \
    } while (0)

}  // namespace cartnet

using namespace cartnet;

extern "C" {

int64_t cartnet_layer_splitk_bytes(int32_t prec, int32_t D, int32_t num_nodes, int64_t num_edges) {
    int64_t a = cartnet_gemm_tn_workspace(prec, D, D, num_edges);
    int64_t b = cartnet_gemm_tn_workspace(prec, D, D, num_nodes);
    int64_t c = cartnet_gemm_tn_workspace(prec, 2 * D, D, num_edges);
    int64_t d = cartnet_gemm_tn_workspace(prec, 4 * D, D, num_nodes);
    int64_t m = a > b ? a : b;
    if (c > m) m = c;
    return (m > d ? m : d) + 256;
}

int cartnet_layer_pack_weights(const cartnet_layer_t* L, cartnet_stream_t stream) {
    CN_CHECK_ARG(L && L->G1 && L->A1 && L->G2 && L->A2 && L->bg1 && L->ba1 && L->b1, "layer_pack_weights: null parameter");
    CN_CHECK_ARG(L->W1n_t && L->W1e_t && L->G2_t && L->A2_t && L->W1nT_t && L->W1eT_t && L->G2T_t && L->A2T_t, "layer_pack_weights: null output");
    const int64_t total = 8 * (int64_t)L->D * L->D;
    CN_DISPATCH_PREC(L->prec, {
        pack_weights_kernel<T><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
            L->G1, L->A1, L->G2, L->A2, L->bg1, L->ba1, L->D, (T*)L->W1n_t, (T*)L->W1e_t, (T*)L->G2_t, (T*)L->A2_t,
            (T*)L->W1nT_t, (T*)L->W1eT_t, (T*)L->G2T_t, (T*)L->A2T_t, L->b1);
    });
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_layer_fwd(const cartnet_layer_t* L, cartnet_stream_t st) {
    CN_CHECK_ARG(L, "layer_fwd: null");
    const int D = L->D, N = L->num_nodes, prec = L->prec;
    const int64_t E = L->num_edges;
    CN_CHECK_ARG(E < (int64_t)1 << 31, "layer_fwd: more than 2^31 edges per call");
    CN_CHECK_ARG(!L->training || E > 1, "layer_fwd: training-mode BatchNorm needs more than 1 edge");
    // per-node projections P = x W1n^T : [:, 0:2D] dst-role (gate|aggr), [:, 2D:4D] src-role
    {
        cartnet_gemm_t d = gemm_desc(prec, N, 4 * D, D, L->x_t, D, L->W1n_t, D);
        if (prec == CARTNET_PREC_BF16X3) { d.out_f32 = (float*)L->P; d.ldo = 4 * D; }      // gathered operands stay plain fp32 (see cartnet_gemm_t)
        else { d.out_t = L->P; d.ldt = 4 * D; }
        CN_TRY(cartnet_gemm(&d, st));
    }
    // per-edge first Linear with gathered projections + SiLU                         (cartnet.py:237,256)
    if (E > 0) {
        cartnet_gemm_t d = gemm_desc(prec, (int)E, 2 * D, D, L->e_t, D, L->W1e_t, D);
        d.bias = L->b1;
        d.gather0 = L->P; d.gidx0 = L->dst32;
        d.gather1 = toff((const void*)L->P, prec, 2 * D); d.gidx1 = L->src32;
        d.ldg = 4 * D;
        d.z_out = L->Z; d.ldz = 2 * D;
        d.act = CARTNET_ACT_SILU;
        d.out_t = L->H; d.ldt = 2 * D;
        CN_TRY(cartnet_gemm(&d, st));
        // second Linears                                                             (cartnet.py:190,195)
        // g is stored centred (g - center, T): BatchNorm removes the shift, the bits go to the part it keeps
        CN_TRY(cartnet_gate_center(L->H, 2 * D, E, D, L->G2, L->bg2, L->bn1_rm, L->training, prec, L->bias_c, L->center, L->hsum,
                                   L->partial, st));
        cartnet_gemm_t dg = gemm_desc(prec, (int)E, D, D, L->H, 2 * D, L->G2_t, D);
        dg.bias = L->bias_c; dg.out_t = L->g_t; dg.ldt = D;
        if (L->training) {
            // edge BatchNorm statistics (global barrier over E rows, cartnet.py:238) ride in this GEMM's epilogue
            CN_TRY(cartnet_gemm_colstats(&dg, L->center, L->mean1, L->var1, L->bn1_rm, L->bn1_rv, L->momentum1, L->partial, st));
        } else {
            CN_TRY(cartnet_gemm(&dg, st));
        }
        cartnet_gemm_t ds = gemm_desc(prec, (int)E, D, D, toff((const void*)L->H, prec, D), 2 * D, L->A2_t, D);
        ds.bias = L->ba2; ds.out_t = L->s_t; ds.ldt = D;
        CN_TRY(cartnet_gemm(&ds, st));
    }
    const float *mean1 = nullptr, *var1 = L->bn1_rv, *mean2 = L->bn2_rm, *var2 = L->bn2_rv;   // eval: g_t is centred on bn1_rm
    if (L->training) { mean1 = L->mean1; var1 = L->var1; }
    CN_TRY(cartnet_edge_gate_aggregate(L->g_t, L->s_t, L->e, L->dist, L->row_ptr, N, E, D, mean1, var1, L->bn1_w, L->bn1_b, L->eps,
                                       L->radius, L->use_envelope, L->e_out, L->e_out_t, L->gn_t, prec, L->m, st));
    if (L->training) {
        CN_TRY(cartnet_colstats(L->m, 0, prec, N, D, D, nullptr, L->mean2, L->var2, L->bn2_rm, L->bn2_rv, L->momentum2, L->partial, st));
        mean2 = L->mean2; var2 = L->var2;
    }
    return cartnet_node_update(L->m, L->x, N, D, mean2, var2, L->bn2_w, L->bn2_b, L->eps, L->x_out, L->x_out_t, prec, st);   // cartnet.py:269,223
}

int cartnet_layer_bwd(const cartnet_layer_t* L, cartnet_stream_t st) {
    CN_CHECK_ARG(L && L->dx_out, "layer_bwd: null gradient input");      // de_out may be null (= zero)
    const int D = L->D, N = L->num_nodes, prec = L->prec;
    const int64_t E = L->num_edges;
    const float* var1 = L->training ? L->var1 : L->bn1_rv;
    const float *mean2 = L->training ? L->mean2 : L->bn2_rm, *var2 = L->training ? L->var2 : L->bn2_rv;
    // node side: x' = silu(BN2(m)) + x
    CN_TRY(cartnet_node_update_bwd_reduce(L->dx_out, L->m, N, D, mean2, var2, L->bn2_w, L->bn2_b, L->eps, L->sums2, L->partial, st));
    CN_TRY(cartnet_node_update_bwd_apply(L->dx_out, L->m, N, D, mean2, var2, L->bn2_w, L->bn2_b, L->eps, L->sums2, L->training, L->dm, st));
    // edge side: sig = env * sigmoid(BN1(g)); e' = e + sig; m = segsum(sig * s)
    // gn_t null: the forward pass kept the centred pre-activation g_t instead of writing a normalised copy; the backward
    // kernels normalise on the fly (training: batch mean of the stored values; eval: g_t is centred on the running mean)
    const void* gsrc = L->gn_t ? L->gn_t : L->g_t;
    const float* gmean = (!L->gn_t && L->training) ? L->mean1 : nullptr;
    const float* gvar = L->gn_t ? nullptr : var1;
    CN_TRY(cartnet_edge_gate_bwd_reduce(gsrc, L->s_t, L->dist, L->dst32, L->de_out, L->dm, E, D, L->bn1_w, L->bn1_b, L->radius,
                                        L->use_envelope, L->ds_t, L->dghat_t, prec, L->sums1, L->partial, gmean, gvar, L->eps, st));
    CN_TRY(cartnet_edge_gate_bwd_apply(gsrc, L->dghat_t, E, D, var1, L->bn1_w, L->eps, L->sums1, L->training, L->dg_t, prec, gmean,
                                       L->gn_t ? 0 : 1, st));
    bias_grad_kernel<<<ceil_div(D, 128), 128, 0, (cudaStream_t)st>>>(L->sums1, L->sums2, L->bn1_w, var1, L->eps, L->training, D, L->dba2,
                                                                      L->dbg2, L->dbn1_w, L->dbn1_b, L->dbn2_w, L->dbn2_b);
    CN_LAUNCH_CHECK();
    // second Linears: dgrad (* SiLU') into dZ = [dZ_gate | dZ_aggr], wgrad straight into dG2 / dA2
    {
        cartnet_gemm_t d = gemm_desc(prec, (int)E, D, D, L->dg_t, D, L->G2T_t, D);
        d.act = CARTNET_ACT_MUL_DSILU; d.z_in = L->Z; d.ldzin = 2 * D; d.out_t = L->dZ; d.ldt = 2 * D;
        CN_TRY(cartnet_gemm(&d, st));
        cartnet_gemm_t a = gemm_desc(prec, (int)E, D, D, L->ds_t, D, L->A2T_t, D);
        a.act = CARTNET_ACT_MUL_DSILU; a.z_in = (const char*)L->Z + (int64_t)D * (int64_t)zsize(prec); a.ldzin = 2 * D;
        a.out_t = toff(L->dZ, prec, D); a.ldt = 2 * D;
        CN_TRY(cartnet_gemm(&a, st));
    }
    CN_TRY(cartnet_gemm_tn(prec, D, D, E, L->dg_t, D, L->H, 2 * D, L->dG2, D, L->splitk, L->splitk_bytes, st));
    CN_TRY(cartnet_gemm_tn(prec, D, D, E, L->ds_t, D, toff((const void*)L->H, prec, D), 2 * D, L->dA2, D, L->splitk, L->splitk_bytes, st));
    // first Linear, edge part: de = dZ W1e + de_out (residual e' = e + sig)
    if (prec == CARTNET_PREC_BF16X3 && E < CARTNET_NT_PAIR_MIN_ROWS) {
        // pair operands on a single CTA per tile: the resident weight slice for K = 2D would be only 64 rows (128 KB /
        // (2D x 4 B)), i.e. four passes over dZ per launch with two 32 KB stages in flight (measured 0.80 ms at ADP-64). The
        // two K halves -- the gate and the aggregate branch -- are contracted by two launches with 128-row slices instead,
        // the second accumulating onto the first through the residual input (same element, same thread): 2 x 0.37 ms.
        // Edge counts that run as CTA pairs (cta_group::2, 128 columns per pair at K = 2D) take ONE launch: 0.60 ms.
        for (int h = 0; h < 2; ++h) {
            cartnet_gemm_t d = gemm_desc(prec, (int)E, D, D, toff((const void*)L->dZ, prec, (int64_t)h * D), 2 * D,
                                         toff((const void*)L->W1eT_t, prec, (int64_t)h * D), 2 * D);
            d.resid = h == 0 ? L->de_out : L->de_in; d.ldr = D; d.out_f32 = L->de_in; d.ldo = D;
            CN_TRY(cartnet_gemm(&d, st));
        }
    } else {
        cartnet_gemm_t d = gemm_desc(prec, (int)E, D, 2 * D, L->dZ, 2 * D, L->W1eT_t, 2 * D);
        d.resid = L->de_out; d.ldr = D; d.out_f32 = L->de_in; d.ldo = D;      // null resid: plain store
        CN_TRY(cartnet_gemm(&d, st));
    }
    // d(W_e) = dZ^T e, written block-wise into the reference layout dG1[:, 2D:3D], dA1[:, 2D:3D]
    {
        float* blocks[2] = {L->dG1 + 2 * D, L->dA1 + 2 * D};
        CN_TRY(cartnet_gemm_tn_blocks(prec, 2 * D, D, E, L->dZ, 2 * D, L->e_t, D, blocks, 2, 3 * D, L->splitk, L->splitk_bytes, st));
    }
    // first Linear, node part: transpose of the two lifts = segmented sums by dst and by src
    CN_TRY(cartnet_segment_sum_pair(L->dZ, 2 * D, L->row_ptr, L->col_ptr, L->perm_src, N, 2 * D, L->dP, 4 * D, 1, prec, st));
    // d(b1) = sum_e dZ = column sums of d(P_dst) over N rows; [0:D] -> dbg1, [D:2D] -> dba1
    CN_TRY(cartnet_colsum(L->dP, 1, prec, N, D, 4 * D, L->dbg1, L->partial, st));
    CN_TRY(cartnet_colsum(toff((const void*)L->dP, prec, D), 1, prec, N, D, 4 * D, L->dba1, L->partial, st));
    {
        // bf16x3: the resident weight slice (32 x K x 4 B) must fit 128 KB: K = 4D > 1024 is contracted in two halves,
        // the second one accumulating onto the first through the residual input (same element, same thread)
        const int ks = (prec == CARTNET_PREC_BF16X3 && D > 256) ? 2 : 1;
        const int Kh = 4 * D / ks;
        for (int h = 0; h < ks; ++h) {
            cartnet_gemm_t d = gemm_desc(prec, N, D, Kh, toff((const void*)L->dP, prec, (int64_t)h * Kh), 4 * D,
                                         toff((const void*)L->W1nT_t, prec, (int64_t)h * Kh), 4 * D);
            d.resid = h == 0 ? L->dx_out : L->dx_in; d.ldr = D; d.out_f32 = L->dx_in; d.ldo = D;
            CN_TRY(cartnet_gemm(&d, st));
        }
    }
    // d(W_i), d(W_j) = dP^T x, four [D,D] blocks: G1_i, A1_i, G1_j, A1_j
    {
        float* blocks[4] = {L->dG1, L->dA1, L->dG1 + D, L->dA1 + D};
        CN_TRY(cartnet_gemm_tn_blocks(prec, 4 * D, D, N, L->dP, 4 * D, L->x_t, D, blocks, 4, 3 * D, L->splitk, L->splitk_bytes, st));
    }
    return 0;
}

}  // extern "C"
