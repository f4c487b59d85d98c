"""autograd Functions that compose the C-ABI kernels into the two hot-path operators:

  * edge_encoder -- the edge branch of Encoder.forward   (/root/reference/models/cartnet.py:133-138,156-159)
  * cartnet_layer -- CartNet_layer.forward (message / aggregate / update)   (cartnet.py:204-274)

Forward and backward are both hand-written (no autograd replay of eager ops). The first Linear of
the gate / aggregate MLPs is split W1 = [W_i | W_j | W_e]: W_i x_i and W_j x_j are
per-NODE projections (one small GEMM), gathered per edge inside the epilogue of the per-EDGE GEMM
e @ W_e^T, which halves the edge FLOPs and removes the [E,768] concatenations. Mathematically
identical to the reference; only the fp32 summation order differs (inside the 1e-5 budget).

`prec` selects the GEMM operand type T: fp32 SIMT (1e-5 parity) or bf16 tcgen05 (2e-3 parity).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from ._lib import NT_PAIR_MIN_ROWS
from .ops import ACT_MUL_DSILU, ACT_SILU, PREC_BF16, PREC_BF16X3, PREC_FP32, PREC_TF32, f32_storage, needs_shadow, t_dtype, z_dtype


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _to_t(w: torch.Tensor, prec: int) -> torch.Tensor:
    """weights are tiny (<= 1 MB): a torch cast is plumbing, not hot path"""
    w = w.detach()
    if prec in (PREC_TF32, PREC_BF16X3):
        return ops.cast(w.contiguous(), prec)          # fp32 words rounded to tf32 (the MMA would truncate otherwise) / hi|lo bf16 pairs
    return w.contiguous() if f32_storage(prec) else w.to(torch.bfloat16).contiguous()


class _EdgeEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cart_dist, cart_dir, means, betas, Wa, ba, Wb, bb, upper, invariant, prec, holder):
        dim_edge = int(Wa.shape[1])
        D2, D = int(Wa.shape[0]), int(Wb.shape[0])
        KF = _round_up(dim_edge, {PREC_FP32: 4, PREC_BF16: 64, PREC_TF32: 32, PREC_BF16X3: 64}[prec])   # tcgen05: whole 128-byte K blocks
        T = t_dtype(prec)
        dev = cart_dist.device
        E = int(cart_dist.shape[0])
        feat = ops.edge_features(cart_dist, cart_dir, means, betas, upper, invariant, KF, prec)
        Wa_t = _to_t(F.pad(Wa.detach(), (0, KF - dim_edge)), prec)
        Wb_t = _to_t(Wb, prec)
        need_bwd = any(ctx.needs_input_grad)       # inference (torch.no_grad): the pre-activations are never read -- not stored
        Z1 = torch.empty(E, D2, dtype=z_dtype(prec), device=dev) if need_bwd else None
        H1 = torch.empty(E, D2, dtype=T, device=dev)
        ops.gemm(prec, feat, Wa_t, bias=ba.detach(), z_out=Z1, act=ACT_SILU, out_t=H1)
        Z2 = torch.empty(E, D, dtype=z_dtype(prec), device=dev) if need_bwd else None
        e0 = torch.empty(E, D, dtype=torch.float32, device=dev)
        e0_t = torch.empty(E, D, dtype=T, device=dev) if needs_shadow(prec) else None
        ops.gemm(prec, H1, Wb_t, bias=bb.detach(), z_out=Z2, act=ACT_SILU, out_f32=e0, out_t=e0_t)
        holder["e_t"] = e0_t if needs_shadow(prec) else e0
        if need_bwd:
            ctx.save_for_backward(feat, Z1, H1, Z2, Wb)
        ctx.prec, ctx.dim_edge = prec, dim_edge
        return e0

    @staticmethod
    def backward(ctx, de0):
        feat, Z1, H1, Z2, Wb = ctx.saved_tensors
        prec = ctx.prec
        T = t_dtype(prec)
        dz2, dbb = ops.dsilu_mul(de0.contiguous(), Z2, prec, want_colsum=True)     # [E, D]; sum_e dz2 from the same pass
        dWb = ops.gemm_tn(prec, dz2, H1)                                # [D, 2D]
        dz1 = torch.empty(Z1.shape, dtype=T, device=Z1.device)
        ops.gemm(prec, dz2, _to_t(Wb.t(), prec), act=ACT_MUL_DSILU, z_in=Z1, out_t=dz1)
        dWa_full = ops.gemm_tn(prec, dz1, feat)                         # [2D, KF]
        dWa = dWa_full[:, :ctx.dim_edge]
        # feat carries a column of ones right after the features when the K padding leaves room (ops.edge_features):
        # sum_e dz1 is that column of the weight gradient, no reduction pass over [E, 2D]
        dba = dWa_full[:, ctx.dim_edge].contiguous() if feat.shape[1] > ctx.dim_edge else ops.colsum(dz1, prec)
        return None, None, None, None, dWa, dba, dWb, dbb, None, None, None, None


def edge_encoder(cart_dist, cart_dir, means, betas, Wa, ba, Wb, bb, upper, invariant, prec):
    """Returns (edge_attr fp32 [E,D], T-typed operand copy for the first layer)."""
    holder = {}
    e0 = _EdgeEncoderFn.apply(cart_dist, cart_dir, means, betas, Wa, ba, Wb, bb, float(upper), bool(invariant),
                              prec, holder)
    return e0, holder["e_t"]


class _LinearSiluFn(torch.autograd.Function):
    """y = SiLU(x W^T + b) for the node branch of the encoder (/root/reference/models/cartnet.py:125-127,154):
    [N, 2D] x [D, 2D]^T on the library's GEMMs with the bias + SiLU epilogue and a hand-written backward, instead of
    three eager cuBLAS SIMT sgemms. Node-side and tiny, so the bf16 mode runs it on the tf32 tensor-core path: no
    additional bf16 rounding enters the parity budget."""

    @staticmethod
    def forward(ctx, x, W, b, prec):
        gp = PREC_TF32 if prec == PREC_BF16 else prec
        M, N = int(x.shape[0]), int(W.shape[0])
        x_t = ops.cast(x.detach().contiguous(), gp)
        Z = torch.empty(M, N, dtype=z_dtype(gp), device=x.device)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        ops.gemm(gp, x_t, _to_t(W, gp), bias=b.detach(), z_out=Z, act=ACT_SILU, out_f32=y)
        ctx.save_for_backward(x_t, Z, W)
        ctx.gp = gp
        return y

    @staticmethod
    def backward(ctx, dy):
        x_t, Z, W = ctx.saved_tensors
        gp = ctx.gp
        dz, db = ops.dsilu_mul(dy.contiguous(), Z, gp, want_colsum=True)      # [M, N]; the bias gradient from the same pass
        if dz.shape[1] % 128 != 0 and x_t.shape[1] % 128 == 0:
            dW = ops.gemm_tn(gp, x_t, dz).t()                           # [K, N]^T: the tcgen05 TN kernel owns 128-row accumulators
        else:
            dW = ops.gemm_tn(gp, dz, x_t)                               # [N, K]
        dx = torch.empty(x_t.shape[0], x_t.shape[1], dtype=torch.float32, device=dy.device)
        ops.gemm(gp, dz, _to_t(W.t(), gp), out_f32=dx)
        return dx, dW, db, None


def linear_silu(x, W, b, prec):
    return _LinearSiluFn.apply(x, W, b, prec)


class _EmbeddingRowsFn(torch.autograd.Function):
    """y = W[idx] -- nn.Embedding's forward (/root/reference/models/cartnet.py:113,146) with a DETERMINISTIC backward.
    torch's CUDA embedding backward accumulates partial segments with float atomics once there are more than 3072 indices
    (the ADP-64 batch has 12.5 k atoms): the one gradient of a training step that was not bit-reproducible. Here the rows
    of dy are grouped by index with a stable sort and summed in a fixed order in two levels -- pieces of at most 64
    consecutive rows of one index (many short segments: parallel), then the pieces of each index -- both by the library's
    CSR segment sum. No host synchronisation: the piece table has a fixed upper size and is padded with empty pieces."""

    PIECE = 64

    @staticmethod
    def forward(ctx, W, idx):
        ctx.save_for_backward(idx)
        ctx.rows = int(W.shape[0])
        return W.index_select(0, idx)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        T, N, C = ctx.rows, int(idx.numel()), int(dy.shape[1])
        dev = dy.device
        out = torch.zeros(T, C, dtype=torch.float32, device=dev)
        if N == 0:
            return out, None
        R = _EmbeddingRowsFn.PIECE
        order = torch.sort(idx, stable=True)
        sidx, perm = order.values, order.indices.to(torch.int32)
        counts = torch.zeros(T, dtype=torch.int64, device=dev).scatter_add_(0, idx, torch.ones_like(idx))   # (bincount would sync)
        tstart = torch.cumsum(counts, 0) - counts                       # first sorted position of every index value
        pos = torch.arange(N, device=dev) - tstart[sidx]                # position inside its group
        first = (pos % R) == 0                                          # row opens a piece
        piece = torch.cumsum(first.to(torch.int64), 0) - 1              # piece id of every sorted row
        P = (N + R - 1) // R + T                                        # upper bound on the number of pieces
        ptr1 = torch.full((P + 2,), N, dtype=torch.int32, device=dev)   # unused pieces are empty ([N, N)); slot P+1 is a dump
        rows = torch.arange(N, device=dev, dtype=torch.int32)
        ptr1.scatter_(0, torch.where(first, piece, torch.full_like(piece, P + 1)), rows)   # (mask indexing would sync)
        ptr1 = ptr1[:P + 1]
        part = torch.empty(P, C, dtype=torch.float32, device=dev)
        ops.segment_sum(dy.contiguous(), ptr1, perm, P, part, PREC_FP32)
        # pieces of one index value are consecutive: its first piece is the piece of its first sorted row
        piece_ext = torch.cat([piece, piece[-1:] + 1])
        ptr2 = torch.cat([piece_ext[tstart.clamp(max=N)], piece_ext[-1:]]).to(torch.int32)
        ops.segment_sum(part, ptr2, None, T, out, PREC_FP32)
        return out, None


def embedding_rows(W, idx):
    return _EmbeddingRowsFn.apply(W, idx)


class _CholeskyTailFn(torch.autograd.Function):
    """(h [n, Dh], W1 [6, Dh], b1 [6]) -> U [n,3,3] = L^T L with L upper triangular, softplus diagonal
    (/root/reference/models/cartnet.py:293-303): one launch forward, two backward."""

    @staticmethod
    def forward(ctx, h, W1, b1):
        h = h.detach().contiguous()
        U, p6 = ops.cholesky_head_fwd(h, W1.detach().contiguous(), b1.detach().contiguous())
        ctx.save_for_backward(h, p6, W1)
        return U

    @staticmethod
    def backward(ctx, dU):
        h, p6, W1 = ctx.saved_tensors
        dh, dW1, db1 = ops.cholesky_head_bwd(dU.contiguous(), h, p6, W1.detach().contiguous())
        return dh, dW1, db1


def cholesky_tail(h, W1, b1):
    return _CholeskyTailFn.apply(h, W1, b1)


class _LossPairFn(torch.autograd.Function):
    """(pred, true) -> (MAE, MSE), the loss pair of /root/reference/train/metrics.py:15-28, one launch forward and one
    backward instead of nn.L1Loss + nn.MSELoss (two reductions forward, ~6 eager launches backward)."""

    @staticmethod
    def forward(ctx, pred, true):
        p = pred.detach().contiguous().to(torch.float32)
        t = true.detach().contiguous().to(torch.float32)
        if p.shape != t.shape:
            raise ValueError("compute_loss: pred %s and true %s differ in shape" % (tuple(p.shape), tuple(t.shape)))
        out = ops.loss_l1_mse(p, t)
        ctx.save_for_backward(p, t)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, dmae, dmse):
        p, t = ctx.saved_tensors
        dm = None if dmae is None else dmae.detach().reshape(1).to(torch.float32).contiguous()
        dq = None if dmse is None else dmse.detach().reshape(1).to(torch.float32).contiguous()
        return ops.loss_l1_mse_bwd(p, t, dm, dq), None


def compute_loss(pred, true):
    """Drop-in for train/metrics.py::compute_loss: returns (MAE, MSE) as 0-dim tensors with autograd through `pred`."""
    return _LossPairFn.apply(pred, true)


class _LayerFn(torch.autograd.Function):
    """Inputs: x [N,D], e [E,D] (fp32) and the packed weights
         W1n [4D,D] = [G1_i ; A1_i ; G1_j ; A1_j],  W1e [2D,D] = [G1_e ; A1_e],  b1 [2D] = [bg1 ; ba1],
         G2, A2 [D,D], bg2, ba2 [D], BatchNorm affine (w1,b1n over edges; w2,b2n over nodes).
       cfg: dict(plan, dist, x_t, e_t, prec, training, radius, use_envelope, momentum, bn buffers, holder)."""

    @staticmethod
    def forward(ctx, x, e, W1n, W1e, b1, G2, A2, bg2, ba2, w1, b1n, w2, b2n, cfg):
        prec, plan, dist = cfg["prec"], cfg["plan"], cfg["dist"]
        training = cfg["training"]
        T = t_dtype(prec)
        dev = x.device
        N, D = int(x.shape[0]), int(x.shape[1])
        E = int(e.shape[0])
        x = x.detach().contiguous()
        e = e.detach().contiguous()
        x_t = cfg.get("x_t")
        e_t = cfg.get("e_t")
        if x_t is None:
            x_t = ops.cast(x, prec)
        if e_t is None:
            e_t = ops.cast(e, prec)
        W1n_t, W1e_t, G2_t, A2_t = (_to_t(w, prec) for w in (W1n, W1e, G2, A2))

        # per-node projections P = x W1n^T : [:, 0:2D] dst-role (gate|aggr), [:, 2D:4D] src-role
        P = torch.empty(N, 4 * D, dtype=T, device=dev)
        if prec == PREC_BF16X3:
            ops.gemm(prec, x_t, W1n_t, out_f32=P)       # gathered operands are plain fp32 in this mode (only added, never contracted)
        else:
            ops.gemm(prec, x_t, W1n_t, out_t=P)
        # per-edge first Linear with gathered projections, SiLU                       (cartnet.py:237,256)
        need_bwd = any(ctx.needs_input_grad)
        Z = torch.empty(E, 2 * D, dtype=z_dtype(prec), device=dev) if need_bwd else None     # inference: pre-activations are not stored
        H = torch.empty(E, 2 * D, dtype=T, device=dev)
        ops.gemm(prec, e_t, W1e_t, bias=b1.detach(), gather0=P[:, :2 * D], gidx0=plan.dst32,
                 gather1=P[:, 2 * D:], gidx1=plan.src32, z_out=Z, act=ACT_SILU, out_t=H)
        # second Linears                                                              (cartnet.py:190,195)
        # g is stored centred (g - center) in T: BatchNorm removes the shift, the bits go to the part it keeps
        g = torch.empty(E, D, dtype=T, device=dev)                  # scratch after this pass
        s = torch.empty(E, D, dtype=T, device=dev)
        mean1, var1 = None, cfg["rv1"]                              # eval: g is centred on the running mean
        if E > 0:
            bias_c, center = ops.gate_center(H[:, :D], G2.detach(), bg2.detach(), cfg["rm1"], training, prec)
            if training:
                # edge BatchNorm statistics (global barrier over E rows, cartnet.py:238) come with the GEMM
                mean1, var1 = ops.gemm_colstats(prec, H[:, :D], G2_t, bias_c, g, cfg["rm1"], cfg["rv1"], cfg["momentum1"],
                                                shift=center)
            else:
                ops.gemm(prec, H[:, :D], G2_t, bias=bias_c, out_t=g)
            ops.gemm(prec, H[:, D:], A2_t, bias=ba2.detach(), out_t=s)
        # no normalised copy of g is written: the backward kernels normalise the stored (centred) g on the fly
        e_out, e_out_t, m, _ = ops.edge_gate_aggregate(g, s, e, dist, plan.row_ptr, N, mean1, var1, w1.detach(),
                                                       b1n.detach(), cfg["radius"], cfg["use_envelope"], prec, True, want_gn=False)
        if training:
            mean2, var2 = ops.colstats(m, cfg["rm2"], cfg["rv2"], cfg["momentum2"])
        else:
            mean2, var2 = cfg["rm2"], cfg["rv2"]
        x_out, x_out_t = ops.node_update(m, x, mean2, var2, w2.detach(), b2n.detach(), prec, True)   # cartnet.py:269,223
        cfg["holder"]["x_t"], cfg["holder"]["e_t"] = x_out_t, e_out_t

        if need_bwd:
            ctx.save_for_backward(x_t, e_t, Z, H, g, s, m, var1, mean2, var2, W1n, W1e, G2, A2, w1, b1n, w2, b2n)
        ctx.cfg = dict(prec=prec, plan=plan, dist=dist, training=training, radius=cfg["radius"],
                       use_envelope=cfg["use_envelope"], mean1=mean1)
        return x_out, e_out

    @staticmethod
    def backward(ctx, dx_out, de_out):
        (x_t, e_t, Z, H, gn, s, m, var1, mean2, var2, W1n, W1e, G2, A2, w1, b1n, w2, b2n) = ctx.saved_tensors
        c = ctx.cfg
        prec, plan, training = c["prec"], c["plan"], c["training"]
        T = t_dtype(prec)
        dev = gn.device
        N, D = int(m.shape[0]), int(m.shape[1])
        E = int(gn.shape[0])
        if dx_out is None:
            dx_out = torch.zeros(N, D, dtype=torch.float32, device=dev)
        if de_out is None:
            de_out = torch.zeros(E, D, dtype=torch.float32, device=dev)
        dx_out = dx_out.contiguous()
        de_out = de_out.contiguous()

        # node side: x' = silu(BN2(m)) + x
        dm, sums2 = ops.node_update_bwd(dx_out, m, mean2, var2, w2, b2n, training)
        # edge side: sig = env * sigmoid(BN1(g)); e' = e + sig; m = segsum(sig * s)
        ds_t, dg_t, sums1 = ops.edge_gate_bwd(gn, s, c["dist"], plan.dst32, de_out, dm, var1, w1, b1n,
                                              c["radius"], c["use_envelope"], training, prec, g_mean=c["mean1"], input_is_g=True)
        # second Linears: dgrad (+ SiLU') and wgrad
        dZ = torch.empty(E, 2 * D, dtype=T, device=dev)
        ops.gemm(prec, dg_t, _to_t(G2.t(), prec), act=ACT_MUL_DSILU, z_in=Z[:, :D], out_t=dZ[:, :D])
        ops.gemm(prec, ds_t, _to_t(A2.t(), prec), act=ACT_MUL_DSILU, z_in=Z[:, D:], out_t=dZ[:, D:])
        dG2 = ops.gemm_tn(prec, dg_t, H[:, :D])
        dA2 = ops.gemm_tn(prec, ds_t, H[:, D:])
        # bias gradients without extra passes over [E, D]: d(ba2) = sum_e ds is a third output of the reduction
        # above; d(bg2) = sum_e dg is identically zero under batch statistics (BatchNorm removes any shift of g;
        # the reference's value is pure rounding noise) and gamma*rstd*sum_e dghat under running statistics
        dba2 = sums1[2 * D:].clone()
        if training:
            dbg2 = torch.zeros(D, dtype=torch.float32, device=dev)
        else:
            dbg2 = w1.detach() * torch.rsqrt(var1 + ops.EPS_BN) * sums1[:D]
        # first Linear, edge part: de = dZ W1e + de_out (residual e' = e + sig)
        de_in = torch.empty(E, D, dtype=torch.float32, device=dev)
        W1eT = _to_t(W1e.t(), prec)
        if prec == PREC_BF16X3 and E < NT_PAIR_MIN_ROWS:      # as in csrc/layer.cu: two K-half launches unless the GEMM runs as CTA pairs
            for h in range(2):
                ops.gemm(prec, dZ[:, h * D:(h + 1) * D], W1eT[:, h * D:(h + 1) * D], resid=de_out if h == 0 else de_in, out_f32=de_in)
        else:
            ops.gemm(prec, dZ, W1eT, resid=de_out, out_f32=de_in)
        dW1e = ops.gemm_tn(prec, dZ, e_t)
        # first Linear, node part: transpose of the two lifts = segmented sums by dst and by src
        dP = torch.empty(N, 4 * D, dtype=T, device=dev)
        ops.segment_sum_pair(dZ, plan.row_ptr, plan.col_ptr, plan.perm_src, N, dP, prec)
        # sum_e dZ = sum_n (sum_{e -> n} dZ): N rows instead of E (two D-wide reductions, as in csrc/layer.cu)
        db1 = torch.cat([ops.colsum(dP[:, :D], prec), ops.colsum(dP[:, D:2 * D], prec)])
        dx_in = torch.empty(N, D, dtype=torch.float32, device=dev)
        W1nT = _to_t(W1n.t(), prec)
        ks = 2 if (prec == PREC_BF16X3 and D > 256) else 1      # as in csrc/layer.cu: K = 4D in two halves
        Kh = 4 * D // ks
        for h in range(ks):
            ops.gemm(prec, dP[:, h * Kh:(h + 1) * Kh], W1nT[:, h * Kh:(h + 1) * Kh], resid=dx_out if h == 0 else dx_in, out_f32=dx_in)
        dW1n = ops.gemm_tn(prec, dP, x_t)
        dw1, db1n = sums1[D:2 * D].clone(), sums1[:D].clone()
        dw2, db2n = sums2[D:].clone(), sums2[:D].clone()
        return dx_in, de_in, dW1n, dW1e, db1, dG2, dA2, dbg2, dba2, dw1, db1n, dw2, db2n, None


def cartnet_layer(x, e, packed, cfg):
    """Python composition of the primitive entry points (the executable specification of cartnet_layer_fwd/bwd;
    used by the CPU host-logic tests with emulated primitives and by the GPU test that checks the native
    orchestration is bit-identical to it). packed = (W1n, W1e, b1, G2, A2, bg2, ba2, w1, b1n, w2, b2n)."""
    return _LayerFn.apply(x, e, *packed, cfg)


# ----------------------------------------------------------------------------------------------------------------
# native orchestration: one C-ABI call per layer forward / backward (cartnet_b200/csrc/layer.cu)
# ----------------------------------------------------------------------------------------------------------------
USE_NATIVE_LAYER = True


def _carve(buf, sizes_shapes):
    out, off = [], 0
    for shape in sizes_shapes:
        n = 1
        for d in shape:
            n *= d
        out.append(buf[off:off + n].view(*shape))
        off += n
    return out


class _NativeLayerFn(torch.autograd.Function):
    """inputs: x, e, then the module's own parameters in the reference layout
       (G1, A1, bg1, ba1, G2, A2, bg2, ba2, bn1_w, bn1_b, bn2_w, bn2_b)."""

    @staticmethod
    def forward(ctx, x, e, G1, A1, bg1, ba1, G2, A2, bg2, ba2, w1, b1n, w2, b2n, cfg):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        prec, plan, dist, training = cfg["prec"], cfg["plan"], cfg["dist"], cfg["training"]
        T = t_dtype(prec)
        dev = x.device
        if not x.is_cuda:
            raise RuntimeError("cartnet_b200: CartNet_layer needs CUDA tensors (no CPU fallback exists)")
        N, D = int(x.shape[0]), int(x.shape[1])
        E = int(e.shape[0])
        # Backward reads x, e and the parameters through raw pointers (ctx.L): saving the autograd-visible tensors as well
        # costs nothing (same storage) and keeps autograd's version check -- an in-place update of any of them between
        # forward and backward raises the standard error instead of silently producing gradients of the new values
        ctx.save_for_backward(x, e, G1, A1, bg1, ba1, G2, A2, bg2, ba2, w1, b1n, w2, b2n)
        x = x.detach().contiguous()
        e = e.detach().contiguous()
        x_t, e_t = cfg.get("x_t"), cfg.get("e_t")
        if x_t is None:
            x_t = ops.cast(x, prec)
        if e_t is None:
            e_t = ops.cast(e, prec)
        shadow = needs_shadow(prec)
        DD = D * D
        gn_t = None                                                  # no normalised copy: backward reads the centred g_t
        need_bwd = any(ctx.needs_input_grad)                         # inference: the pre-activations Z are neither stored nor allocated
        zw = 2 * D * z_dtype(prec).itemsize // T.itemsize            # width of a Z row in T words (fp16 Z in the pair mode)
        tbuf = torch.empty(16 * DD + N * 4 * D + E * 2 * D + (E * zw if need_bwd else 0) + 2 * E * D, dtype=T, device=dev)
        (W1n_t, W1e_t, G2_t, A2_t, W1nT_t, W1eT_t, G2T_t, A2T_t, P, H, s_t, g_t, Z) = _carve(tbuf, [
            (4 * D, D), (2 * D, D), (D, D), (D, D), (D, 4 * D), (D, 2 * D), (D, D), (D, D), (N, 4 * D), (E, 2 * D),
            (E, D), (E, D), (E, zw) if need_bwd else (0, zw)])
        Z = Z.view(z_dtype(prec)) if need_bwd else None
        fbuf = torch.empty(N * D + 9 * D, dtype=torch.float32, device=dev)
        m, mean1, var1, mean2, var2, b1, center, bias_c, hsum = _carve(fbuf, [(N, D), (D,), (D,), (D,), (D,), (2 * D,), (D,), (D,), (D,)])
        x_out = torch.empty(N, D, dtype=torch.float32, device=dev)
        e_out = torch.empty(E, D, dtype=torch.float32, device=dev)
        x_out_t = torch.empty(N, D, dtype=T, device=dev) if shadow else None
        # the T-typed copy of e' only feeds the next layer's GEMMs: the model's last layer skips the 2 B/element store
        want_e_t = shadow and bool(cfg.get("want_e_operand", True))
        e_out_t = torch.empty(E, D, dtype=T, device=dev) if want_e_t else None
        part = ops._partial(dev, int(lib.cartnet_colstats_workspace(2 * D)))
        L = _lib.LayerDesc()
        L.prec, L.training, L.use_envelope, L.D = prec, int(training), int(bool(cfg["use_envelope"])), D
        L.num_nodes, L.num_edges = N, E
        L.radius, L.eps, L.momentum1, L.momentum2 = float(cfg["radius"]), ops.EPS_BN, float(cfg["momentum1"]), float(cfg["momentum2"])
        p = ops._p
        L.src32, L.dst32, L.row_ptr, L.col_ptr, L.perm_src = p(plan.src32), p(plan.dst32), p(plan.row_ptr), p(plan.col_ptr), p(plan.perm_src)
        L.dist, L.x, L.e, L.x_t, L.e_t = p(dist), p(x), p(e), p(x_t), p(e_t)
        params = dict(G1=G1, A1=A1, bg1=bg1, ba1=ba1, G2=G2, A2=A2, bg2=bg2, ba2=ba2, bn1_w=w1, bn1_b=b1n, bn2_w=w2, bn2_b=b2n)
        for k, v in params.items():
            setattr(L, k, p(v.detach()))
        L.bn1_rm, L.bn1_rv, L.bn2_rm, L.bn2_rv = p(cfg["rm1"]), p(cfg["rv1"]), p(cfg["rm2"]), p(cfg["rv2"])
        for k, v in dict(W1n_t=W1n_t, W1e_t=W1e_t, G2_t=G2_t, A2_t=A2_t, W1nT_t=W1nT_t, W1eT_t=W1eT_t, G2T_t=G2T_t, A2T_t=A2T_t,
                         b1=b1, P=P, Z=Z, H=H, g_t=g_t, center=center, bias_c=bias_c, hsum=hsum, s_t=s_t, gn_t=gn_t, m=m, mean1=mean1, var1=var1, mean2=mean2, var2=var2, x_out=x_out,
                         e_out=e_out, x_out_t=x_out_t, e_out_t=e_out_t, partial=part).items():
            setattr(L, k, p(v))
        st = ops._stream()
        _lib.check(lib.cartnet_layer_pack_weights(C.byref(L), st), "layer_pack_weights")
        _lib.check(lib.cartnet_layer_fwd(C.byref(L), st), "layer_fwd")
        cfg["holder"]["x_t"] = x_out_t if shadow else x_out
        cfg["holder"]["e_t"] = e_out_t if shadow else e_out      # None when skipped: a later consumer casts on demand
        ctx.set_materialize_grads(False)     # an unused edge_attr output arrives as None in backward, not as zeros
        ctx.L = L
        ctx.params = tuple(params.values())      # the Parameters themselves: backward may write straight into their .grad
        ctx.keep = (x, e, x_t, e_t, tbuf, fbuf, dist, plan, cfg["rm1"], cfg["rv1"], cfg["rm2"], cfg["rv2"]) + tuple(
            v.detach() for v in params.values())
        ctx.dims = (N, E, D, prec)
        return x_out, e_out

    @staticmethod
    def backward(ctx, dx_out, de_out):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        if ctx.keep is None:
            raise RuntimeError("cartnet_b200: CartNet_layer backward was called a second time, but its saved activations "
                               "have already been freed (run the forward pass again; retain_graph is not supported)")
        ctx.saved_tensors      # noqa: B018 -- version check of x, e and the parameters (see forward)
        L = ctx.L
        N, E, D, prec = ctx.dims
        T = t_dtype(prec)
        dev = ctx.keep[0].device
        dx_out = torch.zeros(N, D, dtype=torch.float32, device=dev) if dx_out is None else dx_out.contiguous()
        de_out = None if de_out is None else de_out.contiguous()      # None: no gradient into e_out (last layer)
        tbuf = torch.empty(3 * E * D + E * 2 * D + N * 4 * D, dtype=T, device=dev)
        ds_t, dg_t, dghat_t, dZ, dP = _carve(tbuf, [(E, D), (E, D), (E, D), (E, 2 * D), (N, 4 * D)])
        fbuf = torch.empty(N * D + 5 * D, dtype=torch.float32, device=dev)
        dm, sums1, sums2 = _carve(fbuf, [(N, D), (3 * D,), (2 * D,)])
        dx_in = torch.empty(N, D, dtype=torch.float32, device=dev)
        de_in = torch.empty(E, D, dtype=torch.float32, device=dev)
        gbuf = torch.empty(8 * D * D + 8 * D, dtype=torch.float32, device=dev)
        (dG1, dA1, dG2, dA2, dbg1, dba1, dbg2, dba2, dw1, db1n, dw2, db2n) = _carve(gbuf, [
            (D, 3 * D), (D, 3 * D), (D, D), (D, D), (D,), (D,), (D,), (D,), (D,), (D,), (D,), (D,)])
        # Parameters whose owner promised that .grad is (re)written once per backward (ddp.FlatGradAllReduce(direct=True)):
        # the kernels store the gradient straight into .grad and autograd gets None -- no AccumulateGrad `+=` launch each
        grads = [dG1, dA1, dbg1, dba1, dG2, dA2, dbg2, dba2, dw1, db1n, dw2, db2n]
        direct = [False] * len(grads)
        for i, prm in enumerate(ctx.params):
            owner = getattr(prm, "_cn_direct_owner", None)
            owner = owner() if owner is not None else None
            tgt = prm.grad if owner is not None and owner.direct else None
            if (tgt is not None and getattr(prm, "_cn_written_epoch", -1) != owner.epoch and tgt.is_contiguous()
                    and tgt.dtype == torch.float32 and tgt.device == dev and tgt.shape == grads[i].shape):
                grads[i], direct[i] = tgt, True
                prm._cn_written_epoch = owner.epoch      # a second backward before the next zero() accumulates via autograd
        (dG1, dA1, dbg1, dba1, dG2, dA2, dbg2, dba2, dw1, db1n, dw2, db2n) = grads
        nbytes = int(lib.cartnet_layer_splitk_bytes(prec, D, N, E))
        ws = ops._workspace(dev, nbytes)
        part = ops._partial(dev, int(lib.cartnet_colstats_workspace(2 * D)))
        p = ops._p
        for k, v in dict(dx_out=dx_out, de_out=de_out, dm=dm, ds_t=ds_t, dg_t=dg_t, dghat_t=dghat_t, dZ=dZ, dP=dP, sums1=sums1,
                         sums2=sums2, dx_in=dx_in, de_in=de_in, dG1=dG1, dA1=dA1, dbg1=dbg1, dba1=dba1, dG2=dG2, dA2=dA2,
                         dbg2=dbg2, dba2=dba2, dbn1_w=dw1, dbn1_b=db1n, dbn2_w=dw2, dbn2_b=db2n, partial=part, splitk=ws).items():
            setattr(L, k, p(v))
        L.splitk_bytes = ws.numel() * 4
        _lib.check(lib.cartnet_layer_bwd(C.byref(L), ops._stream()), "layer_bwd")
        ctx.keep = None
        ctx.params = None
        return (dx_in, de_in) + tuple(None if d else g for g, d in zip(grads, direct)) + (None,)


def cartnet_layer_native(x, e, params, cfg):
    """params = (G1, A1, bg1, ba1, G2, A2, bg2, ba2, bn1_w, bn1_b, bn2_w, bn2_b) in the reference's own layout."""
    return _NativeLayerFn.apply(x, e, *params, cfg)
