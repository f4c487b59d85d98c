"""TEST INFRASTRUCTURE ONLY: plain-torch executable specification of every primitive in
cartnet_b200/ops.py (same names, same signatures, same in-place conventions).

Two uses:
  * CPU tests monkeypatch `cartnet_b200.ops` with these functions to check the HOST logic
    (weight packing, the hand-derived backward composition, BatchNorm buffer updates, shadow
    plumbing) against the oracle without a GPU;
  * GPU tests compare each CUDA primitive against the function of the same name here.
The product never imports this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from cartnet_b200.ops import ACT_MUL_DSILU, ACT_NONE, ACT_SILU, PREC_BF16, PREC_BF16X3, PREC_FP32, PREC_TF32, GraphPlan, f32_storage, needs_shadow, t_dtype  # noqa: F401

EPS_BN = 1e-5


def _f(t):
    return t.to(torch.float64) if HIGH else t.to(torch.float32)


HIGH = False   # set True to emulate in fp64 (tolerance budgeting)


def _rna_tf32(t):
    """fp32 -> nearest tf32 (ties away), like cvt.rna.tf32.f32"""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)


def _pair_bf16(t):
    """fp32 -> hi + lo with hi = bf16(t), lo = bf16(t - hi): what a bf16x3 operand pair represents (~16 mantissa bits)"""
    t = t.to(torch.float32)
    hi = t.to(torch.bfloat16).to(torch.float32)
    lo = (t - hi).to(torch.bfloat16).to(torch.float32)
    return hi + lo


def _shadow(t, prec):
    """value -> T-typed storage of precision mode `prec` (the rounding a kernel's store4<T> applies)"""
    if not needs_shadow(prec):
        return t.to(torch.float32)
    if prec == PREC_TF32:
        return _rna_tf32(t.to(torch.float32))
    if prec == PREC_BF16X3:
        return _pair_bf16(t)
    return t.to(t_dtype(prec))


def _dsilu(z):
    s = torch.sigmoid(z)
    return s * (1 + z * (1 - s))


def _env(dist, radius, use_envelope):
    if not use_envelope:
        return torch.ones_like(dist)
    return 0.5 * (torch.cos(dist * math.pi / radius) + 1.0) * (dist < radius)


def _bn(x, mean, var, w, b):
    rstd = 1.0 / torch.sqrt(var + EPS_BN)
    xhat = (x - mean) * rstd
    return xhat * w + b, xhat, rstd


def graph_plan(edge_index, num_nodes, assume_dst_sorted=False):
    E = edge_index.shape[1]
    src, dst = edge_index[0], edge_index[1]
    if ((src < 0) | (src >= num_nodes) | (dst < 0) | (dst >= num_nodes)).any():
        raise IndexError("edge_index out of range")
    perm_dst = None
    if E > 1 and (dst[1:] < dst[:-1]).any():
        perm_dst = torch.sort(dst, stable=True)[1]
        src, dst = src[perm_dst], dst[perm_dst]

    def csr(keys):
        perm = torch.sort(keys, stable=True)[1].to(torch.int32)
        ptr = torch.zeros(num_nodes + 1, dtype=torch.int32)
        ptr[1:] = torch.cumsum(torch.bincount(keys, minlength=num_nodes), 0).to(torch.int32)
        return ptr, perm
    row_ptr, _ = csr(dst)
    col_ptr, perm_src = csr(src)
    return GraphPlan(num_nodes, E, src.to(torch.int32), dst.to(torch.int32), row_ptr, col_ptr, perm_src, perm_dst)


def edge_features(cart_dist, cart_dir, means, betas, cutoff_upper, invariant, ld, prec):
    d = cart_dist.unsqueeze(-1)
    alpha = 5.0 / cutoff_upper
    cut = 0.5 * (torch.cos(d * math.pi / cutoff_upper) + 1.0) * (d < cutoff_upper)
    rbf = cut * torch.exp(-betas * (torch.exp(alpha * (-d)) - means) ** 2)
    parts = [rbf] if invariant else [rbf, cart_dir]
    feat = torch.cat(parts, dim=-1)
    used = feat.shape[1]
    feat = F.pad(feat, (0, ld - used))
    if ld > used:
        feat[:, used] = 1.0          # ones column: the bias gradient rides in the weight-gradient GEMM
    return _shadow(feat, prec)


def gemm(prec, A, B, *, bias=None, gather0=None, gidx0=None, gather1=None, gidx1=None, z_out=None,
         act=ACT_NONE, z_in=None, resid=None, out_f32=None, out_t=None):
    v = _f(A) @ _f(B).t()
    if bias is not None:
        v = v + _f(bias)
    if gather0 is not None:
        v = v + _f(gather0)[gidx0.long()]
    if gather1 is not None:
        v = v + _f(gather1)[gidx1.long()]
    if z_out is not None:      # pre-activations are only used elementwise: fp16 words in the pair mode (the copy rounds)
        z_out.copy_(v.to(torch.float32).clamp(-65504.0, 65504.0) if prec == PREC_BF16X3 else _shadow(v, prec))
    if act == ACT_SILU:
        v = F.silu(v)
    elif act == ACT_MUL_DSILU:
        v = v * _dsilu(_f(z_in))
    if resid is not None:
        v = v + _f(resid)
    if out_f32 is not None:
        out_f32.copy_(v.to(torch.float32))
    if out_t is not None:
        out_t.copy_(_shadow(v, prec))


def gemm_colstats(prec, A, B, bias, out_t, running_mean=None, running_var=None, momentum=0.1, shift=None):
    v = (_f(A) @ _f(B).t() + _f(bias)).float()
    out_t.copy_(_shadow(v, prec))
    return colstats(v, running_mean, running_var, momentum, shift=shift)      # sums are taken before the rounding to T


def gemm_tn(prec, A, B):
    return (_f(A).t() @ _f(B)).to(torch.float32)


def colstats(x, running_mean=None, running_var=None, momentum=0.1, shift=None, prec=PREC_FP32):
    xx = x.to(torch.float64)
    n = x.shape[0]
    mean = xx.mean(0)
    var = (xx * xx).mean(0) - mean * mean
    var = var.clamp(min=0)
    if running_mean is not None:
        unb = var * n / (n - 1) if n > 1 else var
        true_mean = mean if shift is None else mean + shift.double()
        running_mean.copy_(((1 - momentum) * running_mean.double() + momentum * true_mean).float())
        running_var.copy_(((1 - momentum) * running_var.double() + momentum * unb).float())
    return mean.float(), var.float()


def gate_center(H_g, G2, bg2, running_mean, training, prec):
    if not training:
        return (bg2 - running_mean).float(), running_mean.clone()
    E = H_g.shape[0]
    step = E // 4096 if E > 4096 else 1
    rows = (E + step - 1) // step
    hsum = _f(H_g)[::step][:rows].sum(0)
    mu = (_f(G2) @ hsum) / rows
    return (-mu).float(), (_f(bg2) + mu).float()


def colsum(x, prec):
    return x.to(torch.float64).sum(0).float()


def edge_gate_aggregate(g, s, e, dist, row_ptr, num_nodes, bn_mean, bn_var, bn_w, bn_b, radius, use_envelope, prec,
                        want_shadow, want_gn=True):
    mean = torch.zeros_like(bn_var) if bn_mean is None else bn_mean
    ghat, gn, _ = _bn(_f(g), _f(mean), _f(bn_var), _f(bn_w), _f(bn_b))
    sig = _env(_f(dist), radius, use_envelope).unsqueeze(-1) * torch.sigmoid(ghat)
    e_out = (_f(e) + sig).float()
    counts = (row_ptr[1:] - row_ptr[:-1]).long()
    dst = torch.repeat_interleave(torch.arange(num_nodes), counts)
    m = torch.zeros(num_nodes, e.shape[1], dtype=sig.dtype).index_add_(0, dst, sig * _f(s)).float()
    e_t = _shadow(e_out, prec)
    return e_out, e_t, m, (_shadow(gn.float(), prec) if want_gn else None)


def node_update(m, x, bn_mean, bn_var, bn_w, bn_b, prec, want_shadow):
    y, _, _ = _bn(_f(m), _f(bn_mean), _f(bn_var), _f(bn_w), _f(bn_b))
    x_out = (F.silu(y) + _f(x)).float()
    return x_out, _shadow(x_out, prec)


def node_update_bwd(dx_out, m, bn_mean, bn_var, bn_w, bn_b, training):
    y, yhat, rstd = _bn(_f(m), _f(bn_mean), _f(bn_var), _f(bn_w), _f(bn_b))
    dy = _f(dx_out) * _dsilu(y)
    s1, s2 = dy.sum(0), (dy * yhat).sum(0)
    n = m.shape[0]
    corr = (s1 / n + yhat * s2 / n) if training else 0.0
    dm = _f(bn_w) * rstd * (dy - corr)
    return dm.float(), torch.cat([s1, s2]).float()


def edge_gate_bwd(gn_t, s_t, dist, dst32, de_out, dm, bn_var, bn_w, bn_b, radius, use_envelope, training, prec,
                  g_mean=None, input_is_g=False):
    gn = _f(gn_t)
    rstd = 1.0 / torch.sqrt(_f(bn_var) + EPS_BN)
    if input_is_g:           # the stored, centred pre-activation itself: normalise here
        gn = (gn - (0.0 if g_mean is None else _f(g_mean))) * rstd
    sg = torch.sigmoid(gn * _f(bn_w) + _f(bn_b))
    env = _env(_f(dist), radius, use_envelope).unsqueeze(-1)
    dmd = _f(dm)[dst32.long()]
    ds = env * sg * dmd
    dsig = _f(s_t) * dmd if de_out is None else _f(de_out) + _f(s_t) * dmd
    dghat = dsig * env * sg * (1 - sg)
    s1, s2 = dghat.sum(0), (dghat * gn).sum(0)           # sums are taken before dghat is rounded to T
    dghat = _f(_shadow(dghat.float(), prec))
    n = gn.shape[0]
    corr = (s1 / n + gn * s2 / n) if training else 0.0
    dg = _f(bn_w) * rstd * (dghat - corr)
    return _shadow(ds, prec), _shadow(dg, prec), torch.cat([s1, s2, ds.sum(0)]).float()


def segment_sum(x, ptr, perm, num_nodes, out, prec):
    counts = (ptr[1:] - ptr[:-1]).long()
    seg = torch.repeat_interleave(torch.arange(num_nodes), counts)
    n = int(ptr[-1])                       # rows beyond the last segment are not read (the kernels walk the CSR rows only)
    rows = _f(x)[:n] if perm is None else _f(x)[perm.long()[:n]]
    res = torch.zeros(num_nodes, x.shape[1], dtype=rows.dtype).index_add_(0, seg, rows)
    out_is_t = not (out.dtype == torch.float32 and not f32_storage(prec))
    out.copy_(_shadow(res, prec) if out_is_t else res.to(out.dtype))
    return out


def segment_sum_pair(x, row_ptr, col_ptr, perm_src, num_nodes, out, prec):
    C = x.shape[1]
    segment_sum(x, row_ptr, None, num_nodes, out[:, :C], prec)
    segment_sum(x, col_ptr, perm_src, num_nodes, out[:, C:], prec)
    return out


def dsilu_mul(dy, z, prec, want_colsum=False):
    v = _f(dy) * _dsilu(_f(z))
    y = _shadow(v.float(), prec)
    return (y, v.double().sum(0).float()) if want_colsum else y


def cast(x, prec):
    return _shadow(x, prec)


def _chol_L(p6):
    d = F.softplus(p6[:, :3])
    z = torch.zeros_like(d[:, 0])
    return torch.stack([torch.stack([d[:, 0], p6[:, 3], p6[:, 4]], dim=-1),
                        torch.stack([z, d[:, 1], p6[:, 5]], dim=-1),
                        torch.stack([z, z, d[:, 2]], dim=-1)], dim=1)


def cholesky_head_fwd(h, W1, b1):
    p6 = (_f(h) @ _f(W1).t() + _f(b1)).float()
    L = _chol_L(_f(p6))
    return torch.bmm(L.transpose(1, 2), L).float(), p6


def cholesky_head_bwd(dU, h, p6, W1):
    p = _f(p6)
    L = _chol_L(p)
    G = _f(dU)
    dL = torch.bmm(L, G + G.transpose(1, 2))
    dp = torch.stack([dL[:, 0, 0] * torch.sigmoid(p[:, 0]), dL[:, 1, 1] * torch.sigmoid(p[:, 1]), dL[:, 2, 2] * torch.sigmoid(p[:, 2]),
                      dL[:, 0, 1], dL[:, 0, 2], dL[:, 1, 2]], dim=-1)
    return (dp @ _f(W1)).float(), (dp.t() @ _f(h)).float(), dp.sum(0).float()


def loss_l1_mse(pred, true):
    d = pred.double() - true.double()
    return torch.stack([d.abs().mean(), (d * d).mean()]).float()


def loss_l1_mse_bwd(pred, true, dmae, dmse):
    d = pred.double() - true.double()
    n = d.numel()
    ga = 0.0 if dmae is None else dmae.double()
    gq = 0.0 if dmse is None else dmse.double()
    return ((ga * torch.sign(d) + gq * 2.0 * d) / n).float()


ALL = ["loss_l1_mse", "loss_l1_mse_bwd", "cholesky_head_fwd", "cholesky_head_bwd", "graph_plan", "edge_features", "gemm", "gemm_colstats", "gemm_tn", "colstats", "gate_center", "colsum", "edge_gate_aggregate", "node_update",
       "node_update_bwd", "edge_gate_bwd", "segment_sum", "segment_sum_pair", "dsilu_mul", "cast"]


def install(monkeypatch):
    """Route cartnet_b200.ops through this module (CPU host-logic tests only)."""
    import sys
    from cartnet_b200 import functional, ops
    me = sys.modules[__name__]
    monkeypatch.setattr(functional, "USE_NATIVE_LAYER", False)   # Python composition = spec of cartnet_layer_fwd/bwd
    for name in ALL:
        monkeypatch.setattr(ops, name, getattr(me, name))
