"""GPU: every CUDA primitive against its plain-torch specification (tests/emul_ops.py) on seeded inputs,
called through the C ABI (cartnet_b200/ops.py -> ctypes -> libcartnet_b200.so)."""
import numpy as np
import pytest
import torch

import common
import emul_ops as EM
from cartnet_b200 import ops
from cartnet_b200.ops import ACT_MUL_DSILU, ACT_NONE, ACT_SILU, PREC_BF16, PREC_FP32, PREC_TF32

pytestmark = pytest.mark.gpu

PRECS = [PREC_FP32, PREC_BF16]
GEMM_PRECS = [PREC_FP32, PREC_BF16, PREC_TF32]     # TF32 shares every non-GEMM kernel with FP32 (T = float)


def tol(prec):
    return 2e-5 if prec == PREC_FP32 else 1.5e-2


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed + 1000 * len(shape) + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


def both(fn_name, args_cpu, kw_cpu=None):
    kw_cpu = kw_cpu or {}
    to = lambda v: v.cuda() if torch.is_tensor(v) else v
    ref = getattr(EM, fn_name)(*args_cpu, **kw_cpu)
    got = getattr(ops, fn_name)(*[to(a) for a in args_cpu], **{k: to(v) for k, v in kw_cpu.items()})
    return ref, got


def test_edge_features_golden(golden_feat):
    d = torch.from_numpy(golden_feat["dist"]).cuda()
    means, betas = torch.from_numpy(golden_feat["means"]).cuda(), torch.from_numpy(golden_feat["betas"]).cuda()
    cdir = torch.nn.functional.normalize(rnd(d.numel(), 3), dim=-1).cuda()
    f = ops.edge_features(d, cdir, means, betas, 5.0, False, 72, PREC_FP32)
    ref = torch.from_numpy(golden_feat["rbf"])
    assert float((f[:, :64].cpu() - ref).abs().max()) < 2e-6          # reference RBF values, absolute (values in [0,1])
    assert torch.equal(f[:, 64:67], cdir)
    # K padding: one column of ones (bias gradient rides in the weight-gradient GEMM), then zeros
    assert torch.equal(f[:, 67], torch.ones_like(d)) and float(f[:, 68:].abs().max()) == 0.0
    assert torch.equal(f.cpu(), EM.edge_features(d.cpu(), cdir.cpu(), means.cpu(), betas.cpu(), 5.0, False, 72, PREC_FP32)[:, :72]) or \
        float((f.cpu() - EM.edge_features(d.cpu(), cdir.cpu(), means.cpu(), betas.cpu(), 5.0, False, 72, PREC_FP32)).abs().max()) < 2e-6
    f2 = ops.edge_features(d, None, means, betas, 5.0, True, 64, PREC_FP32)
    assert torch.equal(f2, f[:, :64])


@pytest.mark.parametrize("prec", GEMM_PRECS)
@pytest.mark.parametrize("M,N,K", [(1000, 512, 256), (77, 256, 256), (4096 + 33, 256, 512), (300, 1024, 256), (513, 512, 128)])
def test_gemm_plain(prec, M, N, K):
    T = ops.t_dtype(prec)
    A, B = rnd(M, K, seed=1).to(T), rnd(N, K, seed=2, scale=K ** -0.5).to(T)
    out = torch.empty(M, N, dtype=torch.float32)
    EM.gemm(prec, A, B, out_f32=out)
    outg = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(prec, A.cuda(), B.cuda(), out_f32=outg)
    # fp32 / bf16: identical (rounded) operands, fp32 accumulation; tf32: the tensor core drops 13 mantissa bits
    assert common.rel_err(outg, out) < {PREC_FP32: 2e-6, PREC_BF16: 1e-5, PREC_TF32: 2e-3}[prec]


@pytest.mark.parametrize("prec", GEMM_PRECS)
def test_gemm_fused_epilogues(prec):
    T = ops.t_dtype(prec)
    M, N, K, NN = 1500, 512, 256, 211
    A, B = rnd(M, K, seed=1).to(T), rnd(N, K, seed=2, scale=K ** -0.5).to(T)
    bias = rnd(N, seed=3)
    P = rnd(NN, 2 * N, seed=4).to(T)
    g = torch.Generator().manual_seed(9)
    i0 = torch.randint(0, NN, (M,), generator=g, dtype=torch.int32)
    i1 = torch.randint(0, NN, (M,), generator=g, dtype=torch.int32)
    resid = rnd(M, N, seed=5)
    zin = rnd(M, N, seed=6).to(T)
    big = torch.zeros(M, 2 * N, dtype=T)       # strided output views (ldt = 2N)

    def run(mod, dev):
        mv = lambda t: t.to(dev)
        outs = {}
        # (1) edge-GEMM-1 style: bias + two gathers, store z, silu, store T
        z, h = torch.empty(M, N, dtype=T, device=dev), torch.empty(M, N, dtype=T, device=dev)
        mod.gemm(prec, mv(A), mv(B), bias=mv(bias), gather0=mv(P)[:, :N], gidx0=mv(i0), gather1=mv(P)[:, N:], gidx1=mv(i1),
                 z_out=z, act=ACT_SILU, out_t=h)
        outs["z"], outs["h"] = z, h
        # (2) dgrad style: * silu'(z_in), strided T output
        bb = mv(big).clone()
        mod.gemm(prec, mv(A), mv(B), act=ACT_MUL_DSILU, z_in=mv(zin), out_t=bb[:, N:])
        outs["dz"] = bb
        # (3) residual + fp32 out, A given as a strided column view
        A2 = mv(torch.cat([A, A], dim=1))
        o = torch.empty(M, N, dtype=torch.float32, device=dev)
        mod.gemm(prec, A2[:, K:], mv(B), resid=mv(resid), out_f32=o)
        outs["res"] = o
        return outs
    ref, got = run(EM, "cpu"), run(ops, "cuda")
    for k in ref:
        assert common.rel_err(got[k].float(), ref[k].float()) < (3e-6 if prec == PREC_FP32 else 1e-2), k


@pytest.mark.parametrize("prec", GEMM_PRECS)
@pytest.mark.parametrize("K,M,N", [(5000, 512, 256), (333, 256, 256), (70001, 256, 512), (1200, 1024, 256), (900, 512, 128)])
def test_gemm_tn(prec, K, M, N):
    T = ops.t_dtype(prec)
    A, B = rnd(K, M, seed=1).to(T), rnd(K, N, seed=2).to(T)
    big = torch.cat([A, A], dim=1)
    ref = EM.gemm_tn(prec, A, B)
    got = ops.gemm_tn(prec, big.cuda()[:, M:], B.cuda())
    assert common.rel_err(got, ref) < (2e-3 if prec == PREC_TF32 else 5e-6)
    got2 = ops.gemm_tn(prec, big.cuda()[:, M:], B.cuda())
    assert torch.equal(got, got2)                      # deterministic split-K
    # row blocks scattered into a wider parameter-gradient layout (dG1[:, 2D:3D] / dA1[:, 2D:3D] in cartnet_layer_bwd)
    nb = M // 256
    if nb in (2, 4):
        dest = torch.zeros(nb, 256, 3 * N, device="cuda")
        ops.gemm_tn(prec, big.cuda()[:, M:], B.cuda(), out_blocks=[dest[i, :, N:2 * N] for i in range(nb)])
        assert torch.equal(dest[:, :, N:2 * N].reshape(M, N), got)
        assert float(dest[:, :, :N].abs().max()) == 0.0 and float(dest[:, :, 2 * N:].abs().max()) == 0.0


def test_colstats_and_running_update():
    x = rnd(30011, 256, seed=1) * 0.01 + 3.0           # |mean| >> std: the cancellation-prone case
    rm, rv = rnd(256, seed=2), rnd(256, seed=3).abs() + 0.5
    rm_g, rv_g = rm.cuda(), rv.cuda()
    mean, var = EM.colstats(x, rm, rv, 0.1)
    mg, vg = ops.colstats(x.cuda(), rm_g, rv_g, 0.1)
    assert common.rel_err(mg, mean) < 1e-6 and common.rel_err(vg, var) < 1e-4
    assert common.rel_err(rm_g, rm) < 1e-6 and common.rel_err(rv_g, rv) < 1e-6
    ref = torch.nn.functional.batch_norm(x, None, None, training=True)           # torch's own batch statistics
    mine = (x - mg.cpu()) / torch.sqrt(vg.cpu() + 1e-5)
    assert float((ref - mine).abs().max()) < 2e-3 * float(ref.abs().max())
    assert common.rel_err(ops.colsum(x.cuda(), PREC_FP32), x.double().sum(0).float()) < 1e-6


@pytest.mark.parametrize("prec", PRECS + [PREC_TF32])
def test_colstats_of_centred_T_tensor_with_shift(prec):
    """statistics of a tensor stored centred in T (x = true - shift): mean/var are those of the stored values, the
    running mean tracks true = stored + shift."""
    shift = rnd(256, seed=5) * 3.0
    x = EM.cast(rnd(20003, 256, seed=1) * 0.7 + 0.05, prec)
    rm, rv = rnd(256, seed=2), rnd(256, seed=3).abs() + 0.5
    rm_g, rv_g = rm.cuda(), rv.cuda()
    mean, var = EM.colstats(x, rm, rv, 0.1, shift=shift, prec=prec)
    mg, vg = ops.colstats(x.cuda(), rm_g, rv_g, 0.1, shift=shift.cuda(), prec=prec)
    assert float((mg.cpu() - mean).abs().max()) < 1e-6 and common.rel_err(vg, var) < 1e-5
    assert common.rel_err(rm_g, rm) < 1e-6 and common.rel_err(rv_g, rv) < 1e-6
    assert common.rel_err(rm_g, 0.9 * rnd(256, seed=2) + 0.1 * (x.double().mean(0).float() + shift)) < 1e-5


@pytest.mark.parametrize("prec", GEMM_PRECS)
@pytest.mark.parametrize("M", [1000, 128 * 148 * 3 + 77])
def test_gemm_colstats(prec, M):
    """second Linear of MLP_gate with the BatchNorm sums accumulated in its epilogue (tensor-core modes) / by the
    statistics pass (fp32 mode): output, mean, variance and running-buffer update against the specification."""
    N = K = 256
    T = ops.t_dtype(prec)
    A = EM.cast(torch.nn.functional.silu(rnd(M, 2 * K, seed=1)), prec)[:, K:]          # strided view like H[:, D:]
    B, bias, shift = EM.cast(rnd(N, K, seed=2, scale=K ** -0.5), prec), rnd(N, seed=3) * 0.1, rnd(N, seed=4)
    rm, rv = rnd(N, seed=5), rnd(N, seed=6).abs() + 0.5
    rm_g, rv_g = rm.cuda(), rv.cuda()
    out = torch.empty(M, N, dtype=T)
    mean, var = EM.gemm_colstats(prec, A, B, bias, out, rm, rv, 0.1, shift=shift)
    Ag = EM.cast(torch.nn.functional.silu(rnd(M, 2 * K, seed=1)), prec).cuda()[:, K:]
    outg = torch.empty(M, N, dtype=T, device="cuda")
    mg, vg = ops.gemm_colstats(prec, Ag, B.cuda(), bias.cuda(), outg, rm_g, rv_g, 0.1, shift=shift.cuda())
    t = {PREC_FP32: 2e-6, PREC_BF16: 8e-3, PREC_TF32: 2e-3}[prec]
    assert common.rel_err(outg.float(), out.float()) < t
    ts = {PREC_FP32: 2e-5, PREC_BF16: 1e-4, PREC_TF32: 2e-3}[prec]      # tensor-core modes: fp32 partial sums per warp
    assert float((mg.cpu() - mean).abs().max()) < ts * float(var.max().sqrt()) and common.rel_err(vg, var) < ts
    assert common.rel_err(rm_g, rm) < ts and common.rel_err(rv_g, rv) < ts
    mg2, vg2 = ops.gemm_colstats(prec, Ag, B.cuda(), bias.cuda(), outg, None, None, 0.1)
    assert torch.equal(mg, mg2) and torch.equal(vg, vg2)                                # deterministic reduction order


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("E", [900, 70001])
def test_gate_center(prec, E):
    """centre of the gate pre-activation: eval = running mean; training = bg2 + G2 mean(H_g sample), within a fraction
    of a standard deviation of the true column mean of g (all BatchNorm needs)."""
    D = 256
    Hfull = EM.cast(torch.nn.functional.silu(rnd(E, 2 * D, seed=1) + 0.5), prec)
    H = Hfull[:, :D]
    G2, bg2, rm = rnd(D, D, seed=2, scale=D ** -0.5), rnd(D, seed=3), rnd(D, seed=4)
    for training in (True, False):
        ref = EM.gate_center(H, G2, bg2, rm, training, prec)
        got = ops.gate_center(Hfull.cuda()[:, :D], G2.cuda(), bg2.cuda(), rm.cuda(), training, prec)
        assert float((got[0].cpu() - ref[0]).abs().max()) < 1e-5 and float((got[1].cpu() - ref[1]).abs().max()) < 1e-5
        assert float((got[0].cpu() + got[1].cpu() - bg2).abs().max()) < 1e-5          # bias_c + center = bg2
    g = H.float() @ G2.t() + bg2
    assert float(((got_tr := ops.gate_center(Hfull.cuda()[:, :D], G2.cuda(), bg2.cuda(), rm.cuda(), True, prec))[1].cpu()
                  - g.mean(0)).abs().max()) < 0.25 * float(g.std(0).max())


def _bn_args(D, seed):
    return rnd(D, seed=seed) * 0.3, rnd(D, seed=seed + 1).abs() * 0.5 + 0.2, rnd(D, seed=seed + 2) * 0.2 + 1.0, rnd(D, seed=seed + 3) * 0.2


def _graph(N, E, seed):
    g = torch.Generator().manual_seed(seed)
    dst = torch.sort(torch.randint(0, N, (E,), generator=g))[0]
    dst[dst == 3] = 4                                       # leave node 3 without in-edges
    src = torch.randint(0, N, (E,), generator=g)
    return EM.graph_plan(torch.stack([src, dst]), N)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("use_env", [True, False])
def test_edge_gate_aggregate(prec, use_env):
    N, E, D = 301, 17011, 256
    plan = _graph(N, E, 1)
    g, s, e = EM.cast(rnd(E, D, seed=1), prec), EM.cast(rnd(E, D, seed=2), prec), rnd(E, D, seed=3)
    dist = torch.rand(E, generator=torch.Generator().manual_seed(4)) * 5.5
    mean, var, w, b = _bn_args(D, 5)
    if not use_env:
        mean = None                                        # eval mode: g arrives centred on the running mean
    ref, got = both("edge_gate_aggregate", (g, s, e, dist, plan.row_ptr, N, mean, var, w, b, 5.0, use_env, prec, True))
    tT = 2e-6 if prec == PREC_FP32 else 8e-3
    assert common.rel_err(got[0], ref[0]) < 2e-6
    assert common.rel_err(got[1].float(), ref[1].float()) < tT
    assert common.rel_err(got[2], ref[2]) < 1e-5
    assert common.rel_err(got[3].float(), ref[3].float()) < tT          # normalised gate pre-activation saved for backward
    assert float(got[2][3].abs().max()) == 0.0             # node without in-edges gets exactly 0
    got2 = ops.edge_gate_aggregate(g.cuda(), s.cuda(), e.cuda(), dist.cuda(), plan.row_ptr.cuda(), N,
                                   None if mean is None else mean.cuda(), var.cuda(),
                                   w.cuda(), b.cuda(), 5.0, use_env, prec, True, want_gn=False)
    assert torch.equal(got2[2], got[2]) and got2[3] is None   # deterministic reduction; gn is optional


@pytest.mark.parametrize("training", [True, False])
def test_node_update_fwd_bwd(training):
    N, D = 1234, 256
    m, x, dx = rnd(N, D, seed=1), rnd(N, D, seed=2), rnd(N, D, seed=3)
    mean, var, w, b = _bn_args(D, 7)
    ref, got = both("node_update", (m, x, mean, var, w, b, PREC_BF16, True))
    assert common.rel_err(got[0], ref[0]) < 2e-6 and common.rel_err(got[1].float(), ref[1].float()) < 8e-3
    ref, got = both("node_update_bwd", (dx, m, mean, var, w, b, training))
    assert common.rel_err(got[0], ref[0]) < 1e-5 and common.rel_err(got[1], ref[1]) < 1e-5


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("with_de", [True, False])
def test_edge_gate_bwd(prec, training, with_de):
    N, E, D = 150, 9001, 256
    plan = _graph(N, E, 2)
    gn, s = EM.cast(rnd(E, D, seed=1), prec), EM.cast(rnd(E, D, seed=2), prec)
    de = rnd(E, D, seed=3) if with_de else None            # None: no gradient enters e_out (last layer)
    dm = rnd(N, D, seed=4)
    dist = torch.rand(E, generator=torch.Generator().manual_seed(4)) * 5.5
    _, var, w, b = _bn_args(D, 5)
    ref, got = both("edge_gate_bwd", (gn, s, dist, plan.dst32, de, dm, var, w, b, 5.0, True, training, prec))
    t = 1e-5 if prec == PREC_FP32 else 8e-3
    assert common.rel_err(got[0].float(), ref[0].float()) < t
    assert common.rel_err(got[1].float(), ref[1].float()) < t
    assert common.rel_err(got[2], ref[2]) < 1e-5
    # the layer keeps the stored (centred) g instead of a normalised copy: the kernels normalise on the fly
    g_mean = rnd(D, seed=6, scale=0.3) if training else None
    kw = dict(g_mean=g_mean, input_is_g=True)
    ref, got = both("edge_gate_bwd", (gn, s, dist, plan.dst32, de, dm, var, w, b, 5.0, True, training, prec), kw)
    for i in range(2):
        assert common.rel_err(got[i].float(), ref[i].float()) < t
    assert common.rel_err(got[2], ref[2]) < 2e-5


@pytest.mark.parametrize("prec", PRECS)
def test_segment_sum_dst_and_src(prec):
    N, E, C = 97, 6007, 512
    plan = _graph(N, E, 3)
    T = ops.t_dtype(prec)
    x = rnd(E, C, seed=1).to(T)
    for ptr, perm in ((plan.row_ptr, None), (plan.col_ptr, plan.perm_src)):
        ref = EM.segment_sum(x, ptr, perm, N, torch.empty(N, 2 * C, dtype=T)[:, C:], prec)
        out = torch.zeros(N, 2 * C, dtype=T, device="cuda")
        got = ops.segment_sum(x.cuda(), ptr.cuda(), None if perm is None else perm.cuda(), N, out[:, C:], prec)
        assert common.rel_err(got.float(), ref.float()) < (1e-5 if prec == PREC_FP32 else 8e-3)
        assert float(out[:, :C].abs().max()) == 0.0


@pytest.mark.parametrize("prec", PRECS)
def test_segment_sum_pair(prec):
    N, E, C = 97, 6007, 512
    plan = _graph(N, E, 3)
    T = ops.t_dtype(prec)
    x = rnd(E, C, seed=1).to(T)
    ref = EM.segment_sum_pair(x, plan.row_ptr, plan.col_ptr, plan.perm_src, N, torch.empty(N, 2 * C, dtype=T), prec)
    got = ops.segment_sum_pair(x.cuda(), plan.row_ptr.cuda(), plan.col_ptr.cuda(), plan.perm_src.cuda(), N,
                               torch.empty(N, 2 * C, dtype=T, device="cuda"), prec)
    assert common.rel_err(got.float(), ref.float()) < (1e-5 if prec == PREC_FP32 else 8e-3)
    one = ops.segment_sum(x.cuda(), plan.col_ptr.cuda(), plan.perm_src.cuda(), N, torch.empty(N, C, dtype=T, device="cuda"), prec)
    # the pair kernel adds a segment's rows on 8 interleaved row lanes (a different, equally fixed order)
    assert common.rel_err(one.float(), got[:, C:].float()) < (2e-6 if prec == PREC_FP32 else 8e-3)
    again = ops.segment_sum_pair(x.cuda(), plan.row_ptr.cuda(), plan.col_ptr.cuda(), plan.perm_src.cuda(), N,
                                 torch.empty(N, 2 * C, dtype=T, device="cuda"), prec)
    assert torch.equal(again, got)                          # deterministic


@pytest.mark.parametrize("prec", PRECS)
def test_elementwise(prec):
    T = ops.t_dtype(prec)
    dy, z = rnd(777, 512, seed=1), rnd(777, 512, seed=2).to(T)
    ref, got = both("dsilu_mul", (dy, z, prec))
    assert common.rel_err(got.float(), ref.float()) < (2e-6 if prec == PREC_FP32 else 8e-3)
    (ref2, cs_ref), (got2, cs) = both("dsilu_mul", (dy, z, prec), dict(want_colsum=True))     # bias gradient from the same pass
    assert torch.equal(got2, got) and common.rel_err(cs, cs_ref) < 1e-5
    ref, got = both("cast", (dy, prec))
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_linear_silu_matches_torch(precision):
    """Node branch of the encoder (cartnet.py:125-127): y = SiLU(x W^T + b) on the library GEMMs with the hand-written
    backward, against eager torch fp32. bf16 mode runs this node-side GEMM on the tf32 path."""
    from cartnet_b200 import functional as CF
    prec = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16": ops.PREC_BF16}[precision]
    M, K, N = 777, 512, 256
    x = rnd(M, K, seed=3).cuda().requires_grad_(True)
    W = rnd(N, K, seed=4, scale=K ** -0.5).cuda().requires_grad_(True)
    b = rnd(N, seed=5, scale=0.1).cuda().requires_grad_(True)
    dy = rnd(M, N, seed=6).cuda()
    y = CF.linear_silu(x, W, b, prec)
    y.backward(dy)
    got = [y.detach(), x.grad.clone(), W.grad.clone(), b.grad.clone()]
    x.grad = W.grad = b.grad = None
    yr = torch.nn.functional.silu(torch.nn.functional.linear(x, W, b))
    yr.backward(dy)
    tol = 2e-6 if precision == "fp32" else 2e-3
    for g, r in zip(got, [yr.detach(), x.grad, W.grad, b.grad]):
        assert common.rel_err(g, r) < tol


@pytest.mark.parametrize("n,Dh", [(1, 128), (777, 128), (5000, 128), (33, 64), (40, 256), (0, 128)])
def test_cholesky_head_tail(n, Dh):
    """cartnet_cholesky_head_fwd / _bwd (cartnet.py:293-303: Linear(Dh, 6), softplus diagonal, upper-triangular L,
    U = L^T L) against the plain-torch specification, and against autograd through that specification."""
    h, W1, b1 = rnd(n, Dh, seed=1), rnd(6, Dh, seed=2, scale=Dh ** -0.5), rnd(6, seed=3)
    b1[0] = 25.0 if n else b1[0]                           # exercises the softplus threshold branch
    dU = rnd(n, 3, 3, seed=4)
    U_ref, p_ref = EM.cholesky_head_fwd(h, W1, b1)
    U, p6 = ops.cholesky_head_fwd(h.cuda(), W1.cuda(), b1.cuda())
    if n == 0:
        assert U.shape == (0, 3, 3)
        dh, dW1, db1 = ops.cholesky_head_bwd(dU.cuda(), h.cuda(), p6, W1.cuda())
        assert float(dW1.abs().max()) == 0.0 and float(db1.abs().max()) == 0.0
        return
    assert common.rel_err(p6, p_ref) < 2e-6 and common.rel_err(U, U_ref) < 5e-6
    assert torch.equal(U, U.transpose(1, 2))               # exactly symmetric, like L^T L in the reference
    hh, ww, bb = (t.clone().double().requires_grad_(True) for t in (h, W1, b1))
    pa = hh @ ww.t() + bb
    La = EM._chol_L(pa)
    (torch.bmm(La.transpose(1, 2), La) * dU.double()).sum().backward()
    dh, dW1, db1 = ops.cholesky_head_bwd(dU.cuda(), h.cuda(), p6, W1.cuda())
    for got, ref in ((dh, hh.grad), (dW1, ww.grad), (db1, bb.grad)):
        assert common.rel_err(got, ref) < 1e-5
    dh2, dW2, db2 = ops.cholesky_head_bwd(dU.cuda(), h.cuda(), p6, W1.cuda())
    assert torch.equal(dW1, dW2) and torch.equal(db1, db2) and torch.equal(dh, dh2)     # deterministic reduction
