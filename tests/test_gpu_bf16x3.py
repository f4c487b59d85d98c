"""GPU: the split-precision tensor-core mode (precision="bf16x3", CARTNET_PREC_BF16X3): operands are hi|lo bf16 pairs
(~16 mantissa bits), every product is three tcgen05 kind::f16 MMAs with fp32 accumulation in TMEM.

north_star: "the bf16/TF32 tensor-core path within 2e-3 relative" -- in TRAINING mode (batch statistics; the edge
BatchNorm amplifies operand rounding ~15x, which plain bf16 / tf32 operands do not survive) as well as in eval mode,
predictions AND gradients, on every reference golden case."""
import numpy as np
import pytest
import torch

import common
import emul_ops as EM
import cartnet_b200
from cartnet_b200 import ops
from cartnet_b200.ops import ACT_MUL_DSILU, ACT_SILU, PREC_BF16X3 as X3
from oracle import cartnet_oracle as O
from oracle import fixtures

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + 1000 * len(shape) + sum(shape))
    return torch.randn(*shape, generator=g) * scale


def pack(x_cpu):
    """fp32 values (CPU) -> opaque pair buffer on the GPU"""
    return ops.cast(x_cpu.float().cuda().contiguous(), X3)


def unpack(t):
    return ops.uncast(t, X3).cpu()


def test_cast_uncast_roundtrip_is_the_pair_rounding():
    x = rnd(777, 192, seed=3) * torch.logspace(-6, 3, 192)
    got = unpack(pack(x))
    assert torch.equal(got, EM._pair_bf16(x))                      # hi = bf16(x), lo = bf16(x - hi), bit for bit
    assert common.rel_err(got, x) < 2 ** -16
    assert torch.equal(unpack(pack(got)), got)                     # idempotent
    # a column slice at a multiple of 64 is a valid pair tensor
    t = pack(x)
    assert torch.equal(ops.uncast(t[:, 64:128], X3).cpu(), got[:, 64:128])


@pytest.mark.parametrize("M,N,K", [(1000, 512, 256), (77, 256, 256), (4096 + 33, 256, 512), (300, 1024, 256), (513, 512, 128), (200, 256, 1024)])
def test_gemm_plain(M, N, K):
    A, B = EM._pair_bf16(rnd(M, K, seed=1) + 0.5), EM._pair_bf16(rnd(N, K, seed=2, scale=K ** -0.5))
    ref = A.double() @ B.double().t()
    outg = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(X3, pack(A), pack(B), out_f32=outg)
    # identical operands; the kernel drops lo*lo (2^-18) and accumulates in fp32
    assert common.rel_err(outg, ref) < 2e-5
    # plain bf16 operands would be ~100x worse
    bf = A.to(torch.bfloat16).double() @ B.to(torch.bfloat16).double().t()
    assert common.rel_err(bf, ref) > 20 * common.rel_err(outg, ref)


def test_gemm_fused_epilogues():
    M, N, K, NN = 1500, 512, 256, 211
    A, B = EM._pair_bf16(rnd(M, K, seed=1)), EM._pair_bf16(rnd(N, K, seed=2, scale=K ** -0.5))
    bias = rnd(N, seed=3)
    P = EM._pair_bf16(rnd(NN, 2 * N, seed=4))
    g = torch.Generator().manual_seed(9)
    i0 = torch.randint(0, NN, (M,), generator=g, dtype=torch.int32)
    i1 = torch.randint(0, NN, (M,), generator=g, dtype=torch.int32)
    resid = rnd(M, N, seed=5)
    zin = rnd(M, N, seed=6).half()                  # stored pre-activations are fp16 in this mode
    # reference (emulation, CPU)
    z_r, h_r = torch.empty(M, N, dtype=torch.float16), torch.empty(M, N)
    EM.gemm(X3, A, B, bias=bias, gather0=P[:, :N], gidx0=i0, gather1=P[:, N:], gidx1=i1, z_out=z_r, act=ACT_SILU, out_t=h_r)
    dz_r = torch.zeros(M, 2 * N)
    EM.gemm(X3, A, B, act=ACT_MUL_DSILU, z_in=zin, out_t=dz_r[:, N:])
    o_r = torch.empty(M, N)
    EM.gemm(X3, A, B, resid=resid, out_f32=o_r)
    # device
    Ag, Bg, Pg = pack(A), pack(B), P.cuda()          # gathered operands are plain fp32 in this mode
    z = torch.empty(M, N, dtype=torch.float16, device="cuda")
    h = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(X3, Ag, Bg, bias=bias.cuda(), gather0=Pg[:, :N], gidx0=i0.cuda(), gather1=Pg[:, N:], gidx1=i1.cuda(), z_out=z,
             act=ACT_SILU, out_t=h)
    big = pack(torch.zeros(M, 2 * N))
    ops.gemm(X3, Ag, Bg, act=ACT_MUL_DSILU, z_in=zin.cuda(), out_t=big[:, N:])        # z: fp16 in this mode
    A2 = pack(torch.cat([A, A], dim=1))
    o = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(X3, A2[:, K:], Bg, resid=resid.cuda(), out_f32=o)
    assert common.rel_err(z.float(), z_r.float()) < 1e-3    # z_out: fp16 in this mode (1-ulp flips where the fp32 values differ)
    assert z.dtype == torch.float16
    assert common.rel_err(unpack(h), h_r) < 1e-3            # SiLU through tanh.approx (MUFU, ~2^-11)
    got_dz = unpack(big)
    assert float(got_dz[:, :N].abs().max()) == 0.0
    assert common.rel_err(got_dz[:, N:], dz_r[:, N:]) < 1e-3
    assert common.rel_err(o, o_r) < 3e-5


@pytest.mark.parametrize("prec", [X3, ops.PREC_TF32])
@pytest.mark.parametrize("N,K", [(256, 256), (512, 256), (256, 512), (512, 128), (512, 512), (1024, 512), (512, 1024), (128, 256), (192, 64)])
def test_cta_pair_gemms_equal_single_cta_gemms(N, K, prec, monkeypatch):
    """Edge-sized GEMMs (M >= 32768) run as CTA pairs (tcgen05 cta_group::2: M = 256 per MMA, half of the weight slice per
    SM). Same products in the same order per output element -> bit-identical to the single-CTA kernel, for every fused
    epilogue the layer uses, including the BatchNorm sums and a ragged last tile."""
    M, NN = 40000 + 77, 311
    A, B = ops.cast(rnd(M, K, seed=1).cuda(), prec), ops.cast(rnd(N, K, seed=2, scale=K ** -0.5).cuda(), prec)
    bias = rnd(N, seed=3).cuda()
    P = rnd(NN, 2 * N, seed=4).cuda()
    g = torch.Generator().manual_seed(9)
    i0 = torch.randint(0, NN, (M,), generator=g, dtype=torch.int32).cuda()
    i1 = torch.randint(0, NN, (M,), generator=g, dtype=torch.int32).cuda()
    resid = rnd(M, N, seed=5).cuda()
    zin = rnd(M, N, seed=6).cuda().to(ops.z_dtype(prec))

    def run():
        out = {}
        z = torch.empty(M, N, dtype=ops.z_dtype(prec), device="cuda")
        h = torch.empty(M, N, dtype=torch.float32, device="cuda")
        ops.gemm(prec, A, B, bias=bias, gather0=P[:, :N], gidx0=i0, gather1=P[:, N:], gidx1=i1, z_out=z, act=ACT_SILU, out_t=h)
        out["z"], out["h"] = z, h
        dz = torch.zeros(M, 2 * N, dtype=torch.float32, device="cuda")
        ops.gemm(prec, A, B, act=ACT_MUL_DSILU, z_in=zin, out_t=dz[:, N:])
        out["dz"] = dz
        o = torch.empty(M, N, dtype=torch.float32, device="cuda")
        ops.gemm(prec, A, B, resid=resid, out_f32=o)
        out["res"] = o
        t = torch.empty(M, N, dtype=torch.float32, device="cuda")
        ops.gemm(prec, A, B, bias=bias, out_t=t)
        out["t"] = t
        c = torch.empty(M, N, dtype=torch.float32, device="cuda")
        mean, var = ops.gemm_colstats(prec, A, B, bias, c)
        out["c"], out["mean"], out["var"] = c, mean, var
        torch.cuda.synchronize()
        return out
    monkeypatch.setenv("CARTNET_NT_CG", "1")
    single = run()
    monkeypatch.setenv("CARTNET_NT_CG", "2")
    pair = run()
    for k in single:
        if k in ("mean", "var"):       # the per-CTA partial sums are grouped differently
            assert common.rel_err(pair[k], single[k]) < 1e-5, k
        else:
            assert torch.equal(pair[k], single[k]), k
    ref = ops.uncast(A, prec).double() @ ops.uncast(B, prec).double().t() + resid.double()
    assert common.rel_err(pair["res"], ref) < (2e-5 if prec == X3 else 2e-3)


def test_preactivations_are_fp16_and_saturate():
    """z_out / dsilu_mul: fp16 storage (csrc/common.cuh::ZOf). |z| beyond the fp16 range saturates to +-65504, where
    silu' is exactly 1 / 0 -- the stored value is only ever used for that."""
    M, N, K = 300, 128, 64
    A = torch.zeros(M, K); A[:, 0] = torch.linspace(-3e5, 3e5, M)
    B = torch.zeros(N, K); B[:, 0] = 1.0
    z = torch.empty(M, N, dtype=torch.float16, device="cuda")
    h = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(X3, pack(A), pack(B), z_out=z, act=ACT_SILU, out_t=h)
    want = EM._pair_bf16(A)[:, :1].clamp(-65504.0, 65504.0).half().expand(M, N)
    assert torch.isfinite(z).all() and torch.equal(z.cpu(), want)
    dy = rnd(M, N, seed=3)
    got = unpack(ops.dsilu_mul(dy.cuda(), z, X3))
    ref = EM._pair_bf16(dy * EM._dsilu(want.float()))
    assert common.rel_err(got, ref) < 1e-3
    # ordinary magnitudes: the fp16 rounding of z moves dy * silu'(z) by < 2e-4 relative
    zf = rnd(777, 512, seed=2) * 3
    exact = dy.new_tensor(0) + rnd(777, 512, seed=1) * EM._dsilu(zf)
    got = unpack(ops.dsilu_mul(rnd(777, 512, seed=1).cuda(), zf.half().cuda(), X3))
    assert common.rel_err(got, exact) < 2e-4


@pytest.mark.parametrize("K,M,N", [(5000, 512, 256), (333, 256, 256), (70001, 256, 512), (1200, 1024, 256), (900, 512, 128)])
def test_gemm_tn(K, M, N):
    A, B = EM._pair_bf16(rnd(K, M, seed=1)), EM._pair_bf16(rnd(K, N, seed=2))
    ref = A.double().t() @ B.double()
    big = pack(torch.cat([A, A], dim=1))
    got = ops.gemm_tn(X3, big[:, M:], pack(B))
    assert common.rel_err(got, ref) < 2e-5
    assert torch.equal(got, ops.gemm_tn(X3, big[:, M:], pack(B)))           # deterministic split-K
    nb = M // 256
    if nb in (2, 4):
        dest = torch.zeros(nb, 256, 3 * N, device="cuda")
        ops.gemm_tn(X3, big[:, M:], pack(B), out_blocks=[dest[i, :, N:2 * N] for i in range(nb)])
        assert torch.equal(dest[:, :, N:2 * N].reshape(M, N), got)


def test_gemm_colstats_and_colsum():
    M, N, K = 30011, 256, 256
    A, B = EM._pair_bf16(rnd(M, K, seed=1) * 0.05 + 1.0), EM._pair_bf16(rnd(N, K, seed=2, scale=K ** -0.5))
    bias = rnd(N, seed=3)
    rm, rv = rnd(N, seed=4), rnd(N, seed=5).abs() + 0.5
    out_r = torch.empty(M, N)
    rm_r, rv_r = rm.clone(), rv.clone()
    mean_r, var_r = EM.gemm_colstats(X3, A, B, bias, out_r, rm_r, rv_r, 0.1)
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    rm_g, rv_g = rm.cuda(), rv.cuda()
    mean, var = ops.gemm_colstats(X3, pack(A), pack(B), bias.cuda(), out, rm_g, rv_g, 0.1)
    assert common.rel_err(unpack(out), out_r) < 3e-5
    assert common.rel_err(mean, mean_r) < 1e-5 and common.rel_err(var, var_r) < 2e-3
    assert common.rel_err(rm_g, rm_r) < 1e-5 and common.rel_err(rv_g, rv_r) < 1e-4
    assert common.rel_err(ops.colsum(out, X3), unpack(out).double().sum(0).float()) < 1e-6


def _model(kw, seed, lrad, precision):
    torch.manual_seed(0)
    model = cartnet_b200.CartNet(common.DIM_IN, common.DIM_RBF, common.NUM_LAYERS, radius=lrad, precision=precision, **kw)
    model.load_state_dict(fixtures.make_state_dict(model.state_dict(), seed))
    return model.cuda()


@pytest.mark.parametrize("name", list(common.MODEL_CASES))
def test_training_and_eval_within_2e3_of_reference_golden(golden_model, name):
    """The stated tensor-core tolerance (2e-3 relative) on everything the golden fixture holds, in TRAINING mode:
    prediction, loss, node / edge features after four layers, every parameter gradient (2e-3 of its own magnitude plus
    1e-4 of the largest gradient norm for the analytically-zero / cancellation-dominated ones), the BatchNorm running
    buffers, and the eval-mode prediction after the step."""
    shape, sizes, seed, kw, lrad = common.MODEL_CASES[name]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes), cholesky=kw["cholesky"],
                                        temperature=kw["temperature"]).to("cuda")
    res = common.run_train_step(_model(kw, seed, lrad, "bf16x3"), batch0)
    errs, gerrs = common.check_against_golden(res, golden_model, name, tol=2e-3, gtol=2e-3)
    print(name, errs, max(gerrs.values()))
    assert errs["pred"] < 5e-4 and errs["pred_eval"] < 5e-4         # measured ~1e-4 / ~2e-5: regression guard
    mae_ref = float(np.abs(golden_model[name + "/pred_eval"] - batch0.y.cpu().numpy()).mean())
    mae = float((res["pred_eval"] - batch0.y).abs().mean())
    assert abs(mae - mae_ref) / mae_ref < 5e-4                      # validation MAE identical to 3 significant digits


def test_native_layer_orchestration_is_bit_identical_to_python_composition(monkeypatch):
    from cartnet_b200 import functional as CF
    shape, sizes, seed, kw, lrad = common.MODEL_CASES["adp"]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes)).to("cuda")
    res = {}
    for native in (True, False):
        monkeypatch.setattr(CF, "USE_NATIVE_LAYER", native)
        res[native] = common.run_train_step(_model(kw, seed, lrad, "bf16x3"), batch0)
    a, b = res[True], res[False]
    assert torch.equal(a["pred"], b["pred"]) and torch.equal(a["pred_eval"], b["pred_eval"]) and torch.equal(a["e"], b["e"])
    for k in b["grads"]:
        assert torch.equal(a["grads"][k], b["grads"][k]), k


def test_larger_batch_against_oracle_and_determinism():
    shape, count, seed = "adp", 6, 31
    kw = dict(invariant=False, temperature=True, use_envelope=True, atom_types=True, cholesky=True)
    batch_cpu = fixtures.make_oracle_batch(shape, count, seed)
    torch.manual_seed(0)
    orc = O.OracleCartNet(256, 64, 4, **kw)
    sd = fixtures.make_state_dict(orc.state_dict(), seed)
    orc.load_state_dict(sd)
    ref = common.run_train_step(orc, batch_cpu)
    got = []
    for _ in range(2):
        model = cartnet_b200.CartNet(256, 64, 4, precision="bf16x3", **kw)
        model.load_state_dict(sd)
        model.cuda()
        got.append(common.run_train_step(model, batch_cpu.clone().to("cuda")))
    assert common.rel_err(got[0]["pred"], ref["pred"]) < 2e-3
    assert common.rel_err(got[0]["pred_eval"], ref["pred_eval"]) < 2e-3
    scale = max(float(v.norm()) for v in ref["grads"].values())
    for k, g in ref["grads"].items():
        assert float((got[0]["grads"][k].cpu() - g).norm()) <= 2e-3 * scale, k        # 2e-3 of the largest gradient norm
        assert float((got[0]["grads"][k].cpu() - g).abs().max()) <= 2e-3 * float(g.abs().max()) + 1e-4 * scale, k
    assert torch.equal(got[0]["pred"], got[1]["pred"])
    for k in got[0]["grads"]:
        assert torch.equal(got[0]["grads"][k], got[1]["grads"][k]), k


@pytest.mark.parametrize("dim_in,layers", [(128, 2), (512, 1)])
def test_other_hidden_widths(dim_in, layers):
    shape, sizes, seed, kw, lrad = common.MODEL_CASES["adp"]
    batch_cpu = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes))
    torch.manual_seed(0)
    orc = O.OracleCartNet(dim_in, 64, layers, **kw)
    sd = fixtures.make_state_dict(orc.state_dict(), seed)
    orc.load_state_dict(sd)
    ref = common.run_train_step(orc, batch_cpu)
    model = cartnet_b200.CartNet(dim_in, 64, layers, precision="bf16x3", **kw)
    model.load_state_dict(sd)
    model.cuda()
    got = common.run_train_step(model, batch_cpu.clone().to("cuda"))
    assert common.rel_err(got["pred"], ref["pred"]) < 2e-3
    assert common.rel_err(got["pred_eval"], ref["pred_eval"]) < 2e-3


@pytest.mark.parametrize("temperature,atom_types", [(True, True), (False, True), (True, False), (False, False)])
@pytest.mark.parametrize("cholesky", [True, False])
def test_encoder_variants_and_heads(temperature, atom_types, cholesky):
    """All four node-encoder variants (cartnet.py:144-154) with both heads in the default tensor-core mode: training step
    against the oracle at 2e-3 (prediction, eval prediction, every gradient)."""
    seed = 51
    batch = fixtures.make_oracle_batch("mp", 5, seed, cholesky=cholesky, temperature=True)
    if not cholesky:
        batch.y = batch.y.reshape(-1)
    kw = dict(invariant=False, temperature=temperature, use_envelope=True, atom_types=atom_types, cholesky=cholesky)
    torch.manual_seed(0)
    orc = O.OracleCartNet(256, 64, 2, **kw)
    sd = fixtures.make_state_dict(orc.state_dict(), seed)
    orc.load_state_dict(sd)
    model = cartnet_b200.CartNet(256, 64, 2, precision="bf16x3", **kw)
    model.load_state_dict(sd)
    model.cuda()
    ref = common.run_train_step(orc, batch)
    got = common.run_train_step(model, batch.clone().to("cuda"))
    assert common.rel_err(got["pred"], ref["pred"]) < 2e-3
    assert common.rel_err(got["pred_eval"], ref["pred_eval"]) < 2e-3
    scale = max(float(v.norm()) for v in ref["grads"].values())
    for k, g in ref["grads"].items():
        assert float((got["grads"][k].cpu() - g).abs().max()) <= 2e-3 * float(g.abs().max()) + 2e-3 * 5e-2 * scale, k


def test_unsorted_edges_empty_graph_and_invariance():
    """Edge cases in the default mode: an edge_index that is not dst-sorted (the layer sorts and un-sorts), a graph without
    edges in eval mode, and -- with invariant=True -- independence of cart_dir (SURVEY.md §4 property 1), bit for bit."""
    shape, sizes, seed, kw, lrad = common.MODEL_CASES["adp"]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes)).to("cuda")
    model = _model(kw, seed, lrad, "bf16x3").eval()
    with torch.no_grad():
        b1 = batch0.clone()
        p1, _ = model(b1)
        perm = torch.randperm(batch0.num_edges, generator=torch.Generator().manual_seed(1)).cuda()
        b2 = batch0.clone()
        b2.edge_index = b2.edge_index[:, perm].contiguous()
        b2.cart_dist, b2.cart_dir = b2.cart_dist[perm], b2.cart_dir[perm]
        p2, _ = model(b2)
    # same kernels on the same edges; only the order inside a destination row (hence the fp32 summation order) differs
    assert common.rel_err(p2, p1) < 1e-5 and common.rel_err(b2.edge_attr, b1.edge_attr[perm]) < 1e-5
    from cartnet_b200.batch import CrystalBatch
    fields = dict(x=torch.tensor([6, 8]), batch=torch.zeros(2, dtype=torch.int64), natoms=torch.tensor([2]),
                  temperature=torch.tensor([0.3]), non_H_mask=torch.tensor([True, True]), y=torch.zeros(2, 3, 3),
                  edge_index=torch.zeros(2, 0, dtype=torch.int64), cart_dist=torch.zeros(0), cart_dir=torch.zeros(0, 3))
    orc = O.OracleCartNet(256, 64, 4, **kw)
    orc.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    orc.eval()
    with torch.no_grad():
        pr, _ = orc(CrystalBatch(**fields))
        pg, _ = model(CrystalBatch(**fields).to("cuda"))
    assert common.rel_err(pg, pr) < 2e-3
    shape, sizes, seed, kw, lrad = common.MODEL_CASES["invariant_noenv"]
    bi = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes)).to("cuda")
    mi = _model(kw, seed, lrad, "bf16x3").eval()
    with torch.no_grad():
        pa, _ = mi(bi.clone())
        bj = bi.clone()
        bj.cart_dir = torch.randn_like(bj.cart_dir)
        pb, _ = mi(bj)
    assert torch.equal(pa, pb)
