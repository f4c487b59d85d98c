"""GPU: the product model (CUDA kernels through the C ABI) against the reference's golden vectors and the
oracle, forward + backward, training and eval mode."""
import numpy as np
import pytest
import torch

import common
import cartnet_b200
from cartnet_b200 import _lib
from oracle import cartnet_oracle as O
from oracle import fixtures

pytestmark = pytest.mark.gpu


def _model(kw, seed, lrad, precision):
    torch.manual_seed(0)
    model = cartnet_b200.CartNet(common.DIM_IN, common.DIM_RBF, common.NUM_LAYERS, radius=lrad, precision=precision, **kw)
    model.load_state_dict(fixtures.make_state_dict(model.state_dict(), seed))
    return model.cuda()


def test_native_library_loaded():
    _lib.load()
    assert any("libcartnet_b200.so" in l for l in open("/proc/self/maps"))
    assert _lib.load().cartnet_device_ok(0) == 1


@pytest.mark.parametrize("name", list(common.MODEL_CASES))
def test_fp32_path_matches_reference_golden(golden_model, name):
    """north_star: ADP tensors / edge features of the whole 4-layer model within 1e-5 relative of the reference,
    training mode (batch statistics) and eval mode, plus gradients and BatchNorm buffers. Node features after four
    compounded training-mode layers get 2e-5: on these cases the REFERENCE's own fp32 result is already 8.2e-6
    away from the exact (fp64) value (BatchNorm over 65 nodes amplifies rounding), so two correct fp32
    implementations cannot be expected closer than that; the per-layer test above holds 1e-5."""
    shape, sizes, seed, kw, lrad = common.MODEL_CASES[name]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes), cholesky=kw["cholesky"],
                                        temperature=kw["temperature"]).to("cuda")
    res = common.run_train_step(_model(kw, seed, lrad, "fp32"), batch0)
    errs, gerrs = common.check_against_golden(res, golden_model, name, tol=1e-5, gtol=2e-4, tol_x=2e-5)
    print(name, errs, max(gerrs.values()))


@pytest.mark.parametrize("name", list(common.MODEL_CASES))
@pytest.mark.parametrize("training", [True, False])
def test_fp32_layer_matches_reference_layer(name, training):
    """north_star, literally: "fp32 node features ... match the reference PyG LAYER within 1e-5 relative".
    Every CartNet_layer (and the edge encoder) is fed the oracle's own inputs -- the oracle is bit-identical to
    the reference's forward (tests/golden/make_golden.py asserts pred error 0.0) -- and its outputs are compared
    layer by layer, so rounding differences cannot compound through the four BatchNorm'd layers."""
    shape, sizes, seed, kw, lrad = common.MODEL_CASES[name]
    batch_cpu = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes), cholesky=kw["cholesky"],
                                           temperature=kw["temperature"])
    torch.manual_seed(0)
    orc = O.OracleCartNet(common.DIM_IN, common.DIM_RBF, common.NUM_LAYERS, layer_radius=lrad, **kw)
    sd = fixtures.make_state_dict(orc.state_dict(), seed)
    orc.load_state_dict(sd)
    model = cartnet_b200.CartNet(common.DIM_IN, common.DIM_RBF, common.NUM_LAYERS, radius=lrad, precision="fp32", **kw)
    model.load_state_dict(sd)
    model.cuda()
    orc.train(training); model.train(training)
    with torch.no_grad():
        bo = orc.encoder(batch_cpu.clone())
        bg = model.encoder(batch_cpu.clone().to("cuda"))
        assert common.rel_err(bg.edge_attr, bo.edge_attr) < 1e-5 and common.rel_err(bg.x, bo.x) < 1e-5
        for lo, lg in zip(orc.layers, model.layers):
            bin_g = batch_cpu.clone().to("cuda")
            bin_g.x, bin_g.edge_attr = bo.x.clone().cuda(), bo.edge_attr.clone().cuda()     # identical inputs
            bo = lo(bo)
            out = lg(bin_g)
            ex, ee = common.rel_err(out.x, bo.x), common.rel_err(out.edge_attr, bo.edge_attr)
            assert ex < 1e-5 and ee < 1e-5, (name, training, ex, ee)
            if training:
                assert common.rel_err(lg.norm.running_var, lo.norm.running_var) < 1e-5
                assert common.rel_err(lg.norm2.running_mean, lo.norm2.running_mean) < 1e-5


@pytest.mark.parametrize("precision", ["bf16x3", "tf32", "bf16"])
def test_tensor_core_path(golden_model, precision):
    """north_star: tensor-core path within 2e-3 relative, validation MAE identical to 3 significant digits.
    bf16x3 (the default tensor-core mode, what bench.py times) is held to 2e-3 in TRAINING mode as well -- prediction and
    every gradient (tests/test_gpu_bf16x3.py runs the same check on all three golden cases). The single-rounding modes
    (tf32, bf16) are eval-mode modes: under batch statistics the edge BatchNorm amplifies their operand rounding ~15x
    (measured 5e-3 / 3e-2 on this case), which is why neither is the default; their training-mode numbers are only
    bounded loosely here as a regression guard and are NOT a parity claim."""
    name = "adp"
    shape, sizes, seed, kw, lrad = common.MODEL_CASES[name]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes)).to("cuda")
    res = common.run_train_step(_model(kw, seed, lrad, precision), batch0)
    gm = golden_model
    assert common.rel_err(res["pred_eval"], torch.from_numpy(gm[name + "/pred_eval"])) < 2e-3
    if precision == "bf16x3":
        common.check_against_golden(res, gm, name, tol=2e-3, gtol=2e-3)
    else:
        assert common.rel_err(res["pred"], torch.from_numpy(gm[name + "/pred"])) < {"tf32": 1e-2, "bf16": 6e-2}[precision]   # regression guard only
    mae_ref = float(np.abs(gm[name + "/pred_eval"] - batch0.y.cpu().numpy()).mean())
    mae = float((res["pred_eval"] - batch0.y).abs().mean())
    assert abs(mae - mae_ref) / mae_ref < 5e-4                 # validation MAE identical to 3 significant digits


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_larger_batch_against_oracle_and_determinism(precision):
    shape, count, seed = "adp", 6, 31
    kw = dict(invariant=False, temperature=True, use_envelope=True, atom_types=True, cholesky=True)
    batch_cpu = fixtures.make_oracle_batch(shape, count, seed)
    torch.manual_seed(0)
    orc = O.OracleCartNet(256, 64, 4, **kw)
    sd = fixtures.make_state_dict(orc.state_dict(), seed)
    orc.load_state_dict(sd)
    ref = common.run_train_step(orc, batch_cpu)
    model = cartnet_b200.CartNet(256, 64, 4, precision=precision, **kw)
    model.load_state_dict(sd)
    model.cuda()
    got = common.run_train_step(model, batch_cpu.clone().to("cuda"))
    tol = 1e-5 if precision == "fp32" else 6e-2       # bf16: regression guard only (training-mode parity is bf16x3's, tests/test_gpu_bf16x3.py)
    assert common.rel_err(got["pred"], ref["pred"]) < tol
    assert common.rel_err(got["pred_eval"], ref["pred_eval"]) < (1e-5 if precision == "fp32" else 2e-3)
    if precision == "fp32":
        for k, g in ref["grads"].items():
            scale = max(float(v.abs().max()) for v in ref["grads"].values())
            assert float((got["grads"][k].cpu() - g).abs().max()) <= 2e-4 * float(g.abs().max()) + 1e-5 * scale, k
    # bit-reproducible: no atomics anywhere on the path
    model2 = cartnet_b200.CartNet(256, 64, 4, precision=precision, **kw)
    model2.load_state_dict(sd)
    model2.cuda()
    got2 = common.run_train_step(model2, batch_cpu.clone().to("cuda"))
    assert torch.equal(got["pred"], got2["pred"])
    for k in got["grads"]:
        assert torch.equal(got["grads"][k], got2["grads"][k]), k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_native_layer_orchestration_is_bit_identical_to_python_composition(monkeypatch, precision):
    """cartnet_layer_fwd / cartnet_layer_bwd (csrc/layer.cu) issue the same kernels in the same order as the Python
    composition of the primitives (cartnet_b200/functional.py::_LayerFn, the version the CPU host-logic tests check
    against the reference goldens): outputs, gradients and BatchNorm buffers must agree bit for bit."""
    from cartnet_b200 import functional as CF
    shape, sizes, seed, kw, lrad = common.MODEL_CASES["adp"]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes)).to("cuda")
    res = {}
    for native in (True, False):
        monkeypatch.setattr(CF, "USE_NATIVE_LAYER", native)
        res[native] = common.run_train_step(_model(kw, seed, lrad, precision), batch0)
    a, b = res[True], res[False]
    assert torch.equal(a["pred"], b["pred"]) and torch.equal(a["pred_eval"], b["pred_eval"]) and torch.equal(a["e"], b["e"])
    for k in b["grads"]:
        assert torch.equal(a["grads"][k], b["grads"][k]), k
    for k in b["bufs"]:
        assert torch.equal(a["bufs"][k], b["bufs"][k]), k


def test_unsorted_edges_and_layer_standalone():
    name = "adp"
    shape, sizes, seed, kw, lrad = common.MODEL_CASES[name]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes)).to("cuda")
    model = _model(kw, seed, lrad, "fp32").eval()
    with torch.no_grad():
        b1 = batch0.clone()
        p1, _ = model(b1)
        perm = torch.randperm(batch0.num_edges, generator=torch.Generator().manual_seed(1)).cuda()
        b2 = batch0.clone()
        b2.edge_index = b2.edge_index[:, perm].contiguous()
        b2.cart_dist, b2.cart_dir = b2.cart_dist[perm], b2.cart_dir[perm]
        p2, _ = model(b2)
    assert common.rel_err(p2, p1) < 1e-5
    assert common.rel_err(b2.edge_attr, b1.edge_attr[perm]) < 1e-5


def test_graph_kernel_feeds_model_end_to_end():
    """radius graph built on the GPU -> model forward == oracle graph -> oracle model."""
    from cartnet_b200 import build_graph, synthetic
    from cartnet_b200.batch import CrystalBatch
    seed = 41
    structs = synthetic.make_structures("adp", 3, seed, sizes=np.array([30, 55, 18]))
    batch_cpu = fixtures.make_oracle_batch("adp", 3, seed, sizes=np.array([30, 55, 18]))
    pos = torch.from_numpy(np.concatenate([s["pos"] for s in structs])).cuda()
    cell = torch.from_numpy(np.stack([s["cell"] for s in structs])).cuda()
    nat = torch.tensor([30, 55, 18]).cuda()
    gr = build_graph(pos, cell, nat, 5.0)
    assert torch.equal(gr["edge_index"].cpu(), batch_cpu.edge_index)
    b = batch_cpu.clone().to("cuda")
    b.edge_index, b.cart_dist, b.cart_dir = gr["edge_index"], gr["cart_dist"], gr["cart_dir"]
    kw = dict(invariant=False, temperature=True, use_envelope=True, atom_types=True, cholesky=True)
    orc = O.OracleCartNet(256, 64, 4, **kw)
    sd = fixtures.make_state_dict(orc.state_dict(), seed)
    orc.load_state_dict(sd); orc.eval()
    model = cartnet_b200.CartNet(256, 64, 4, precision="fp32", **kw)
    model.load_state_dict(sd); model.cuda().eval()
    with torch.no_grad():
        pr, _ = orc(batch_cpu.clone())
        pg, _ = model(b)
    assert common.rel_err(pg, pr) < 1e-5


def test_graph_without_edges_eval():
    """Isolated atoms (no pair within the radius): E = 0 flows through the graph build, the encoder and every layer in
    eval mode; nodes without in-edges receive a zero message (cartnet.py:259-260) and only the BatchNorm shift survives."""
    from cartnet_b200 import build_graph
    from cartnet_b200.batch import CrystalBatch
    pos = torch.tensor([[0.5, 0.5, 0.5], [10.0, 10.0, 10.0]], device="cuda")
    cell = (torch.eye(3, device="cuda") * 20.0)[None]
    gr = build_graph(pos, cell, torch.tensor([2], device="cuda"), 1.0)
    assert gr["edge_index"].shape == (2, 0)
    kw = dict(invariant=False, temperature=True, use_envelope=True, atom_types=True, cholesky=True)
    model = cartnet_b200.CartNet(256, 64, 2, precision="fp32", **kw).cuda().eval()
    orc = O.OracleCartNet(256, 64, 2, **kw)
    orc.load_state_dict(model.state_dict())
    orc.eval()
    fields = dict(x=torch.tensor([6, 8]), batch=torch.zeros(2, dtype=torch.int64), natoms=torch.tensor([2]),
                  temperature=torch.tensor([0.3]), non_H_mask=torch.tensor([True, True]), y=torch.zeros(2, 3, 3),
                  edge_index=gr["edge_index"].cpu(), cart_dist=gr["cart_dist"].cpu(), cart_dir=gr["cart_dir"].cpu())
    with torch.no_grad():
        pr, _ = orc(CrystalBatch(**fields))
        pg, _ = model(CrystalBatch(**fields).to("cuda"))
    assert common.rel_err(pg, pr) < 1e-5
    model.train()
    with pytest.raises(ValueError):                       # BatchNorm over zero edges, like nn.BatchNorm1d
        model(CrystalBatch(**fields).to("cuda"))


@pytest.mark.parametrize("temperature,atom_types", [(True, True), (False, True), (True, False), (False, False)])
@pytest.mark.parametrize("cholesky", [True, False])
def test_encoder_variants_match_oracle(temperature, atom_types, cholesky):
    """All four node-encoder variants of Encoder.forward (cartnet.py:144-154: embedding+temperature, embedding+bias,
    temperature only, single learned vector) with both heads, training step against the oracle (fp32 mode)."""
    seed = 51
    batch = fixtures.make_oracle_batch("mp", 5, seed, cholesky=cholesky, temperature=True)
    if not cholesky:
        batch.y = batch.y.reshape(-1)
    kw = dict(invariant=False, temperature=temperature, use_envelope=True, atom_types=atom_types, cholesky=cholesky)
    torch.manual_seed(0)
    orc = O.OracleCartNet(256, 64, 2, **kw)
    sd = fixtures.make_state_dict(orc.state_dict(), seed)
    orc.load_state_dict(sd)
    model = cartnet_b200.CartNet(256, 64, 2, precision="fp32", **kw)
    model.load_state_dict(sd)
    model.cuda()
    ref = common.run_train_step(orc, batch)
    got = common.run_train_step(model, batch.clone().to("cuda"))
    assert common.rel_err(got["pred"], ref["pred"]) < 1e-5
    assert common.rel_err(got["pred_eval"], ref["pred_eval"]) < 1e-5
    scale = max(float(v.abs().max()) for v in ref["grads"].values())
    for k, g in ref["grads"].items():
        if k.endswith("MLP_gate.2.bias"):
            continue                                  # analytically zero gradient (see tests/test_gpu_configs.py)
        assert float((got["grads"][k].cpu() - g).abs().max()) <= 2e-4 * float(g.abs().max()) + 1e-5 * scale, k


@pytest.mark.parametrize("name", list(common.MODEL_CASES))
def test_tf32_mode_eval_within_2e3_on_every_golden_case(golden_model, name):
    """north_star: "the bf16/TF32 tensor-core path within 2e-3 relative". The tf32 mode (operands rounded to nearest tf32
    where they are produced, fp32 accumulation in TMEM) holds that in eval mode on every reference golden case.
    The golden eval prediction was taken after one training step, so the reference's own post-step BatchNorm buffers
    (golden `buf/*`) are loaded first: the comparison is at exactly the reference's state."""
    shape, sizes, seed, kw, lrad = common.MODEL_CASES[name]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes), cholesky=kw["cholesky"],
                                        temperature=kw["temperature"]).to("cuda")
    model = _model(kw, seed, lrad, "tf32").eval()
    pre = name + "/buf/"
    bufs = {k[len(pre):]: torch.from_numpy(golden_model[k]) for k in golden_model.files if k.startswith(pre)}
    missing, unexpected = model.load_state_dict(bufs, strict=False)
    assert not unexpected and len(bufs) == 6 * common.NUM_LAYERS
    with torch.no_grad():
        pred, _ = model(batch0.clone())
    assert common.rel_err(pred, torch.from_numpy(golden_model[name + "/pred_eval"])) < 2e-3


def test_direct_gradient_writes_are_bit_identical():
    """ddp.FlatGradAllReduce(direct=True): the layer backward stores weight gradients straight into the flat
    gradient buffer (no AccumulateGrad launch per parameter); values must equal the autograd-accumulated ones."""
    from cartnet_b200.ddp import FlatGradAllReduce
    shape, sizes, seed, kw, lrad = common.MODEL_CASES["adp"]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes)).to("cuda")
    flats = []
    for direct in (False, True):
        torch.manual_seed(0)
        model = cartnet_b200.CartNet(256, 64, 4, radius=lrad, precision="bf16", **kw)
        model.load_state_dict(fixtures.make_state_dict(model.state_dict(), seed))
        model.cuda().train()
        sync = FlatGradAllReduce(model.parameters(), direct=direct)
        for _ in range(2):                       # second pass: stale values from the first must be overwritten
            sync.zero()
            pred, true = model(batch0.clone())
            torch.nn.functional.l1_loss(pred, true).backward()
        flats.append(sync.flat.clone())
        assert all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in model.parameters())
        if direct:      # a second backward WITHOUT zero() (gradient accumulation) must add, not overwrite
            pred, true = model(batch0.clone())
            torch.nn.functional.l1_loss(pred, true).backward()
            ref2 = 2.0 * flats[0]
            assert common.rel_err(sync.flat, ref2) < 1e-5
    assert torch.equal(flats[0], flats[1])


@pytest.mark.parametrize("dim_in,layers", [(128, 2), (512, 1)])
def test_other_hidden_widths(dim_in, layers):
    """`--dim_in` other than the default 256 (main.py:141): every kernel shape constraint holds for multiples of 128
    (weight-gradient GEMM on single 128-row accumulators, narrower resident weight slices, D/2-wide head)."""
    shape, sizes, seed, kw, lrad = common.MODEL_CASES["adp"]
    batch_cpu = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes))
    torch.manual_seed(0)
    orc = O.OracleCartNet(dim_in, 64, layers, **kw)
    sd = fixtures.make_state_dict(orc.state_dict(), seed)
    orc.load_state_dict(sd)
    ref = common.run_train_step(orc, batch_cpu)
    # bf16 row: regression guard, not a parity claim (the tensor-core parity mode is bf16x3, tests/test_gpu_bf16x3.py)
    for precision, tol_train, tol_eval in (("fp32", 2e-5, 1e-5), ("bf16", 1e-1, 5e-3)):
        model = cartnet_b200.CartNet(dim_in, 64, layers, precision=precision, **kw)
        model.load_state_dict(sd)
        model.cuda()
        got = common.run_train_step(model, batch_cpu.clone().to("cuda"))
        assert common.rel_err(got["pred"], ref["pred"]) < tol_train, precision
        assert common.rel_err(got["pred_eval"], ref["pred_eval"]) < tol_eval, precision
        if precision == "fp32":
            scale = max(float(v.abs().max()) for v in ref["grads"].values())
            for k, g in ref["grads"].items():
                assert float((got["grads"][k].cpu() - g).abs().max()) <= 2e-4 * float(g.abs().max()) + 1e-5 * scale, k


@pytest.mark.parametrize("precision", ["fp32", "bf16", "bf16x3"])
def test_inference_without_grad_is_bit_identical(precision):
    """Under torch.no_grad() the forward pass neither allocates nor stores the pre-activations it would only need for a
    backward pass (specialised GEMM epilogues without the z store): predictions must not change by a bit."""
    shape, sizes, seed, kw, lrad = common.MODEL_CASES["adp"]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes)).to("cuda")
    model = _model(kw, seed, lrad, precision).eval()
    pa, _ = model(batch0.clone())
    with torch.no_grad():
        bb = batch0.clone()
        pb, _ = model(bb)
    assert torch.equal(pa.detach(), pb)
    model.train()                                   # batch statistics, still no autograd graph
    ref = _model(kw, seed, lrad, precision).train()
    pc, _ = ref(batch0.clone())
    with torch.no_grad():
        pd, _ = model(batch0.clone())
    assert torch.equal(pc.detach(), pd)
