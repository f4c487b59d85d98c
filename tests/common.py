"""Helpers shared by CPU and GPU tests."""
import hashlib

import numpy as np
import torch

# tests/golden/make_golden.py::MODEL_CASES (kept in sync by test_golden_cases_in_sync)
MODEL_CASES = {
    "adp": ("adp", [24, 41], 21, dict(invariant=False, temperature=True, use_envelope=True,
                                       atom_types=True, cholesky=True), 5.0),
    "jarvis": ("jarvis", [3, 9, 17], 22, dict(invariant=False, temperature=False, use_envelope=True,
                                              atom_types=True, cholesky=False), 5.0),
    "invariant_noenv": ("mp", [7, 12], 23, dict(invariant=True, temperature=True, use_envelope=False,
                                                atom_types=True, cholesky=True), 5.0),
}
DIM_IN, DIM_RBF, NUM_LAYERS = 256, 64, 4


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def graph_case_names(g):
    return sorted({k.split("/")[0] for k in g.files})


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  -- the 'relative' of the north_star tolerances (1e-5 fp32, 2e-3 bf16)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def run_train_step(model, batch0):
    """One fwd+bwd like train/train.py:171-183 with cfg.loss = MAE; then an eval forward."""
    model.train()
    model.zero_grad(set_to_none=True)
    b = batch0.clone()
    pred, true = model(b)
    loss = torch.nn.functional.l1_loss(pred, true)
    loss.mean().backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    bufs = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.eval()
    with torch.no_grad():
        be = batch0.clone()
        pred_eval, _ = model(be)
    return dict(pred=pred.detach(), loss=loss.detach(), x=b.x.detach(), e=b.edge_attr.detach(), grads=grads,
                bufs=bufs, pred_eval=pred_eval.detach())


def check_against_golden(res, gm, name, tol, gtol, tol_x=None):
    """res from run_train_step; gm = tests/golden/model.npz; tol on activations, gtol on gradients."""
    from oracle import fixtures
    pre = name + "/"
    errs = {}
    errs["pred"] = rel_err(res["pred"], torch.from_numpy(gm[pre + "pred"]))
    errs["pred_eval"] = rel_err(res["pred_eval"], torch.from_numpy(gm[pre + "pred_eval"]))
    errs["loss"] = abs(float(res["loss"]) - float(gm[pre + "loss"])) / abs(float(gm[pre + "loss"]))
    if pre + "x_out_rows" in gm.files:
        errs["x"] = rel_err(torch.from_numpy(fixtures.subsample_rows(res["x"].cpu(), 32)), torch.from_numpy(gm[pre + "x_out_rows"]))
    errs["e"] = rel_err(torch.from_numpy(fixtures.subsample_rows(res["e"].cpu(), 32)), torch.from_numpy(gm[pre + "e_out_rows"]))
    for k, v in errs.items():
        lim = tol_x if (k == "x" and tol_x is not None) else tol
        assert v < lim, (name, k, v, lim)
    gkeys = [k[len(pre + "grad/"):] for k in gm.files if k.startswith(pre + "grad/")]
    assert gkeys
    gscale = max(float(gm[pre + "gradnorm/" + k]) for k in gkeys)
    gerrs = {}
    for k in gkeys:
        ref = torch.from_numpy(gm[pre + "grad/" + k])
        g = res["grads"][k].cpu()
        g2 = g.reshape(g.shape[0], -1)
        mine = torch.from_numpy(fixtures.subsample_rows(g2, 8)) if g2.numel() > 4096 else g
        # gradients that are analytically zero (bias in front of a BatchNorm) are rounding noise: compare
        # against the norm of the largest gradient instead of their own magnitude
        err = float((mine.double() - ref.double()).abs().max())
        scale = float(ref.abs().max())
        assert err <= gtol * scale + gtol * 5e-2 * gscale, (name, k, err, scale, gscale)
        gerrs[k] = err / (scale + 1e-30)
        nrm = float(g.double().norm())
        assert abs(nrm - float(gm[pre + "gradnorm/" + k])) <= 10 * gtol * float(gm[pre + "gradnorm/" + k]) + gtol * 1e-1 * gscale, (name, k)
    for k in gm.files:
        if k.startswith(pre + "buf/"):
            key = k[len(pre + "buf/"):]
            ref = torch.from_numpy(gm[k])
            got = res["bufs"][key].cpu()
            if ref.dtype.is_floating_point:
                assert rel_err(got, ref) < tol, (name, key, rel_err(got, ref))
            else:
                assert torch.equal(got, ref), (name, key)
    return errs, gerrs
