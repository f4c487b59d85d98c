"""CPU (authoring container only: needs /root/reference): the reference's OWN call sites drive the replacement module.

INTEGRATION.md §2 patches one import in models/master.py. Here that patch is applied by aliasing `models.cartnet` to
`cartnet_b200.cartnet`, then the unmodified reference code runs on top of it:
  * models/master.py::create_model()            (master.py:23-33, reads the graphgym cfg like main.py:156-188 fills it)
  * checkpoint written by the REFERENCE model   (train.py:92-95 format) loaded into the replacement (main.py:215-216)
  * train/train.py::train_epoch / eval_epoch    (train.py:148-244: batch.to, model(batch), compute_loss, gradient
                                                 accumulation over 16 iterations, optimizer / OneCycleLR steps, metrics)
  * main.py's Monte-Carlo pattern               (main.py:87-98: clone, rotate cart_dir, compare R^T pred R)
and the same epoch is run with the reference's own CartNet for comparison. The CUDA primitives are replaced by their
plain-torch specification (tests/emul_ops.py) because this container has no GPU; what is checked is the drop-in contract
(constructor, forward(batch) -> (pred, true), in-place batch mutation, state-dict keys, autograd through accumulation)."""
import importlib
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import common
import emul_ops
from oracle import fixtures, ref_loader

import cartnet_b200

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")


class _Logger:
    def __init__(self):
        self.rows = []

    def update_stats(self, **kw):
        self.rows.append({k: v for k, v in kw.items() if k in ("loss", "MAE", "MSE", "lr")})


class _PygLikeBatch(cartnet_b200.CrystalBatch):
    """What train.py needs from a PyG Batch: attribute access, .to(device) (a no-op here: no GPU), .clone()."""

    def to(self, device, non_blocking=False):
        return self


def _loader(n_batches, seed):
    out = []
    for i in range(n_batches):
        b = fixtures.make_oracle_batch("adp", 2, seed + i, sizes=np.array([9, 13]))
        out.append(_PygLikeBatch(**b.__dict__))
    return out


def _fill_cfg(cfg, monkeypatch):
    # the fields main.py:156-188 pokes onto the global graphgym cfg
    for k, v in dict(model="CartNet", dim_in=256, dim_rbf=64, num_layers=2, invariant=False, use_temp=True, envelope=True,
                     use_atom_types=True, radius=5.0, loss="MAE", params_count=0).items():
        monkeypatch.setattr(cfg, k, v, raising=False)
    monkeypatch.setattr(cfg, "dataset", SimpleNamespace(name="ADP"), raising=False)


@pytest.fixture
def reference_modules():
    """Imports the reference through the shim and removes every trace afterwards: once `torch_geometric.graphgym.config`
    is importable, cartnet_b200 reads the GLOBAL cfg.radius / cfg.invariant like the reference does (cartnet.py:156,201),
    which must not leak into the other tests of this process."""
    before_mods, before_path = set(sys.modules), list(sys.path)
    yield ref_loader.load()
    for name in set(sys.modules) - before_mods:
        if name.split(".")[0] in ("torch_geometric", "torch_scatter", "models", "train", "dataset", "wandb"):
            del sys.modules[name]
    sys.path[:] = before_path


def test_reference_call_sites_run_unchanged_on_the_replacement(monkeypatch, tmp_path, reference_modules):
    ref_cartnet, _, _, cfg = reference_modules
    _fill_cfg(cfg, monkeypatch)
    monkeypatch.setattr(torch.nn.Module, "to", lambda self, *a, **k: self)      # `.to("cuda:0")` (master.py:33) without a GPU
    master = importlib.import_module("models.master")
    train = importlib.import_module("train.train")

    # --- the reference's own model: reference checkpoint + reference epoch
    torch.manual_seed(3)
    ref_model = master.create_model()
    assert type(ref_model).__module__ == "models.cartnet"
    ckpt = str(tmp_path / "best.ckpt")
    opt_r = torch.optim.Adam(ref_model.parameters(), lr=1e-3)                    # main.py:208
    torch.save({"model_state": ref_model.state_dict(), "optimizer_state": opt_r.state_dict()}, ckpt)      # train.py:92-95

    def epoch(model, opt):
        loader, logger = _loader(18, 500), _Logger()
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, total_steps=4, pct_start=0.3)      # train.py:59
        train.train_epoch(logger, loader, model, opt, 16, sched)                  # scripts/train_cartnet_adp.sh: accumulation 16
        ev = _Logger()
        train.eval_epoch(ev, _loader(3, 900), model)
        return logger.rows, ev.rows

    rows_r, eval_r = epoch(ref_model, opt_r)

    # --- the one patched import, then the same unmodified call sites
    emul_ops.install(monkeypatch)
    monkeypatch.setitem(sys.modules, "models.cartnet", cartnet_b200.cartnet)
    torch.manual_seed(3)
    model = master.create_model()
    assert type(model).__module__ == "cartnet_b200.cartnet"
    state = torch.load(ckpt)
    model.load_state_dict(state["model_state"])                                  # main.py:215-216, strict
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    opt.load_state_dict(state["optimizer_state"])
    rows, eval_rows = epoch(model, opt)

    assert len(rows) == len(rows_r) == 18
    # iterations 1..16 run on the checkpoint's weights (gradients are being accumulated): same function
    for a, b in zip(rows[:16], rows_r[:16]):
        assert abs(a["loss"] - b["loss"]) <= 1e-5 * abs(b["loss"]) and abs(a["MSE"] - b["MSE"]) <= 1e-5 * abs(b["MSE"])
    # after the optimiser step both keep training; Adam turns fp32 rounding-order noise in near-zero gradients into
    # O(lr) parameter differences, so from here on the two runs agree to ~1e-3, not 1e-5
    for a, b in zip(rows[16:] + eval_rows, rows_r[16:] + eval_r):
        assert abs(a["loss"] - b["loss"]) <= 1e-2 * abs(b["loss"]), (a, b)
    assert rows[15]["lr"] != rows[14]["lr"] and rows[16]["lr"] == rows[15]["lr"]         # the scheduler stepped with the optimiser, after iteration 16

    # --- forward mutates the batch in place like the reference (cartnet.py:154,159,223,225) ...
    b = _loader(1, 77)[0]
    model.eval()
    with torch.no_grad():
        pred, true = model(b)
    assert b.x.dtype == torch.float32 and b.x.shape == (22, 256) and b.edge_attr.shape[1] == 256 and true is b.y
    # ... which is why main.py clones before it rotates (main.py:87-98): Monte-Carlo rotation pattern
    b0 = _loader(1, 77)[0]
    R = cartnet_b200.augment.random_rotations(1, "cpu", torch.Generator().manual_seed(0))[0]
    with torch.no_grad():
        p0, _ = model(b0.clone())
        rot = b0.clone()
        rot.cart_dir = rot.cart_dir @ R                                          # main.py:96
        p1, _ = model(rot)
    assert p0.shape == p1.shape and torch.isfinite(p1).all()
    assert torch.allclose(p1, p1.transpose(1, 2), atol=1e-6)                      # Cholesky head: symmetric (SPD) for any rotation
