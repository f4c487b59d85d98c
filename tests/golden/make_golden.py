#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference files (imported from
/root/reference through oracle/ref_loader.py) on seeded synthetic inputs, and checks the
oracle restatement (oracle/cartnet_oracle.py) against them while doing so.

Run in the authoring container only:   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there read the committed .npz files.
"""
from __future__ import annotations

import hashlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cartnet_b200 import synthetic  # noqa: E402
from oracle import cartnet_oracle as O  # noqa: E402
from oracle import fixtures, ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------ graph cases
def graph_cases():
    rng = np.random.default_rng(11)
    cases = {}

    def synth(n, rho, seed):
        return synthetic.make_crystal(n, rho, np.random.default_rng(seed))

    p, c = synth(60, 9.5, 1)
    cases["adp60"] = dict(pos=p, cell=c[None], natoms=[60], radius=5.0)
    p, c = synth(5, 15.0, 2)
    cases["jarvis5"] = dict(pos=p, cell=c[None], natoms=[5], radius=5.0)             # rep>=2 -> C>=45
    p, c = synth(2, 15.0, 3)
    cases["jarvis2"] = dict(pos=p, cell=c[None], natoms=[2], radius=5.0)             # rep 3 -> C=343
    cases["cubic2"] = dict(pos=np.array([[0, 0, 0], [1.5, 1.5, 1.5]], np.float32),
                           cell=(np.eye(3, dtype=np.float32) * 3.0)[None], natoms=[2], radius=5.0)
    cases["single_atom"] = dict(pos=np.array([[0.3, 0.2, 0.1]], np.float32),
                                cell=(np.eye(3, dtype=np.float32) * 4.0)[None], natoms=[1], radius=5.0)
    p, c = synth(20, 12.0, 4)
    p = (p + rng.normal(0, 6.0, p.shape)).astype(np.float32)                          # atoms outside the cell
    cases["unwrapped20"] = dict(pos=p, cell=c[None], natoms=[20], radius=5.0)
    p1, c1 = synth(12, 15.0, 5)
    p2, c2 = synth(33, 9.5, 6)
    cases["batch2"] = dict(pos=np.concatenate([p1, p2]), cell=np.stack([c1, c2]), natoms=[12, 33],
                           radius=5.0)                                                # max-rep over batch
    p, c = synth(40, 9.5, 7)
    cases["radius3"] = dict(pos=p, cell=c[None], natoms=[40], radius=3.0)
    p, c = synth(16, 9.5, 8)
    cases["radius7p3"] = dict(pos=p, cell=c[None], natoms=[16], radius=7.3)
    p, c = synth(300, 9.5, 9)
    cases["adp300"] = dict(pos=p, cell=c[None], natoms=[300], radius=5.0, hash_only=True)
    p, c = synth(700, 9.5, 10)
    cases["adp700"] = dict(pos=p, cell=c[None], natoms=[700], radius=5.0, hash_only=True)
    p, c = synth(8, 40.0, 12)   # sparse: some atoms may have few neighbours
    cases["sparse8"] = dict(pos=p, cell=c[None], natoms=[8], radius=2.5)
    # kNN neighbour cap (dataset/utils.py:215-233,240-360) -- the path compute_knn / the Comformer baselines use
    p, c = synth(60, 9.5, 1)
    cases["knn12_adp60"] = dict(pos=p, cell=c[None], natoms=[60], radius=5.0, knn=12)
    cases["knn25_adp60"] = dict(pos=p, cell=c[None], natoms=[60], radius=5.0, knn=25)
    cases["knn200_adp60"] = dict(pos=p, cell=c[None], natoms=[60], radius=5.0, knn=200)          # nothing exceeds -> all kept
    cases["knn6_cubic2"] = dict(pos=np.array([[0, 0, 0], [1.5, 1.5, 1.5]], np.float32),          # highly degenerate shells
                                cell=(np.eye(3, dtype=np.float32) * 3.0)[None], natoms=[2], radius=5.0, knn=6)
    # non-periodic axes (dataset/utils.py:141-156: rep = 0 along them) via the batch's own pbc field (:67-77)
    p, c = synth(30, 9.5, 13)
    cases["slab_pbc110"] = dict(pos=p, cell=c[None], natoms=[30], radius=5.0, pbc=[True, True, False])
    cases["wire_pbc001"] = dict(pos=p, cell=c[None], natoms=[30], radius=5.0, pbc=[False, False, True])
    cases["molecule_pbc000"] = dict(pos=p, cell=c[None], natoms=[30], radius=5.0, pbc=[False, False, False])
    p1, c1 = synth(12, 15.0, 5)
    p2, c2 = synth(33, 9.5, 6)
    cases["knn16_batch2"] = dict(pos=np.concatenate([p1, p2]), cell=np.stack([c1, c2]), natoms=[12, 33], radius=5.0, knn=16)
    return cases


def run_graph(dutils):
    out = {}
    for name, cs in graph_cases().items():
        data = SimpleNamespace(pos=torch.from_numpy(cs["pos"]), cell=torch.from_numpy(cs["cell"]),
                               natoms=torch.tensor(cs["natoms"], dtype=torch.int64),
                               pbc=torch.tensor([cs.get("pbc", [True, True, True])]))
        ei, uc, dist, direc = dutils.radius_graph_pbc(data, cs["radius"], cs.get("knn"), pbc=[True, True, True])
        ei, uc, dist, direc = ei.numpy(), uc.numpy(), dist.numpy(), direc.numpy()
        oei, ouc, odist, odir = O.radius_graph_pbc_oracle(cs["pos"], cs["cell"], cs["natoms"], cs["radius"],
                                                          pbc=tuple(cs.get("pbc", [True, True, True])),
                                                          max_num_neighbors_threshold=cs.get("knn"))
        assert np.array_equal(ei, oei), name
        assert np.array_equal(uc, ouc), name
        # torch.sqrt on this host is MKL VML (<=0.54 ulp, not correctly rounded): the oracle / CUDA
        # value is the IEEE sqrt of the bit-exact d^2, i.e. within 1 ulp of the reference-as-run.
        # All reference call sites discard this output (figshare_dataset.py:65) and recompute the
        # distance from `direction`, which IS bit-exact.
        ulp = np.abs(dist.view(np.int32).astype(np.int64) - odist.view(np.int32).astype(np.int64))
        assert ulp.max(initial=0) <= 1 and (ulp == 0).mean() > 0.98 if len(ulp) else True, name
        assert np.array_equal(direc.view(np.uint32), odir.view(np.uint32)), name
        # sortedness claim (SURVEY.md §4 property 3)
        key = ei[1].astype(np.int64) * (1 << 40) + ei[0].astype(np.int64) * (1 << 20)
        assert np.all(np.diff(key) >= 0), name
        pre = name + "/"
        out[pre + "pos"], out[pre + "cell"] = cs["pos"], cs["cell"]
        out[pre + "natoms"] = np.asarray(cs["natoms"], np.int64)
        out[pre + "radius"] = np.float64(cs["radius"])
        out[pre + "knn"] = np.int64(cs.get("knn") or 0)
        out[pre + "pbc"] = np.asarray(cs.get("pbc", [True, True, True]), dtype=bool)
        out[pre + "num_edges"] = np.int64(ei.shape[1])
        out[pre + "sha_edge_index"] = sha(ei)
        out[pre + "sha_unit_cell"] = sha(uc)
        out[pre + "sha_dist"] = sha(dist)
        out[pre + "sha_direction"] = sha(direc)
        if not cs.get("hash_only"):
            out[pre + "edge_index"], out[pre + "unit_cell"] = ei, uc
            out[pre + "dist"], out[pre + "direction"] = dist, direc
        print("graph %-12s n=%-4d E=%-6d ok (oracle bit-exact)" % (name, sum(cs["natoms"]), ei.shape[1]))
    np.savez_compressed(os.path.join(GOLD, "graph.npz"), **out)


# ------------------------------------------------------------------ model cases
MODEL_CASES = {
    # name: (shape, sizes, seed, ctor kwargs, layer radius)
    "adp": ("adp", [24, 41], 21, dict(invariant=False, temperature=True, use_envelope=True,
                                       atom_types=True, cholesky=True), 5.0),
    "jarvis": ("jarvis", [3, 9, 17], 22, dict(invariant=False, temperature=False, use_envelope=True,
                                              atom_types=True, cholesky=False), 5.0),
    "invariant_noenv": ("mp", [7, 12], 23, dict(invariant=True, temperature=True, use_envelope=False,
                                                atom_types=True, cholesky=True), 5.0),
}
DIM_IN, DIM_RBF, NUM_LAYERS = 256, 64, 4


def grads_of(model):
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


def run_model(cartnet_mod, cfg):
    out = {}
    for name, (shape, sizes, seed, kw, lrad) in MODEL_CASES.items():
        cfg.radius, cfg.invariant = lrad, kw["invariant"]
        batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes),
                                            cholesky=kw["cholesky"], temperature=kw["temperature"])
        torch.manual_seed(0)
        ref = cartnet_mod.CartNet(DIM_IN, DIM_RBF, NUM_LAYERS, **kw)
        sd = fixtures.make_state_dict(ref.state_dict(), seed)
        ref.load_state_dict(sd)
        orc = O.OracleCartNet(DIM_IN, DIM_RBF, NUM_LAYERS, layer_radius=lrad, **kw)
        orc.load_state_dict(sd)                                   # identical keys

        res = {}
        for tag, model in (("ref", ref), ("orc", orc)):
            model.train()
            b = batch0.clone()
            pred, true = model(b)
            loss = torch.nn.functional.l1_loss(pred, true)       # train/metrics.py:15-28, cfg.loss=MAE
            loss.mean().backward()
            g = grads_of(model)
            sd_after = {k: v.detach().clone() for k, v in model.state_dict().items()}
            model.eval()
            with torch.no_grad():
                be = batch0.clone()
                pred_eval, _ = model(be)
            res[tag] = dict(pred=pred.detach(), loss=loss.detach(), x=b.x.detach(), e=b.edge_attr.detach(),
                            grads=g, sd_after=sd_after, pred_eval=pred_eval, x_eval=be.x if be.x.dim() == 2 else None)

        def rel(a, b):
            return float((a - b).abs().max() / (b.abs().max() + 1e-30))
        r, o = res["ref"], res["orc"]
        assert rel(o["pred"], r["pred"]) < 2e-6, (name, rel(o["pred"], r["pred"]))
        assert rel(o["pred_eval"], r["pred_eval"]) < 2e-6
        for k in r["grads"]:
            # grads that are analytically zero (the bias in front of a BatchNorm) are pure rounding noise
            err = float((o["grads"][k] - r["grads"][k]).abs().max())
            gscale = max(float(v.abs().max()) for v in r["grads"].values())
            assert err < 5e-5 * float(r["grads"][k].abs().max()) + 1e-5 * gscale, (name, k, err, gscale)
        for k in r["sd_after"]:
            if r["sd_after"][k].dtype.is_floating_point:
                assert rel(o["sd_after"][k], r["sd_after"][k]) < 1e-5, (name, k)
            else:
                assert torch.equal(o["sd_after"][k], r["sd_after"][k]), (name, k)
        print("model %-16s N=%d E=%d loss=%.6f  oracle==reference (pred %.1e)" % (
            name, batch0.num_nodes, batch0.num_edges, float(r["loss"]), rel(o["pred"], r["pred"])))

        pre = name + "/"
        out[pre + "sizes"] = np.asarray(sizes, np.int64)
        out[pre + "seed"] = np.int64(seed)
        out[pre + "num_nodes"], out[pre + "num_edges"] = np.int64(batch0.num_nodes), np.int64(batch0.num_edges)
        out[pre + "sha_edge_index"] = sha(batch0.edge_index.numpy())
        out[pre + "pred"] = r["pred"].numpy()
        out[pre + "pred_eval"] = r["pred_eval"].numpy()
        out[pre + "loss"] = r["loss"].numpy()
        if r["x"].dim() == 2:
            out[pre + "x_out_rows"] = fixtures.subsample_rows(r["x"], 32)
        out[pre + "e_out_rows"] = fixtures.subsample_rows(r["e"], 32)
        for k, g in r["grads"].items():
            g2 = g.reshape(g.shape[0], -1)
            out[pre + "grad/" + k] = fixtures.subsample_rows(g2, 8) if g2.numel() > 4096 else g.numpy()
            out[pre + "gradnorm/" + k] = np.float64(g.double().norm())
        for k, v in r["sd_after"].items():
            if "running_" in k or "num_batches" in k:
                out[pre + "buf/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "model.npz"), **out)


# ------------------------------------------------------------------ featuriser cases
def run_featurisers(mutils):
    d = torch.tensor(np.concatenate([np.linspace(0.5, 5.5, 41), [1e-3, 4.999999, 5.0, 5.000001]]).astype(np.float32))
    rbf = mutils.ExpNormalSmearing(0.0, 5.0, 64, False)
    cut = mutils.CosineCutoff(0, 5.0)
    out = {"dist": d.numpy(), "rbf": rbf(d).numpy(), "cutoff": cut(d).numpy(),
           "means": rbf.means.numpy(), "betas": rbf.betas.numpy()}
    m, b = O.rbf_params(5.0, 64)
    assert torch.equal(m, rbf.means) and torch.equal(b, rbf.betas)
    assert torch.equal(O.exp_normal_smearing(d, m, b, 5.0), rbf(d))
    assert torch.equal(O.cosine_cutoff(d, 5.0), cut(d))
    np.savez_compressed(os.path.join(GOLD, "featurisers.npz"), **out)
    print("featurisers ok")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    cartnet_mod, mutils, dutils, cfg = ref_loader.load()
    run_featurisers(mutils)
    run_graph(dutils)
    run_model(cartnet_mod, cfg)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
