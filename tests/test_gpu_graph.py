"""GPU: the neighbour-list kernels against the reference's golden vectors (bit-exact), against the oracle on
seeded inputs, and through size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest
import torch

import common
from cartnet_b200 import radius_graph_pbc, build_graph, synthetic, ops
from oracle import cartnet_oracle as O
from types import SimpleNamespace

pytestmark = pytest.mark.gpu


def _data(pos, cell, natoms, pbc=(True, True, True)):
    return SimpleNamespace(pos=torch.from_numpy(np.asarray(pos)).cuda(), cell=torch.from_numpy(np.asarray(cell)).cuda(),
                           natoms=torch.as_tensor(np.asarray(natoms), dtype=torch.int64).cuda(),
                           pbc=torch.tensor([[bool(v) for v in pbc]]).cuda())


def test_golden_graphs_bit_exact(golden_graph):
    g = golden_graph
    for name in common.graph_case_names(g):
        pre = name + "/"
        ei, uc, dist, direc = radius_graph_pbc(_data(g[pre + "pos"], g[pre + "cell"], g[pre + "natoms"], g[pre + "pbc"]),
                                               float(g[pre + "radius"]), int(g[pre + "knn"]) or None, pbc=[True, True, True])
        ei, uc, dist, direc = ei.cpu().numpy(), uc.cpu().numpy(), dist.cpu().numpy(), direc.cpu().numpy()
        assert ei.shape[1] == int(g[pre + "num_edges"]), name
        assert common.sha(ei) == str(g[pre + "sha_edge_index"]), name
        assert common.sha(uc) == str(g[pre + "sha_unit_cell"]), name
        assert common.sha(direc) == str(g[pre + "sha_direction"]), name
        if pre + "dist" in g.files:
            ulp = np.abs(dist.view(np.int32).astype(np.int64) - g[pre + "dist"].view(np.int32).astype(np.int64))
            assert ulp.max(initial=0) <= 1, name       # reference-as-run used a non-IEEE sqrt (MKL VML)


@pytest.mark.parametrize("shape,count,seed", [("adp", 6, 1), ("jarvis", 24, 3), ("mp", 12, 4)])
def test_batched_build_equals_per_crystal_oracle(shape, count, seed):
    structs = synthetic.make_structures(shape, count, seed)
    pos = np.concatenate([s["pos"] for s in structs])
    cell = np.stack([s["cell"] for s in structs])
    nat = [len(s["z"]) for s in structs]
    out = build_graph(torch.from_numpy(pos).cuda(), torch.from_numpy(cell).cuda(), torch.tensor(nat).cuda(), 5.0)
    eis, ucs, dists, dirs = [], [], [], []
    off = 0
    for s in structs:      # the reference's data sets call radius_graph_pbc once per crystal
        ei, uc, d, v = O.radius_graph_pbc_oracle(s["pos"], s["cell"][None], [len(s["z"])], 5.0)
        eis.append(ei + off); ucs.append(uc); dists.append(d); dirs.append(v)
        off += len(s["z"])
    ei = np.concatenate(eis, 1)
    assert np.array_equal(out["edge_index"].cpu().numpy(), ei)
    assert np.array_equal(out["unit_cell"].cpu().numpy(), np.concatenate(ucs))
    assert np.array_equal(out["direction"].cpu().numpy().view(np.uint32), np.concatenate(dirs).view(np.uint32))
    assert np.array_equal(out["dist"].cpu().numpy().view(np.uint32), np.concatenate(dists).view(np.uint32))
    cd, cdir = O.edge_vectors(torch.from_numpy(np.concatenate(dirs)))
    assert common.rel_err(out["cart_dist"], cd) < 1e-6 and common.rel_err(out["cart_dir"], cdir) < 1e-6
    assert np.array_equal(out["src32"].cpu().numpy(), ei[0].astype(np.int32))
    assert np.array_equal(out["dst32"].cpu().numpy(), ei[1].astype(np.int32))
    rp = out["row_ptr"].cpu().numpy()
    assert np.array_equal(np.diff(rp), np.bincount(ei[1], minlength=off))


def test_random_cells_sweep_bit_exact():
    """120 random crystals in one batched launch: skewed triclinic cells, tiny cells that need up to 6 repeats per axis
    (both bmm summation orders, C from 27 to > 1000 cells), atoms outside the cell, radii 2.5..7 -- against the oracle run
    per crystal (what the reference's data sets do)."""
    rng = np.random.default_rng(1234)
    cs = set()
    for radius in (2.5, 5.0, 7.0):
        pos_l, cell_l, nat = [], [], []
        for i in range(40):
            n = int(rng.integers(1, 24))
            a = rng.uniform(1.6, 9.0)
            cell = (np.diag(rng.uniform(0.7, 1.4, 3)) * a + rng.uniform(-0.35, 0.35, (3, 3)) * a * np.tri(3, k=-1)).astype(np.float32)
            if i % 5 == 0:
                cell = cell[[1, 2, 0]] * np.float32(-1 if i % 10 == 0 else 1)          # permuted / left-handed lattices
            frac = rng.uniform(-0.4, 1.4, (n, 3))
            pos_l.append((frac @ cell.astype(np.float64)).astype(np.float32)); cell_l.append(cell); nat.append(n)
        out = build_graph(torch.from_numpy(np.concatenate(pos_l)).cuda(), torch.from_numpy(np.stack(cell_l)).cuda(),
                          torch.tensor(nat).cuda(), radius)
        eis, ucs, dirs = [], [], []
        off = 0
        for p, c, n in zip(pos_l, cell_l, nat):
            ei, uc, d, v = O.radius_graph_pbc_oracle(p, c[None], [n], radius)
            r = O.cell_repeats(c, radius)
            cs.add((2 * r[0] + 1) * (2 * r[1] + 1) * (2 * r[2] + 1))
            eis.append(ei + off); ucs.append(uc); dirs.append(v)
            off += n
        assert np.array_equal(out["edge_index"].cpu().numpy(), np.concatenate(eis, 1))
        assert np.array_equal(out["unit_cell"].cpu().numpy(), np.concatenate(ucs))
        assert np.array_equal(out["direction"].cpu().numpy().view(np.uint32), np.concatenate(dirs).view(np.uint32))
    assert min(cs) == 27 and max(cs) > 400      # both bmm summation orders and large repeat counts were exercised


def test_multi_crystal_call_uses_batch_max_reps():
    """radius_graph_pbc on a 2-crystal batch searches max(rep) cells for both (dataset/utils.py:163)."""
    rng = np.random.default_rng(5)
    p1, c1 = synthetic.make_crystal(4, 15.0, rng)
    p2, c2 = synthetic.make_crystal(50, 9.5, rng)
    pos, cell, nat = np.concatenate([p1, p2]), np.stack([c1, c2]), [4, 50]
    ei, uc, dist, direc = radius_graph_pbc(_data(pos, cell, nat), 5.0, None)
    oei, ouc, odist, odir = O.radius_graph_pbc_oracle(pos, cell, nat, 5.0)
    assert np.array_equal(ei.cpu().numpy(), oei) and np.array_equal(uc.cpu().numpy(), ouc)
    assert np.array_equal(direc.cpu().numpy().view(np.uint32), odir.view(np.uint32))


def test_full_size_properties():
    """ADP-64 (BASELINE configs[1]) and one 5k-atom supercell: properties the oracle is too slow to check
    directly -- sortedness, symmetry of the edge multiset, in-range distances, idempotence."""
    for shape, count, seed in (("adp", 64, 2), ("supercell", 1, 5)):
        structs = synthetic.make_structures(shape, count, seed)
        pos = torch.from_numpy(np.concatenate([s["pos"] for s in structs])).cuda()
        cell = torch.from_numpy(np.stack([s["cell"] for s in structs])).cuda()
        nat = torch.tensor([len(s["z"]) for s in structs]).cuda()
        out = build_graph(pos, cell, nat, 5.0)
        ei, uc, d = out["edge_index"], out["unit_cell"], out["dist"]
        E = ei.shape[1]
        assert E > 40 * pos.shape[0]
        key = ei[1] * (1 << 40) + ei[0] * (1 << 16) + ((uc[:, 0] + 8) * 289 + (uc[:, 1] + 8) * 17 + (uc[:, 2] + 8)).long()
        assert bool((key[1:] > key[:-1]).all())                       # strictly sorted by (dst, src, cell): no duplicates
        assert bool(((d * d) <= 25.0 * (1 + 1e-6)).all()) and bool((d > 0.0099).all())
        rkey = ei[0] * (1 << 40) + ei[1] * (1 << 16) + ((8 - uc[:, 0]) * 289 + (8 - uc[:, 1]) * 17 + (8 - uc[:, 2])).long()
        # (i <- j, u) present  <=>  (j <- i, -u) present, except pairs within 1 ulp of the threshold
        a, b = torch.sort(key)[0], torch.sort(rkey)[0]
        mism = int((a != b).sum())
        assert mism <= max(4, E // 100000), mism
        out2 = build_graph(pos, cell, nat, 5.0)
        assert torch.equal(out2["edge_index"], ei) and torch.equal(out2["direction"], out["direction"])
        # direction consistency: pos[dst] - pos[src] - uc @ cell == direction (fp32 tolerance)
        b_of = torch.repeat_interleave(torch.arange(len(structs), device="cuda"), nat)
        off = torch.einsum("ek,ekd->ed", uc, cell[b_of[ei[1]]])
        recon = pos[ei[1]] - pos[ei[0]] - off
        assert float((recon - out["direction"]).abs().max()) < 1e-3


def test_scan_and_csr_primitives():
    g = torch.Generator().manual_seed(0)
    for n, E in ((1, 1), (37, 500), (5000, 200000)):
        keys = torch.randint(0, n, (E,), generator=g)
        ei = torch.stack([torch.randint(0, n, (E,), generator=g), keys]).cuda()
        plan = ops.graph_plan(ei, n)
        import emul_ops
        ref = emul_ops.graph_plan(ei.cpu(), n)
        assert (plan.perm_dst is None) == (ref.perm_dst is None)
        if ref.perm_dst is not None:
            assert torch.equal(plan.perm_dst.cpu(), ref.perm_dst)
        for f in ("src32", "dst32", "row_ptr", "col_ptr", "perm_src"):
            assert torch.equal(getattr(plan, f).cpu(), getattr(ref, f)), (n, E, f)
    with pytest.raises(IndexError):
        ops.graph_plan(torch.tensor([[0, 5], [1, 0]]).cuda(), 3)


def test_knn_cap_against_oracle_random():
    """kNN cap (dataset/utils.py:240-360) on larger random crystals: non-strict (the reference default, bit-exact by
    construction: a value threshold) and strict (ties by edge order) against the oracle."""
    rng = np.random.default_rng(77)
    for n, k in ((150, 12), (90, 25), (40, 3)):
        pos, cell = synthetic.make_crystal(n, 9.5, rng)
        for strict in (False, True):
            ei, uc, dist, direc = radius_graph_pbc(_data(pos, cell[None], [n]), 5.0, k, enforce_max_neighbors_strictly=strict)
            oei, ouc, odist, odir = O.radius_graph_pbc_oracle(pos, cell[None], [n], 5.0, max_num_neighbors_threshold=k,
                                                              enforce_max_neighbors_strictly=strict)
            assert np.array_equal(ei.cpu().numpy(), oei), (n, k, strict)
            assert np.array_equal(uc.cpu().numpy(), ouc)
            assert np.array_equal(direc.cpu().numpy().view(np.uint32), odir.view(np.uint32))
            deg = np.bincount(oei[1], minlength=n)
            assert deg.min() >= min(k, 1) and (strict is False or deg.max() <= k)


def _build_both(pos, cell, nat, radius=5.0, **kw):
    a = ops.nlist_build(pos, cell, nat, radius, batch_max_reps=False, want_i32=True, cells=True, **kw)
    b = ops.nlist_build(pos, cell, nat, radius, batch_max_reps=False, want_i32=True, cells=False, **kw)
    return a, b


@pytest.mark.parametrize("n,seed", [(700, 11), (2000, 12), (5000, 5), (20000, 6)])
def test_cell_list_is_bit_identical_to_all_pairs(n, seed):
    """north_star kernel (1): the cell-list neighbour kernel (binned, 27 bins per destination, in-row key sort) gives the
    same edges in the same order with the same fp32 payload as the all-pairs kernel -- which is bit-exact against the
    reference goldens above and against the 5000-atom oracle hashes (tests/test_gpu_round2.py)."""
    s = synthetic.make_structures("supercell", 1, seed, sizes=np.array([n]))[0]
    pos, cell, nat = torch.from_numpy(s["pos"]).cuda(), torch.from_numpy(s["cell"][None]).cuda(), torch.tensor([n]).cuda()
    a, b = _build_both(pos, cell, nat)
    assert a["edge_index"].shape[1] > 40 * n
    for k in ("edge_index", "unit_cell", "dist", "direction", "cart_dist", "cart_dir", "row_ptr", "src32", "dst32"):
        assert torch.equal(a[k], b[k]), k


def test_cell_list_mixed_batch_atoms_outside_the_cell_and_partial_pbc():
    """One launch over a mixed batch: a binned 3000-atom crystal, a 40-atom crystal (all-pairs path inside the same
    launch), a 1500-atom crystal whose atoms were moved out of the cell by random lattice translations (the reference
    does not wrap positions: images beyond +-rep are NOT found, utils.py:166-170), a strongly sheared cell; then the same
    batch with a non-periodic axis."""
    rng = np.random.default_rng(3)
    structs = synthetic.make_structures("supercell", 4, 21, sizes=np.array([3000, 40, 1500, 1200]))
    shift = rng.integers(-1, 2, size=(1500, 3)).astype(np.float32)
    structs[2]["pos"] = (structs[2]["pos"] + shift @ structs[2]["cell"]).astype(np.float32)
    c = structs[3]["cell"].copy()
    c[1] += 0.9 * c[0]                                            # shear
    frac = rng.random((1200, 3))
    structs[3]["cell"], structs[3]["pos"] = c, (frac @ c.astype(np.float64)).astype(np.float32)
    pos = torch.from_numpy(np.concatenate([s["pos"] for s in structs])).cuda()
    cell = torch.from_numpy(np.stack([s["cell"] for s in structs])).cuda()
    nat = torch.tensor([len(s["pos"]) for s in structs]).cuda()
    for mask in (7, 3):
        a, b = _build_both(pos, cell, nat, pbc_mask=mask)
        for k in ("edge_index", "unit_cell", "dist", "direction", "row_ptr"):
            assert torch.equal(a[k], b[k]), (mask, k)
    # the shifted crystal against the CPU oracle (which restates the reference's +-rep search literally)
    s2 = structs[2]
    ei, uc, _, direction = O.radius_graph_pbc_oracle(s2["pos"], s2["cell"][None], [1500], 5.0)
    g = ops.nlist_build(torch.from_numpy(s2["pos"]).cuda(), torch.from_numpy(s2["cell"][None]).cuda(), torch.tensor([1500]).cuda(), 5.0,
                        batch_max_reps=False)
    assert np.array_equal(g["edge_index"].cpu().numpy(), ei) and np.array_equal(g["unit_cell"].cpu().numpy(), uc)
    assert np.array_equal(g["direction"].cpu().numpy(), direction)
