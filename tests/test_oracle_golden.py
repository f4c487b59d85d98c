"""CPU: pins the oracle restatement against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py). Runs on the GPU box too (no /root/reference needed)."""
import numpy as np
import pytest
import torch

import common
from oracle import cartnet_oracle as O
from oracle import fixtures


def test_featurisers_bit_exact(golden_feat):
    d = torch.from_numpy(golden_feat["dist"])
    m, b = O.rbf_params(5.0, 64)
    assert np.array_equal(m.numpy(), golden_feat["means"]) and np.array_equal(b.numpy(), golden_feat["betas"])
    assert np.array_equal(O.exp_normal_smearing(d, m, b, 5.0).numpy(), golden_feat["rbf"])
    assert np.array_equal(O.cosine_cutoff(d, 5.0).numpy(), golden_feat["cutoff"])


def test_graph_oracle_matches_reference(golden_graph):
    g = golden_graph
    names = common.graph_case_names(g)
    assert len(names) >= 10
    for name in names:
        pre = name + "/"
        ei, uc, dist, direc = O.radius_graph_pbc_oracle(g[pre + "pos"], g[pre + "cell"], g[pre + "natoms"],
                                                        float(g[pre + "radius"]), pbc=tuple(bool(v) for v in g[pre + "pbc"]),
                                                        max_num_neighbors_threshold=int(g[pre + "knn"]) or None)
        assert ei.shape[1] == int(g[pre + "num_edges"]), name
        assert common.sha(ei) == str(g[pre + "sha_edge_index"]), name          # bit-exact: edge set and order
        assert common.sha(uc) == str(g[pre + "sha_unit_cell"]), name
        assert common.sha(direc) == str(g[pre + "sha_direction"]), name
        if pre + "dist" in g.files:   # sqrt: reference-as-run used MKL VML (not correctly rounded) -> <= 1 ulp
            ulp = np.abs(dist.view(np.int32).astype(np.int64) - g[pre + "dist"].view(np.int32).astype(np.int64))
            assert ulp.max(initial=0) <= 1, name


def test_graph_edge_cases(golden_graph):
    g = golden_graph
    # self-image edges are kept, zero-distance pairs dropped (SURVEY.md §4 property 5)
    ei = g["cubic2/edge_index"]
    assert (ei[0] == ei[1]).sum() > 0
    assert g["single_atom/edge_index"].shape[1] == 6 and (g["single_atom/edge_index"] == 0).all()
    # sorted by (dst, src)
    for name in common.graph_case_names(g):
        if name + "/edge_index" in g.files:
            e = g[name + "/edge_index"]
            key = e[1] * (1 << 32) + e[0]
            assert np.all(np.diff(key) >= 0), name


def test_graph_oracle_chunking_invariant():
    rng = np.random.default_rng(3)
    from cartnet_b200 import synthetic
    pos, cell = synthetic.make_crystal(150, 9.5, rng)
    a = O.radius_graph_pbc_oracle(pos, cell[None], [150], 5.0, chunk_rows=7)
    b = O.radius_graph_pbc_oracle(pos, cell[None], [150], 5.0, chunk_rows=64)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("name", list(common.MODEL_CASES))
def test_model_oracle_matches_reference(golden_model, name):
    gm = golden_model
    shape, sizes, seed, kw, lrad = common.MODEL_CASES[name]
    assert list(gm[name + "/sizes"]) == sizes and int(gm[name + "/seed"]) == seed
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes), cholesky=kw["cholesky"],
                                        temperature=kw["temperature"])
    assert common.sha(batch0.edge_index.numpy()) == str(gm[name + "/sha_edge_index"])
    torch.manual_seed(0)
    model = O.OracleCartNet(common.DIM_IN, common.DIM_RBF, common.NUM_LAYERS, layer_radius=lrad, **kw)
    model.load_state_dict(fixtures.make_state_dict(model.state_dict(), seed))
    res = common.run_train_step(model, batch0)
    common.check_against_golden(res, gm, name, tol=5e-6, gtol=1e-4)


def test_param_count_matches_readme():
    # README.md:186 "2.5M"; exact count reproduced from the reference (SURVEY.md quick facts)
    model = O.OracleCartNet(256, 64, 4)
    assert sum(p.numel() for p in model.parameters()) == 2498438
