"""CPU: device-side SO(3) augmentation (cartnet_b200/augment.py) against the reference's per-sample formulas
(/root/reference/dataset/datasetADP.py:33-39) and the geometric invariants it must keep."""
import numpy as np
import torch

from cartnet_b200 import augment
from oracle import fixtures


def test_random_rotations_are_proper():
    R = augment.random_rotations(64, "cpu", torch.Generator().manual_seed(0), dtype=torch.float64)
    eye = torch.eye(3, dtype=torch.float64).expand(64, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-12)
    assert torch.allclose(torch.linalg.det(R), torch.ones(64, dtype=torch.float64), atol=1e-12)


def test_per_crystal_rotation_matches_per_sample_formulas():
    b = fixtures.make_oracle_batch("adp", 3, 5, sizes=np.array([10, 17, 8]))
    b.non_H_index = torch.nonzero(b.non_H_mask).squeeze(-1)
    ref = b.clone()
    R = augment.random_rotations(3, "cpu", torch.Generator().manual_seed(1))
    augment.rotate_batch_(b, R)
    node_g = ref.batch
    edge_g = node_g[ref.edge_index[1]]
    for g in range(3):
        em = edge_g == g
        assert torch.allclose(b.cart_dir[em], ref.cart_dir[em] @ R[g], atol=1e-6)          # datasetADP.py:36
        assert torch.allclose(b.cell[g], ref.cell[g] @ R[g], atol=1e-5)                      # datasetADP.py:37
        am = node_g[ref.non_H_mask] == g
        assert torch.allclose(b.y[am], R[g].T @ ref.y[am] @ R[g], atol=1e-6)                 # datasetADP.py:35
    # invariants: unit directions stay unit, targets stay symmetric with the same eigenvalues, distances untouched
    assert torch.allclose(b.cart_dir.norm(dim=-1), torch.ones(b.num_edges), atol=1e-5)
    assert torch.equal(b.cart_dist, ref.cart_dist)
    assert torch.allclose(torch.linalg.eigvalsh(b.y.double()), torch.linalg.eigvalsh(ref.y.double()), atol=1e-6)
    # rotated geometry is self-consistent: pos_dst - pos_src - offset still points along cart_dir (checked via norms)
    assert torch.allclose(b.pos.norm(dim=-1), ref.pos.norm(dim=-1), atol=1e-4)


def test_single_rotation_form_matches_montecarlo_usage():
    b = fixtures.make_oracle_batch("adp", 2, 6, sizes=np.array([9, 12]))
    ref = b.clone()
    R = augment.random_rotations(1, "cpu", torch.Generator().manual_seed(2))[0]
    augment.rotate_batch_(b, R)
    assert torch.allclose(b.cart_dir, ref.cart_dir @ R, atol=1e-6)                            # main.py:96
    assert torch.allclose(b.y, R.T @ ref.y @ R, atol=1e-6)                                    # main.py:97
