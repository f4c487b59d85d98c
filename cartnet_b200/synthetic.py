"""Synthetic crystal generator shared by tests and bench.py.

Host-side numpy only; no reference data set is available offline (CSD is licensed,
JARVIS/MP need a download), so every workload is generated from a seed. The shapes
follow the reference's data sets: ADP 194.2 atoms/crystal on average
(/root/reference/README.md:95), radius 5 Å (/root/reference/main.py:142).
"""
from __future__ import annotations

import numpy as np

_Z_CHOICES = np.array([1, 6, 7, 8, 9, 15, 16, 17, 35], dtype=np.int64)
_Z_PROBS = np.array([0.45, 0.38, 0.06, 0.08] + [0.03 / 5] * 5, dtype=np.float64)


def crystal_sizes(shape: str, count: int, rng: np.random.Generator) -> np.ndarray:
    """Atoms per crystal for the named workload shape."""
    if shape == "adp":  # lognormal, mean ~194 atoms (README.md:95)
        n = np.clip(np.rint(rng.lognormal(5.05, 0.65, size=count)), 16, 1200)
    elif shape == "jarvis":  # small cells, 2..40 atoms
        n = rng.integers(2, 41, size=count)
    elif shape == "mp":  # MEGNet split, 4..80 atoms
        n = rng.integers(4, 81, size=count)
    elif shape == "supercell":
        n = np.full(count, 5000)
    else:
        raise ValueError("unknown shape %r" % shape)
    return n.astype(np.int64)


def density(shape: str) -> float:
    """Å^3 per atom."""
    return 9.5 if shape in ("adp", "supercell") else 15.0


def make_crystal(n: int, rho: float, rng: np.random.Generator):
    """One triclinic crystal: returns (pos[n,3] f32, cell[3,3] f32 rows = lattice vectors)."""
    a = float((n * rho) ** (1.0 / 3.0))
    cell = np.array([[1.2 * a, 0.0, 0.0],
                     [0.3 * a, 0.9 * a, 0.0],
                     [0.1 * a, -0.2 * a, a / 1.08]], dtype=np.float32)
    frac = rng.random((n, 3))
    pos = (frac @ cell.astype(np.float64)).astype(np.float32)
    return pos, cell


def make_structures(shape: str, count: int, seed: int, sizes=None):
    """List of dicts {pos, cell, z, temperature} for `count` crystals."""
    rng = np.random.default_rng(seed)
    if sizes is None:
        sizes = crystal_sizes(shape, count, rng)
    rho = density(shape)
    out = []
    for n in sizes:
        pos, cell = make_crystal(int(n), rho, rng)
        z = rng.choice(_Z_CHOICES, size=int(n), p=_Z_PROBS)
        out.append({"pos": pos, "cell": cell, "z": z,
                    "temperature": np.float32(rng.standard_normal())})
    return out


def adp_targets(num_non_h: int, rng: np.random.Generator) -> np.ndarray:
    """SPD 3x3 targets y = A^T A + 0.005 I per non-H atom."""
    a = rng.normal(0.0, 0.05, size=(num_non_h, 3, 3))
    y = np.einsum("nji,njk->nik", a, a) + 0.005 * np.eye(3)
    return y.astype(np.float32)
