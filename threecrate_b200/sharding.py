"""Host-side sharding logic for the multi-GPU path (one process per GPU).

kNN / normals: the grid is replicated, rank r takes the contiguous range
[r*n/W, (r+1)*n/W) of the index's cell-sorted order (the device snaps it to whole cells).
ICP: rank r takes source points [r*ns/W, (r+1)*ns/W); the 29 normal-equation scalars are
sum-all-reduced every iteration and every rank solves the same 6x6 system.
"""
from __future__ import annotations

import numpy as np

NUM_ICP_SUMS = 29  # 21 upper-triangular AtA + 6 Atb + sum b^2 + n_valid


def shard_range(rank: int, world: int, n: int) -> tuple[int, int]:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * n // world, (rank + 1) * n // world


def normal_equation_sums(src_t: np.ndarray, tgt: np.ndarray, nrm: np.ndarray) -> np.ndarray:
    """The 29 scalars of registration.rs:412-428,461-469 for already-matched, already-transformed
    pairs (f64).  Used by the CPU tests of the all-reduce decomposition."""
    s = src_t.astype(np.float64)
    d = tgt.astype(np.float64)
    n = nrm.astype(np.float64)
    a = np.concatenate([np.cross(s, n), n], axis=1)            # rows [c, n]
    b = (n * (d - s)).sum(1)
    out = np.zeros(NUM_ICP_SUMS)
    t = 0
    for r in range(6):
        for c in range(r, 6):
            out[t] = (a[:, r] * a[:, c]).sum()
            t += 1
    out[21:27] = (a * b[:, None]).sum(0)
    out[27] = (b * b).sum()
    out[28] = len(s)
    return out


def solve_normal_equations(sums: np.ndarray) -> np.ndarray:
    """x = [alpha, beta, gamma, tx, ty, tz] from the 29 scalars."""
    ata = np.zeros((6, 6))
    t = 0
    for r in range(6):
        for c in range(r, 6):
            ata[r, c] = ata[c, r] = sums[t]
            t += 1
    return np.linalg.solve(ata, sums[21:27])
