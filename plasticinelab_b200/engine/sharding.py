"""Slab decomposition of the particle set across GPUs (host-side logic, numpy only).

The grid is cut along axis 0 into `world` slabs of whole 4-plane blocks.  A particle belongs to the rank whose slab
contains its base cell plane `int(x0 * n_grid - 0.5)` AT PARTITION TIME; ownership then stays fixed for the life of the
env (no migration), and the engine exchanges `halo_w` planes on each side of every boundary (engine/sharded.py).
"""
from __future__ import annotations

import numpy as np


def base_plane(x0: np.ndarray, n_grid: int) -> np.ndarray:
    """Plane index of the stencil base along axis 0: C-style truncation like the kernels (mpm_simulator.py:160)."""
    return (np.asarray(x0, dtype=np.float64) * n_grid - 0.5).astype(np.int64)


def slab_bounds(x0: np.ndarray, n_grid: int, world: int, halo_w: int = 8, align: int = 4):
    """Boundaries b[0]=0 < b[1] < ... < b[world]=n_grid (multiples of `align`) that balance the particle counts, with
    every interior slab at least 2*halo_w planes thick and every boundary at least halo_w planes from the grid faces."""
    assert n_grid % align == 0 and halo_w % align == 0 and world >= 1
    if world == 1:
        return [0, n_grid]
    planes = np.sort(base_plane(x0, n_grid))
    b = [0]
    for r in range(1, world):
        q = planes[min(len(planes) - 1, (len(planes) * r) // world)]
        cut = int(round(q / align)) * align
        lo = max(b[-1] + (2 * halo_w if r > 1 else halo_w), halo_w)
        cut = max(cut, lo)
        cut = min(cut, n_grid - halo_w - 2 * halo_w * (world - 1 - r))
        b.append(cut)
    b.append(n_grid)
    assert all(b[i] < b[i + 1] for i in range(world)), b
    return b


def owned_index(x0: np.ndarray, n_grid: int, bounds, rank: int) -> np.ndarray:
    p = base_plane(x0, n_grid)
    return np.nonzero((p >= bounds[rank]) & (p < bounds[rank + 1]))[0]


def zone(bounds, rank: int, side: int, halo_w: int):
    """Plane range [lo, hi) exchanged with the left (side 0) / right (side 1) neighbour, or None at the domain ends."""
    if side == 0:
        return None if rank == 0 else (bounds[rank] - halo_w, bounds[rank] + halo_w)
    return None if rank == len(bounds) - 2 else (bounds[rank + 1] - halo_w, bounds[rank + 1] + halo_w)


def check_margin(x0: np.ndarray, n_grid: int, bounds, rank: int, halo_w: int) -> bool:
    """True while every stencil of this rank's particles stays inside [own_lo - halo_w, own_hi + halo_w)."""
    if len(x0) == 0:
        return True
    p = base_plane(x0, n_grid)
    lo = bounds[rank] - (halo_w if rank > 0 else 0)
    hi = bounds[rank + 1] + (halo_w if rank < len(bounds) - 2 else 0)
    return bool(p.min() >= lo and p.max() + 2 < hi)
