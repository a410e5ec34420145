"""Initial particle sampling (host side, numpy).

Must reproduce the reference's particle sets bit for bit (`plb/engine/shapes/shape_maker.py:12-76`), which pins three things:
numpy's GLOBAL generator is seeded with 0 while the shapes are built and restored afterwards; string-valued fields of a
shape spec are Python expressions; and the order of draws per shape -- box: one `random((n,3))`; ball: `normal((n,3))` for the
directions, then `random((n,1))` for the radii (cube root for a uniform ball).
"""
from __future__ import annotations

import contextlib

import numpy as np

DIM = 3
_PALETTE = ((127 << 16) + 127, 127 << 8, 127, 127 << 16)      # default packed RGB per object index


@contextlib.contextmanager
def _seeded_global_rng(seed):
    saved = np.random.get_state()
    np.random.seed(seed)
    try:
        yield
    finally:
        np.random.set_state(saved)


def _quat_to_matrix(q):
    w, x, y, z = (float(c) for c in q)
    n = w * w + x * x + y * y + z * z
    s = 0.0 if n == 0 else 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])


def default_count(volume):
    """10k particles per 0.2^3 of volume, at least one (`get_n_particles`)."""
    return max(int(volume / 0.2 ** 3) * 10000, 1)


def sample_box(center, width, n):
    half = 0.5 * (np.full(DIM, width) if isinstance(width, float) else np.asarray(width, dtype=np.float64))
    if n is None:
        n = default_count(np.prod(2 * half))
    return (np.random.random((n, DIM)) * 2 - 1) * half + np.asarray(center)


def sample_ball(center, radius, n):
    if n is None:
        n = default_count((radius ** 3) * 4 * np.pi / 3)
    direction = np.random.normal(size=(n, DIM))
    direction /= np.linalg.norm(direction, axis=-1, keepdims=True)
    rho = np.random.random(size=(n, 1)) ** (1.0 / DIM)
    return direction * rho * radius + np.asarray(center)[:DIM]


class Shapes:
    """`Shapes(cfg.SHAPES).get()` -> (positions [N,3] float64, packed colours [N] int32), objects concatenated in order."""

    def __init__(self, specs):
        self.objects, self.colors = [], []
        self.dim = DIM
        with _seeded_global_rng(0):
            for spec in specs:
                fields = {k: (eval(v) if isinstance(v, str) else v) for k, v in spec.items() if k != "shape"}
                kind = spec["shape"]
                rot, color = fields.pop("init_rot", None), fields.pop("color", None)
                if kind == "box":
                    pts = sample_box(fields["init_pos"], fields["width"], fields.get("n_particles", 10000))
                elif kind == "sphere":
                    pts = sample_ball(fields["init_pos"], fields["radius"], fields.get("n_particles", 10000))
                else:
                    raise NotImplementedError(f"Shape {kind} is not supported!")
                self._append(pts, color, rot)

    def _append(self, pts, color, rot):
        if rot is not None:                      # rigid rotation about the centroid
            c = pts.mean(axis=0)
            pts = (pts[:, :DIM] - c) @ _quat_to_matrix(rot).T + c
        self.objects.append(pts[:, :DIM])
        if color is None or isinstance(color, int):
            packed = _PALETTE[len(self.objects) - 1] if color is None else color
            color = np.full(len(pts), packed, dtype=np.int32)
        self.colors.append(color)

    def get(self):
        if not self.objects:
            raise AssertionError("please add at least one shape into the scene")
        return np.concatenate(self.objects), np.concatenate(self.colors)
