"""Initial particle sampling (host side, numpy).

Behaviour follows `plb/engine/shapes/shape_maker.py:12-76`: numpy's global RNG
is seeded with 0 for the duration of the constructor and restored afterwards;
string-valued entries are `eval`'d; boxes are uniform in the box; spheres are a
normalised Gaussian direction times U^(1/3) times radius, drawn in exactly that
order (normal first, then random) so the particle set is identical.
"""
from __future__ import annotations

import numpy as np

COLORS = [(127 << 16) + 127, (127 << 8), 127, 127 << 16]


class Shapes:
    def __init__(self, cfg):
        self.objects = []
        self.colors = []
        self.dim = 3
        state = np.random.get_state()
        np.random.seed(0)
        try:
            for item in cfg:
                kwargs = {k: (eval(v) if isinstance(v, str) else v) for k, v in item.items() if k != "shape"}
                if item["shape"] == "box":
                    self.add_box(**kwargs)
                elif item["shape"] == "sphere":
                    self.add_sphere(**kwargs)
                else:
                    raise NotImplementedError(f"Shape {item['shape']} is not supported!")
        finally:
            np.random.set_state(state)

    def get_n_particles(self, volume):
        return max(int(volume / 0.2 ** 3) * 10000, 1)

    def add_object(self, particles, color=None, init_rot=None):
        if init_rot is not None:
            w, x, y, z = [float(q) for q in init_rot]
            n = w * w + x * x + y * y + z * z
            s = 2.0 / n if n > 0 else 0.0
            R = np.array([
                [1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])
            origin = particles.mean(axis=0)
            particles = (particles[:, :self.dim] - origin) @ R.T + origin
        self.objects.append(particles[:, :self.dim])
        if color is None or isinstance(color, int):
            tmp = COLORS[len(self.objects) - 1] if color is None else color
            color = np.zeros(len(particles), np.int32)
            color[:] = tmp
        self.colors.append(color)

    def add_box(self, init_pos, width, n_particles=10000, color=None, init_rot=None):
        width = np.array([width] * self.dim) if isinstance(width, float) else np.array(width)
        if n_particles is None:
            n_particles = self.get_n_particles(np.prod(width))
        p = (np.random.random((n_particles, self.dim)) * 2 - 1) * (0.5 * width) + np.array(init_pos)
        self.add_object(p, color, init_rot=init_rot)

    def add_sphere(self, init_pos, radius, n_particles=10000, color=None, init_rot=None):
        if n_particles is None:
            volume = (radius ** 3) * 4 * np.pi / 3
            n_particles = self.get_n_particles(volume)
        p = np.random.normal(size=(n_particles, self.dim))
        p /= np.linalg.norm(p, axis=-1, keepdims=True)
        u = np.random.random(size=(n_particles, 1)) ** (1.0 / self.dim)
        p = p * u * radius + np.array(init_pos)[:self.dim]
        self.add_object(p, color, init_rot=init_rot)

    def get(self):
        assert len(self.objects) > 0, "please add at least one shape into the scene"
        return np.concatenate(self.objects), np.concatenate(self.colors)
