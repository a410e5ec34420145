"""`TaichiEnv`: scene assembly + lifecycle with the reference's surface (`plb/engine/taichi_env.py:9-106`).

Differences by design: there is no Taichi runtime -- a CUDA engine handle is created instead (no CPU fallback);
`render` is out of scope and raises; `dtype` selects the float32 production kernels or the
float64 parity kernels (the reference only supports float64, `mpm_simulator.py:8`).
"""
from __future__ import annotations

import os

import numpy as np

from .. import _capi
from .losses import Loss
from .mpm_simulator import MPMSimulator
from .primitives import Primitives
from .shapes import Shapes
from .tape import Tape, active_tape


class TaichiEnv:
    def __init__(self, cfg, nn=False, loss=True, dtype=None, device=0, max_prim_frames=None, particle_index=None):
        if nn:
            raise NotImplementedError("the Taichi MLP policy (plb/engine/nn/mlp.py) is out of scope (SURVEY.md 8f #3)")
        self.cfg = cfg.ENV
        self.primitives = Primitives(cfg.PRIMITIVES, max_timesteps=cfg.SIMULATOR.max_steps)
        self.shapes = Shapes(cfg.SHAPES)
        self.init_particles, self.particle_colors = self.shapes.get()
        if particle_index is not None:      # one rank's share of the particles (slab decomposition, engine/sharded.py)
            self.init_particles = np.ascontiguousarray(self.init_particles[particle_index])
            self.particle_colors = self.particle_colors[particle_index]
        self.n_particles = cfg.SIMULATOR.n_particles = len(self.init_particles)
        dtype = dtype or os.environ.get("PLB_DTYPE") or cfg.SIMULATOR.get("dtype", "float64")
        conf = _capi.make_config(dict(cfg.SIMULATOR), self.n_particles, len(self.primitives), dtype=dtype,
                                 max_frames=cfg.SIMULATOR.max_steps, max_prim_frames=max_prim_frames, device=device)
        self.engine = _capi.Engine(conf, [p.desc for p in self.primitives])
        self.primitives.bind(self.engine)
        self.simulator = MPMSimulator(cfg.SIMULATOR, self.primitives, self.engine)
        self.simulator._env = self
        self.renderer = None
        self.loss = Loss(cfg.ENV.loss, self.simulator) if loss else None
        self._is_copy = True

    def set_copy(self, is_copy: bool):
        self._is_copy = is_copy

    def initialize(self):
        self.primitives.initialize()
        self.simulator.initialize()
        if self.loss:
            self.loss.initialize()
        self.simulator.reset(self.init_particles)
        if self.loss:
            self.loss.clear()

    def render(self, mode="human", **kwargs):
        raise NotImplementedError("rendering (plb/engine/renderer) is out of scope for this engine (SURVEY.md 2a #7)")

    def tape(self, loss=None):
        return Tape(self)

    def step(self, action=None):
        if action is not None:
            action = np.array(action)
        start = 0 if self._is_copy else self.simulator.cur
        self.simulator.step(is_copy=self._is_copy, action=action)
        t = active_tape()
        if t is not None and not self._is_copy:
            t.record_step(start, self.simulator.substeps)

    def compute_loss(self):
        assert self.loss is not None
        if self._is_copy:
            self.loss.clear()
            return self.loss.compute_loss(0)
        t = active_tape()
        if t is not None:
            t.record_loss(self.simulator.cur)
        return self.loss.compute_loss(self.simulator.cur)

    def get_state(self):
        assert self.simulator.cur == 0
        return {"state": self.simulator.get_state(0), "softness": self.primitives.get_softness(), "is_copy": self._is_copy}

    def set_state(self, state, softness, is_copy):
        self.simulator.cur = 0
        self.simulator.set_state(0, state)
        self.primitives.set_softness(softness)
        self._is_copy = is_copy
        if self.loss:
            self.loss.reset()
            self.loss.clear()
