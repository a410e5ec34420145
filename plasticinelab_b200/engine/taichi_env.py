"""`TaichiEnv`: owns one scene -- primitives, particles, simulator, loss -- and its lifecycle.

Public surface = `plb/engine/taichi_env.py:9-106` (constructor arguments, `initialize`, `set_copy`, `step`, `compute_loss`,
`get_state`, `set_state`, `render`, attributes `simulator`, `primitives`, `loss`, `n_particles`, `init_particles`).
What is different underneath: no Taichi runtime -- a CUDA engine handle is created through the C ABI (no CPU fallback);
`dtype` picks the float32 production kernels or the float64 parity kernels (the reference only has float64,
`mpm_simulator.py:8`); `nn=True` attaches the state-feedback MLP policy (`engine/nn/mlp.py`); `render` is out of scope
(SURVEY.md 2a #7).
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


def _resolve_dtype(requested, sim_cfg):
    return requested or os.environ.get("PLB_DTYPE") or sim_cfg.get("dtype", "float64")


class TaichiEnv:
    def __init__(self, cfg, nn=False, loss=True, dtype=None, device=0, max_prim_frames=None, particle_index=None):
        sim_cfg = cfg.SIMULATOR
        self.cfg = cfg.ENV
        self._is_copy = True
        self.renderer = None

        # scene content: manipulators + sampled particles (optionally one rank's share, engine/sharded.py)
        self.primitives = Primitives(cfg.PRIMITIVES, max_timesteps=sim_cfg.max_steps)
        self.shapes = Shapes(cfg.SHAPES)
        pts, colors = self.shapes.get()
        n_global = len(pts)
        if particle_index is not None:
            pts, colors = np.ascontiguousarray(pts[particle_index]), colors[particle_index]
        self.init_particles, self.particle_colors = pts, colors
        self.n_particles = sim_cfg.n_particles = len(pts)

        # native engine + host mirrors bound to it
        conf = _capi.make_config(dict(sim_cfg), self.n_particles, len(self.primitives), dtype=_resolve_dtype(dtype, sim_cfg),
                                 max_frames=sim_cfg.max_steps, max_prim_frames=max_prim_frames, device=device)
        self.engine = _capi.Engine(conf, [p.desc for p in self.primitives])
        self.primitives.bind(self.engine)
        self.simulator = MPMSimulator(sim_cfg, self.primitives, self.engine)
        self.simulator._env = self
        self.simulator.n_particles_global = n_global      # (a slab rank holds a share; target densities are scaled to the whole body)
        if nn:
            from .nn.mlp import MLP
            self.nn = MLP(self.simulator, self.primitives, (256, 256))     # taichi_env.py:35-36
        self.loss = Loss(cfg.ENV.loss, self.simulator) if loss else None

    # ------------------------------------------------------------------ lifecycle
    def initialize(self):
        for part in (self.primitives, self.simulator, self.loss):
            if part is not None:
                part.initialize()
        self.simulator.reset(self.init_particles)
        if self.loss is not None:
            self.loss.clear()

    def set_copy(self, is_copy: bool):
        """True: RL mode (every step runs frames 0..S and copies back); False: trajectory mode (gradients)."""
        self._is_copy = is_copy

    def tape(self, loss=None):
        return Tape(self)

    def render(self, mode="human", **kwargs):
        raise NotImplementedError("rendering (plb/engine/renderer) is out of scope for this engine (SURVEY.md 2a #7)")

    # ------------------------------------------------------------------ stepping
    def step(self, action=None):
        sim = self.simulator
        first_frame = 0 if self._is_copy else sim.cur
        sim.step(is_copy=self._is_copy, action=None if action is None else np.array(action))
        tape = active_tape()
        if tape is not None and not self._is_copy:
            tape.record_step(first_frame, sim.substeps)

    def compute_loss(self):
        if self.loss is None:
            raise AssertionError("this env was built without a loss")
        if self._is_copy:
            self.loss.clear()
            return self.loss.compute_loss(0)
        frame = self.simulator.cur
        tape = active_tape()
        if tape is not None:
            tape.record_loss(frame)
        return self.loss.compute_loss(frame)

    # ------------------------------------------------------------------ state i/o (frame 0 only, like the reference)
    def get_state(self):
        if self.simulator.cur != 0:
            raise AssertionError("get_state is only valid at frame 0")
        return dict(state=self.simulator.get_state(0), softness=self.primitives.get_softness(), is_copy=self._is_copy)

    def set_state(self, state, softness, is_copy):
        self.simulator.cur = 0
        self.simulator.set_state(0, state)
        self.primitives.set_softness(softness)
        self._is_copy = is_copy
        if self.loss is not None:
            self.loss.reset()
            self.loss.clear()
