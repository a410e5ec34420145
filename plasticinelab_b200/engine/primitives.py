"""Host mirror of the rigid manipulators (`plb/engine/primitive/primitives.py:262-320`, `primive_base.py`).

The kinematic state, the action buffers and their gradients live inside the native engine (C++ float64 on the
host, mirrored to the GPU); these classes expose the reference's names on top of the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _capi
from ..config import CfgNode


class Primitive:
    state_dim = 7

    def __init__(self, cfg: dict, index: int):
        self.cfg = CfgNode(dict(cfg))
        self.index = index
        self.shape = cfg["shape"]
        action = cfg.get("action") or {}
        self.action_dim = int(action.get("dim", 0) or 0)
        if self.shape == "Chopsticks":
            self.state_dim = 8
        self.desc = _capi.primitive_desc(dict(cfg))
        self._engine = None

    @property
    def init_state(self):
        return tuple(self.desc.init_state)[: self.state_dim]

    def get_state(self, f):
        out = np.zeros(8, dtype=np.float64)
        self._engine.call("plb_get_primitive_state", int(f), self.index, _capi.dptr(out))
        return out[: self.state_dim].copy()

    def set_state(self, f, state):
        ss = np.zeros(8, dtype=np.float64)
        self._engine.call("plb_get_primitive_state", int(f), self.index, _capi.dptr(ss))
        state = np.asarray(state, dtype=np.float64).reshape(-1)
        if self.state_dim == 8:
            assert len(state) == 8
        ss[: len(state)] = state
        self._engine.call("plb_set_primitive_state", int(f), self.index, _capi.dptr(ss))


class Primitives:
    def __init__(self, cfgs, max_timesteps=1024):
        self.primitives = [Primitive(dict(c), i) for i, c in enumerate(cfgs)]
        self.action_dims = [0]
        for p in self.primitives:
            self.action_dims.append(self.action_dims[-1] + p.action_dim)
        self.n = len(self.primitives)
        self.max_timesteps = max_timesteps
        self._engine = None
        self._softness = 0.0

    def bind(self, engine):
        self._engine = engine
        for p in self.primitives:
            p._engine = engine

    @property
    def action_dim(self):
        return self.action_dims[-1]

    @property
    def state_dim(self):
        return sum(p.state_dim for p in self.primitives)

    def set_action(self, s, n_substeps, action):
        action = np.ascontiguousarray(np.asarray(action, dtype=np.float64).reshape(-1))
        assert len(action) == self.action_dims[-1]
        self._engine.call("plb_set_action", int(s), int(n_substeps), _capi.dptr(action), len(action))

    def get_grad(self, n, n_substeps=None):
        """(n, sum action_dim) float64, like Primitives.get_grad (primitives.py:295-301)."""
        S = int(n_substeps if n_substeps is not None else self._engine.config.substeps)
        out = np.zeros((int(n), max(self.action_dim, 1)), dtype=np.float64)
        self._engine.call("plb_get_action_grad", int(n), S, _capi.dptr(out))
        return out[:, : self.action_dim]

    def set_softness(self, softness=666.0):
        self._softness = float(softness)
        self._engine.call("plb_set_softness", C.c_double(self._softness))

    def get_softness(self):
        return self._softness

    def __getitem__(self, item):
        if isinstance(item, tuple):
            item = item[0]
        return self.primitives[item]

    def __len__(self):
        return len(self.primitives)

    def __iter__(self):
        return iter(self.primitives)

    def initialize(self):
        for p in self.primitives:
            p.set_state(0, p.init_state)
