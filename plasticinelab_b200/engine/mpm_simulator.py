"""Host mirror of `MPMSimulator` (`plb/engine/mpm_simulator.py`): same attribute and method names, every
compute call forwarded to the CUDA engine through the C ABI (`include/plb_b200.h`).
"""
from __future__ import annotations

import numpy as np

from .. import _capi


class MPMSimulator:
    def __init__(self, cfg, primitives, engine):
        k = _capi.sim_constants(dict(cfg))
        self.cfg = cfg
        self.dim = 3
        self.dtype = "float64" if engine.config.dtype == _capi.PLB_F64 else "float32"
        self.n_particles = engine.config.n_particles
        self.n_grid = k["n_grid"]
        self.dx, self.inv_dx, self.dt = k["dx"], k["inv_dx"], k["dt"]
        self.p_vol, self.p_rho, self.p_mass = k["p_vol"], k["p_rho"], k["p_mass"]
        self.substeps = k["substeps"]
        self.max_steps = engine.config.max_frames
        self.res = (self.n_grid,) * 3
        self.ground_friction = k["ground_friction"]
        self.primitives = primitives
        self.n_primitive = len(primitives)
        self.engine = engine
        self.cur = 0

    def initialize(self):
        pass  # uniform material constants already live in the engine (mpm_simulator.py:53-57)

    def set_materials(self, mu=None, lam=None, yield_stress=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (mu, lam, yield_stress)]
        self.engine.call("plb_set_materials", *[_capi.dptr(a) for a in arrs])

    # ---- one substep and its adjoint (mpm_simulator.py:245-278)
    def substep(self, s):
        self.engine.call("plb_kinematics", int(s), 1)
        self.engine.call("plb_substep_fwd", int(s), int(s) + 1, int(s))

    def substep_grad(self, s):
        self.engine.call("plb_substep_bwd", int(s), int(s))

    # ---- io (mpm_simulator.py:282-363)
    def get_state(self, f):
        n = self.n_particles
        x, v = np.zeros((n, 3)), np.zeros((n, 3))
        F, C = np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
        self.engine.call("plb_get_frame", int(f), _capi.dptr(x), _capi.dptr(v), _capi.dptr(F), _capi.dptr(C))
        out = [x, v, F, C]
        for p in self.primitives:
            out.append(p.get_state(f))
        return out

    def set_state(self, f, state):
        x, v, F, C = [np.ascontiguousarray(a, dtype=np.float64) for a in state[:4]]
        self.engine.call("plb_set_frame", int(f), _capi.dptr(x), _capi.dptr(v), _capi.dptr(F), _capi.dptr(C))
        for s, p in zip(state[4:], self.primitives):
            p.set_state(f, s)
        if f == 0:
            self.engine.call("plb_sort_particles", 0)     # new episode state: restore spatial order

    def reset(self, x):
        n = self.n_particles
        x = np.ascontiguousarray(x, dtype=np.float64)
        F = np.ascontiguousarray(np.broadcast_to(np.eye(3), (n, 3, 3)))
        self.engine.call("plb_set_frame", 0, _capi.dptr(x), _capi.dptr(np.zeros((n, 3))), _capi.dptr(F),
                         _capi.dptr(np.zeros((n, 3, 3))))
        self.engine.call("plb_sort_particles", 0)
        self.cur = 0

    def get_x(self, f):
        x = np.zeros((self.n_particles, 3))
        self.engine.call("plb_get_frame", int(f), _capi.dptr(x), None, None, None)
        return x

    def get_v(self, f):
        v = np.zeros((self.n_particles, 3))
        self.engine.call("plb_get_frame", int(f), None, _capi.dptr(v), None, None)
        return v

    def copyframe(self, source, target):
        self.engine.call("plb_copy_frame", int(source), int(target))
        self.engine.call("plb_copy_primitive_frame", int(source), int(target))

    # ---- env step (mpm_simulator.py:365-376)
    def step(self, is_copy, action=None):
        start = 0 if is_copy else self.cur
        self.cur = start + self.substeps
        if action is not None:
            self.primitives.set_action(start // self.substeps, self.substeps, action)
        self.engine.call("plb_kinematics", int(start), int(self.substeps))
        self.engine.call("plb_step_fwd", int(start), int(start), int(self.substeps))
        if is_copy:
            self.copyframe(self.cur, 0)
            self.cur = 0
