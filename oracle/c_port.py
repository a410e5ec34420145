"""ctypes wrapper of oracle/mpm_oracle.c (plain-C/OpenMP float64 port of the substep and its adjoint; Sphere primitives).
TEST INFRASTRUCTURE / CPU BASELINE ONLY -- same rules as plb_oracle.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_build', 'liboracle_c.so')


class Params(C.Structure):
    _fields_ = [("n", C.c_int), ("n_grid", C.c_int), ("n_prim", C.c_int),
                ("dx", C.c_double), ("inv_dx", C.c_double), ("dt", C.c_double), ("p_vol", C.c_double), ("p_mass", C.c_double),
                ("mu", C.c_double), ("lam", C.c_double), ("yield_stress", C.c_double), ("ground_friction", C.c_double),
                ("grav_dv", C.c_double * 3), ("radius", C.c_double * 8), ("friction", C.c_double * 8), ("softness", C.c_double)]


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class CPort:
    def __init__(self, osim, n, softness):
        """osim: plb_oracle.OracleSim (only Sphere primitives)."""
        self.lib = C.CDLL(_SO)
        p = Params()
        p.n, p.n_grid, p.n_prim = n, osim.n_grid, len(osim.prims)
        p.dx, p.inv_dx, p.dt, p.p_vol, p.p_mass = osim.dx, osim.inv_dx, osim.dt, osim.p_vol, osim.p_mass
        p.mu, p.lam, p.yield_stress, p.ground_friction = osim.mu0, osim.lam0, osim.yield0, osim.ground_friction
        g = osim.gravity.numpy()
        p.grav_dv = (C.c_double * 3)(*(osim.dt * g * 30))
        for k, pr in enumerate(osim.prims):
            assert pr.shape == 'Sphere', 'the C port covers Sphere primitives only'
            p.radius[k] = pr.radius
            p.friction[k] = pr.friction
        p.softness = float(softness)
        self.p = p
        self.n = n
        self.G = osim.n_grid ** 3
        self.grid = np.zeros(self.G * 16)
        self.threads = self.lib.oc_max_threads()

    @staticmethod
    def poses(states):
        out = np.zeros((max(len(states), 1), 8))
        for k, s in enumerate(states):
            s = np.asarray(s, dtype=np.float64)
            out[k, :len(s)] = s
        return out

    def substep_fwd(self, state, pose0, pose1):
        x, v, Cm, F = [np.ascontiguousarray(a, dtype=np.float64) for a in state]
        xo, vo, Co, Fo = np.zeros_like(x), np.zeros_like(v), np.zeros_like(Cm), np.zeros_like(F)
        self.lib.oc_substep_fwd(C.byref(self.p), _p(pose0), _p(pose1), _p(x), _p(v), _p(Cm), _p(F), _p(xo), _p(vo), _p(Co), _p(Fo), _p(self.grid))
        return xo, vo, Co, Fo

    def substep_bwd(self, state, pose0, pose1, adj_next):
        x, v, Cm, F = [np.ascontiguousarray(a, dtype=np.float64) for a in state]
        gxn, gvn, gCn, gFn = [np.ascontiguousarray(a, dtype=np.float64) for a in adj_next]
        gx, gv, gC, gF = np.zeros_like(x), np.zeros_like(v), np.zeros_like(Cm), np.zeros_like(F)
        g0, g1 = np.zeros((max(self.p.n_prim, 1), 8)), np.zeros((max(self.p.n_prim, 1), 8))
        self.lib.oc_substep_bwd(C.byref(self.p), _p(pose0), _p(pose1), _p(x), _p(v), _p(Cm), _p(F), _p(gxn), _p(gvn), _p(gCn), _p(gFn),
                                _p(gx), _p(gv), _p(gC), _p(gF), _p(g0), _p(g1), _p(self.grid))
        return (gx, gv, gC, gF), g0, g1
