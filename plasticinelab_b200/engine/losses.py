"""Host mirror of `Loss` (`plb/engine/losses/loss.py`): names and bookkeeping of the reference, the grid/particle
reductions and their adjoint run in the CUDA engine.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _capi


class _LossScalar:
    """Stands in for the 0-d Taichi field `Loss.loss`: `loss[None]` reads the accumulated loss (solver.py:43)."""

    def __init__(self, owner):
        self._owner = owner

    def __getitem__(self, key):
        out = np.zeros(1)
        self._owner.engine.call("plb_get_loss", _capi.dptr(out))
        return float(out[0])


class Loss:
    def __init__(self, cfg, sim):
        self.cfg = cfg
        self.sim = sim
        self.engine = sim.engine
        self.res = sim.res
        self.n_grid = sim.n_grid
        self.dx = sim.dx
        self.n_particles = getattr(sim, 'n_particles_global', sim.n_particles)
        self.loss = _LossScalar(self)
        self.soft_contact_loss = False
        self.contact_grad_all = True     # reference autodiff of ti.atomic_min (see oracle docstring)
        self._weights = (10.0, 10.0, 1.0)
        self._target_iou = 0.0
        self._iou = 0.0
        self.target_density = None
        self._start_loss = 0.0
        self._init_iou = 0.0
        self._last_loss = 0.0

    # ---- targets (loss.py:46-66)
    def load_target_density(self, path=None, grids=None, target_sdf=None):
        from ..envs.scene import load_target, resample_target      # (late import: envs imports the engine)
        if path is not None and len(path) > 0:
            grids = load_target(path)
        if grids is None:
            return
        grids = np.asarray(grids, dtype=np.float64)
        if grids.shape[0] != self.n_grid:
            grids = resample_target(grids, self.n_grid, self.n_particles * self.sim.p_mass)
        self.target_density = np.ascontiguousarray(grids)
        sdf = None if target_sdf is None else np.ascontiguousarray(target_sdf, dtype=np.float64)
        self.engine.call("plb_set_target", _capi.dptr(self.target_density), _capi.dptr(sdf))
        # iou of the target with itself (loss.py:55-57)
        t = self.target_density
        ma = t.max()
        I = (t * t).sum() / ma / ma
        U = 2 * t.sum() / ma
        self._target_iou = float(I / (U - I))

    def target_sdf(self):
        out = np.zeros(self.res)
        self.engine.call("plb_get_target_sdf", _capi.dptr(out))
        return out

    def initialize(self):
        w = self.cfg.weight
        self.set_weights(w.sdf, w.density, w.contact, self.cfg.soft_contact)
        self.load_target_density(self.cfg.target_path)

    def set_weights(self, sdf, density, contact, is_soft_contact):
        self._weights = (float(sdf), float(density), float(contact))
        self.soft_contact_loss = bool(is_soft_contact)
        self.engine.call("plb_set_loss_weights", C.c_double(sdf), C.c_double(density), C.c_double(contact),
                         int(self.soft_contact_loss), int(self.contact_grad_all))

    # ---- evaluation (loss.py:186-208,269-298)
    def compute_loss_kernel(self, f):
        self.engine.call("plb_loss_fwd", int(f), int(f), None)

    def compute_loss_kernel_grad(self, f):
        self.engine.call("plb_loss_bwd", int(f), int(f))

    def _extract_loss(self, f):
        out = np.zeros(8)
        self.engine.call("plb_loss_fwd", int(f), int(f), _capi.dptr(out))
        self._iou = float(out[4])
        return {"loss": float(out[0]), "contact_loss": float(out[1]), "density_loss": float(out[2]),
                "sdf_loss": float(out[3]), "iou": self._iou, "target_iou": self._target_iou}

    def reset(self):
        self.clear_loss()
        info = self._extract_loss(0)
        self._start_loss = info["loss"]
        self._init_iou = info["iou"]
        self._last_loss = 0

    def compute_loss(self, f):
        info = self._extract_loss(f)
        r = self._start_loss - (info["loss"] - self._last_loss)
        cur_step_loss = info["loss"] - self._last_loss
        self._last_loss = info["loss"]
        denom = info["target_iou"] - self._init_iou
        inc = (info["iou"] - self._init_iou) / denom if denom != 0 else 0.0
        info["reward"] = r
        info["incremental_iou"] = max(min(inc, 1), 0)
        info["loss"] = cur_step_loss
        return info

    def clear_loss(self):
        self.engine.call("plb_clear_loss")

    def clear(self):
        self.clear_loss()
        self._last_loss = 0
