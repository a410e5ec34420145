"""Env-step gradient checkpointing: fwd+bwd episodes whose full trajectory does not fit in HBM.

The reference keeps every substep frame (`[max_steps, N]` fields, `plb/engine/mpm_simulator.py:33-38`) and demonstrates
checkpointing only in a notebook (`plb/optimizer/long_term_gradient.ipynb` cell 4: store the state every T env steps,
re-simulate a segment, call `substep_grad` manually, compare with the full-tape gradient).  Here it is a mode of the
engine's frame storage: the particle frames of ONE env step live in a working window (slots 0..S), the state at every
env-step boundary is kept in a checkpoint slot, and the backward pass re-simulates one env step at a time before running
its adjoint.  Primitive frames keep their global indices (they are tiny), which is why the C ABI separates `slot` from `pf`.

Memory: (S + 1) + (H + 1) frames instead of H*S + 1  (1M particles, S=39, H=50, f32: 8.7 GB instead of 187 GB).
Cost: one extra forward pass.  Result: identical to the un-checkpointed gradient up to float summation order.
"""
from __future__ import annotations

import numpy as np

from .. import _capi


class CheckpointedEpisode:
    """Drives a `TaichiEnv` (built with max_steps >= S + H + 3) through a checkpointed fwd+bwd episode."""

    def __init__(self, env, horizon):
        self.env = env
        self.eng = env.engine
        self.S = env.simulator.substeps
        self.H = int(horizon)
        need = self.S + 1 + self.H + 1
        if env.engine.config.max_frames < need:
            raise ValueError(f"checkpointed episode needs max_steps >= {need} (window {self.S + 1} + {self.H + 1} checkpoints)")
        if env.engine.config.max_prim_frames < self.H * self.S + 1:
            raise ValueError("primitive trajectory buffer too short: pass max_prim_frames >= horizon * substeps + 1")
        self.ckpt0 = self.S + 1                      # first checkpoint slot

    def ckpt(self, k):
        return self.ckpt0 + k

    def forward_backward(self, actions, softness=666.0, sync_losses=False):
        """Frame 0 must hold the start state (after `env.set_state` / `reset`).  Returns (summed loss, grad [H, A])."""
        env, eng, S, H = self.env, self.eng, self.S, self.H
        actions = np.asarray(actions, dtype=np.float64)
        assert len(actions) == H
        A = env.primitives.action_dim
        env._is_copy = False
        env.simulator.cur = 0
        for p in env.primitives:
            p.set_state(0, p.get_state(0))
        env.primitives.set_softness(softness)
        eng.call("plb_zero_grads")
        eng.call("plb_copy_frame", 0, self.ckpt(0))
        infos = []
        for k in range(H):                            # forward: window slots 0..S, checkpoint the boundary state
            if k > 0:
                eng.call("plb_copy_frame", self.ckpt(k), 0)
            eng.call("plb_set_action", k, S, _capi.dptr(np.ascontiguousarray(actions[k])), A)
            eng.call("plb_kinematics", k * S, S)
            eng.call("plb_step_fwd", 0, k * S, S)
            eng.call("plb_copy_frame", S, self.ckpt(k + 1))
            out = np.zeros(8) if sync_losses else None
            eng.call("plb_loss_fwd", self.ckpt(k + 1), (k + 1) * S, _capi.dptr(out))
            infos.append(out)
        for k in reversed(range(H)):                  # backward: seed with the loss adjoint, re-simulate, run the adjoint
            eng.call("plb_loss_bwd", self.ckpt(k + 1), (k + 1) * S)
            eng.call("plb_copy_frame", self.ckpt(k), 0)
            eng.call("plb_step_fwd", 0, k * S, S)
            eng.call("plb_step_bwd", 0, k * S, S)
        grad = env.primitives.get_grad(H, S)
        loss = env.loss.loss[None]
        return loss, grad
