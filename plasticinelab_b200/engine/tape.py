"""Reverse-mode tape over env steps: the replacement for `ti.Tape(loss=env.loss.loss)` (`plb/optimizer/solver.py:36`).

On entry every adjoint (particles, primitive poses, action gradients) and the accumulated loss are cleared and
d(loss) = 1 is implied; `TaichiEnv.step` / `compute_loss` record themselves; on exit the records are replayed in
reverse through the engine's adjoint kernels (loss adjoint, then S substep adjoints per env step).
"""
from __future__ import annotations

_ACTIVE = None


def active_tape():
    return _ACTIVE


class Tape:
    def __init__(self, env=None, loss=None):
        if env is None and loss is not None:
            env = loss._owner.sim._env          # ti.Tape(loss=env.loss.loss) spelling
        self.env = env
        self.records = []

    def __enter__(self):
        global _ACTIVE
        assert _ACTIVE is None, "nested tapes are not supported"
        self.env.simulator.engine.call("plb_zero_grads")
        self.records = []
        _ACTIVE = self
        return self

    def record_step(self, start, n):
        self.records.append(("step", int(start), int(n)))

    def record_loss(self, f):
        self.records.append(("loss", int(f)))

    def record_policy(self, policy, step):
        """`policy.set_action(step, S)` ran (engine/nn/mlp.py): its backward runs when the sweep is back at frame step*S."""
        if not any(r[0] == "policy" and r[1] is policy for r in self.records):
            policy.zero_grad_for_tape()
        self.records.append(("policy", policy, int(step)))

    def __exit__(self, exc_type, exc, tb):
        global _ACTIVE
        _ACTIVE = None
        if exc_type is not None:
            return False
        eng = self.env.simulator.engine
        for rec in reversed(self.records):
            if rec[0] == "loss":
                eng.call("plb_loss_bwd", rec[1], rec[1])
            elif rec[0] == "policy":
                rec[1].backward(rec[2])
            else:
                eng.call("plb_step_bwd", rec[1], rec[1], rec[2])
        return False
