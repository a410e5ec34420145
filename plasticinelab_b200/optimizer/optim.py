"""First-order optimisers over the action sequence (numpy, float64).

Behavioural contract = `plb/optimizer/optim.py:5-78`: `Optimizer(parameters, cfg=None, **overrides)` keeps a reference to
the caller's array, `step(grads)` writes the clipped update INTO that array and returns a copy; Adam uses bias-corrected
moments with t starting at 1; Momentum is an exponential average with weight (1 - momentum) on the new gradient.
The update rules live in small stateless functions so they can be unit-tested on their own.
"""
from __future__ import annotations

import numpy as np

from ..config import CfgNode, make_cls_config


def adam_update(theta, grad, m, v, t, lr, b1, b2, eps):
    """One Adam step at (1-based) time t; returns (new_theta, new_m, new_v)."""
    m = b1 * m + (1.0 - b1) * grad
    v = b2 * v + (1.0 - b2) * np.square(grad)
    m_hat = m / (1.0 - b1 ** t)
    v_hat = v / (1.0 - b2 ** t)
    return theta - lr * m_hat / (np.sqrt(v_hat) + eps), m, v


def momentum_update(theta, grad, buf, lr, momentum):
    """Heavy-ball with an averaged buffer; returns (new_theta, new_buf)."""
    buf = momentum * buf + (1.0 - momentum) * grad
    return theta - lr * buf, buf


class Optimizer:
    """Base class: bounds handling + config plumbing; subclasses provide `_propose(grads)`."""
    _defaults = dict(lr=0.1, bounds=(-1.0, 1.0), type='')

    def __init__(self, parameters: np.ndarray, cfg=None, **kwargs):
        self.cfg = make_cls_config(self, cfg, **kwargs)
        self.lr, self.bounds = self.cfg.lr, self.cfg.bounds
        self.parameters = parameters
        self.initialize()

    def initialize(self):
        pass

    def step(self, grads):
        if grads.shape != self.parameters.shape:
            raise AssertionError(f"gradient shape {grads.shape} != parameter shape {self.parameters.shape}")
        lo, hi = self.bounds
        np.clip(self._propose(grads), lo, hi, out=self.parameters)
        return np.array(self.parameters, copy=True)

    @classmethod
    def default_config(cls):
        merged = {}
        for klass in reversed(cls.__mro__):
            merged.update(getattr(klass, '_defaults', {}))
        return CfgNode(merged)


class Momentum(Optimizer):
    _defaults = dict(momentum=0.9)

    def initialize(self):
        self.momentum = self.cfg.momentum
        self.momentum_buffer = np.zeros(self.parameters.shape, dtype=np.float64)

    def _propose(self, grads):
        new, self.momentum_buffer = momentum_update(self.parameters, grads, self.momentum_buffer, self.lr, self.momentum)
        return new


class Adam(Optimizer):
    _defaults = dict(beta_1=0.9, beta_2=0.999, epsilon=1e-8)

    def initialize(self):
        self.iter = 0
        self.momentum_buffer = np.zeros(self.parameters.shape, dtype=np.float64)
        self.v_buffer = np.zeros(self.parameters.shape, dtype=np.float64)

    def _propose(self, grads):
        self.iter += 1
        c = self.cfg
        new, self.momentum_buffer, self.v_buffer = adam_update(self.parameters, grads.reshape(self.parameters.shape), self.momentum_buffer,
                                                                self.v_buffer, self.iter, self.lr, c.beta_1, c.beta_2, c.epsilon)
        return new
