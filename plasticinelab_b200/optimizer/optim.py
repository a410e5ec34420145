"""Numpy optimisers over the action sequence; behaviour of `plb/optimizer/optim.py:5-78`."""
import numpy as np

from ..config import CfgNode, make_cls_config


class Optimizer:
    def __init__(self, parameters: np.ndarray, cfg=None, **kwargs):
        self.cfg = make_cls_config(self, cfg, **kwargs)
        self.lr = self.cfg.lr
        self.bounds = self.cfg.bounds
        self.parameters = parameters
        self.initialize()

    def step(self, grads):
        assert grads.shape == self.parameters.shape
        self.parameters[:] = self._step(grads).clip(*self.bounds)
        return self.parameters.copy()

    @classmethod
    def default_config(cls):
        return CfgNode(dict(lr=0.1, bounds=(-1.0, 1.0), type=''))


class Momentum(Optimizer):
    def initialize(self):
        self.momentum_buffer = np.zeros_like(self.parameters).astype(np.float64)
        self.momentum = self.cfg.momentum

    def _step(self, grads):
        grads = self.momentum_buffer * self.momentum + grads * (1 - self.momentum)
        self.momentum_buffer[:] = grads
        return self.parameters[:] - self.lr * grads

    @classmethod
    def default_config(cls):
        cfg = Optimizer.default_config()
        cfg.momentum = 0.9
        return cfg


class Adam(Optimizer):
    def initialize(self):
        self.momentum_buffer = np.zeros_like(self.parameters).astype(np.float64)
        self.v_buffer = np.zeros_like(self.momentum_buffer).astype(np.float64)
        self.iter = 0

    def _step(self, grads):
        gd = grads.reshape(*self.parameters.shape)
        b1, b2, eps = self.cfg.beta_1, self.cfg.beta_2, self.cfg.epsilon
        m_t = b1 * self.momentum_buffer + (1 - b1) * gd
        v_t = b2 * self.v_buffer + (1 - b2) * (gd * gd)
        self.momentum_buffer[:] = m_t
        self.v_buffer[:] = v_t
        m_cap = m_t / (1 - (b1 ** (self.iter + 1)))
        v_cap = v_t / (1 - (b2 ** (self.iter + 1)))
        self.iter += 1
        return self.parameters - (self.lr * m_cap) / (np.sqrt(v_cap) + eps)

    @classmethod
    def default_config(cls):
        cfg = Optimizer.default_config()
        cfg.beta_1, cfg.beta_2, cfg.epsilon = 0.9, 0.999, 1e-8
        return cfg
