"""`PlasticineEnv`: the gym surface of the reference (`plb/envs/env.py:12-86`) on top of the CUDA engine."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gym is absent in this image
    import gym
    from gym.spaces import Box
    _EnvBase = gym.Env
except Exception:  # noqa: BLE001
    from .gym_shim import Box, Env as _EnvBase

from ..engine.taichi_env import TaichiEnv
from .scene import load_variants


class PlasticineEnv(_EnvBase):
    def __init__(self, cfg_path, version, nn=False, dtype=None, device=0, cfg_overrides=None):
        self.cfg_path = cfg_path
        cfg = self.load_varaints(cfg_path, version)
        if cfg_overrides:
            cfg_overrides(cfg)
        self.taichi_env = TaichiEnv(cfg, nn, dtype=dtype, device=device)
        self.taichi_env.initialize()
        self.cfg = cfg.ENV
        self.taichi_env.set_copy(True)
        self._init_state = self.taichi_env.get_state()
        self._n_observed_particles = self.cfg.n_observed_particles
        obs = self.reset()
        self.observation_space = Box(-np.inf, np.inf, obs.shape)
        self.action_space = Box(-1, 1, (self.taichi_env.primitives.action_dim,))

    def reset(self):
        self.taichi_env.set_state(**self._init_state)
        self._recorded_actions = []
        return self._get_obs()

    def _get_obs(self, t=0):
        x = self.taichi_env.simulator.get_x(t)
        v = self.taichi_env.simulator.get_v(t)
        outs = [p.get_state(t) for p in self.taichi_env.primitives]
        s = np.concatenate(outs) if outs else np.zeros(0)
        step_size = len(x) // self._n_observed_particles
        return np.concatenate((np.concatenate((x[::step_size], v[::step_size]), axis=-1).reshape(-1), s.reshape(-1)))

    def step(self, action):
        self.taichi_env.step(action)
        loss_info = self.taichi_env.compute_loss()
        self._recorded_actions.append(action)
        obs = self._get_obs()
        r = loss_info["reward"]
        if np.isnan(obs).any() or np.isnan(r):
            raise Exception("NaN..")          # the reference also pickles the action log (env.py:50-56)
        return obs, r, False, loss_info

    def render(self, mode="human"):
        return self.taichi_env.render(mode)

    @classmethod
    def load_varaints(cls, cfg_path, version):      # (sic) reference spelling, env.py:62
        return load_variants(cfg_path, version)
