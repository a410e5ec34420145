"""`PlasticineEnv`: the gym-style task wrapper (`plb/envs/env.py:12-86`) on top of the CUDA engine.

Observation = [x_i, v_i] of every (N // n_observed)-th particle, flattened, followed by the primitive states; reward and
the info dict come from `Loss.compute_loss`; episodes never terminate by themselves (`done` is always False, the
TimeLimit wrapper of `envs.make` ends them after 50 steps)."""
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
        full_cfg = self.load_varaints(cfg_path, version)
        if cfg_overrides is not None:
            cfg_overrides(full_cfg)
        self.cfg = full_cfg.ENV
        self._n_observed_particles = self.cfg.n_observed_particles
        self.taichi_env = sim_env = TaichiEnv(full_cfg, nn, dtype=dtype, device=device)
        sim_env.initialize()
        sim_env.set_copy(True)                       # RL mode: every step restarts from frame 0
        self._init_state = sim_env.get_state()
        first_obs = self.reset()
        self.observation_space = Box(-np.inf, np.inf, first_obs.shape)
        self.action_space = Box(-1, 1, (sim_env.primitives.action_dim,))

    # reference spelling (env.py:62); kept because callers use it
    load_varaints = staticmethod(load_variants)

    def _get_obs(self, t=0):
        sim = self.taichi_env.simulator
        stride = sim.n_particles // self._n_observed_particles
        particles = np.concatenate((sim.get_x(t)[::stride], sim.get_v(t)[::stride]), axis=-1)
        prims = [p.get_state(t) for p in self.taichi_env.primitives]
        return np.concatenate([particles.reshape(-1)] + [np.asarray(s).reshape(-1) for s in prims])

    def reset(self):
        self._recorded_actions = []
        self.taichi_env.set_state(**self._init_state)
        return self._get_obs()

    def step(self, action):
        self.begin_step(action)
        return self.finish_step(action)

    # the two halves of `step`: `begin_step` only enqueues work on this env's CUDA stream (the S substeps are one graph
    # launch), `finish_step` reads loss and observation back (device sync).  `VecPlasticineEnv` calls begin_step on every
    # env before the first finish_step so that the envs' kernels overlap on the GPU.
    def begin_step(self, action):
        self.taichi_env.step(action)

    def finish_step(self, action):
        info = self.taichi_env.compute_loss()
        self._recorded_actions.append(action)
        obs, reward = self._get_obs(), info["reward"]
        if np.isnan(reward) or np.isnan(obs).any():
            raise Exception("NaN..")          # the reference additionally pickles the action log (env.py:50-56)
        return obs, reward, False, info

    def render(self, mode="human"):
        return self.taichi_env.render(mode)
