"""`VecPlasticineEnv`: K independent task envs on one GPU, stepped together (first half).

The reference's RL baselines step ONE env at a time (`plb/algorithms/ppo/ppo/envs.py:80-91` ends up with a 1-env
DummyVecEnv), and a 10k-particle env keeps a B200 a few per cent busy: its kernels are latency-bound on a handful of SMs.
Every `PlasticineEnv` here owns its engine handle and CUDA stream, and an env step is asynchronous until the loss is read
back, so stepping K envs as "enqueue all, then read all" lets their substep graphs run side by side.  No new device code:
the pool only re-orders calls the single env already makes.

API = the usual vector-env one: `reset() -> obs[K, D]`, `step(actions[K, A]) -> (obs[K, D], reward[K], done[K], infos)`;
envs that hit the 50-step limit are reset automatically and return the first observation of the new episode (the
terminal observation is kept in `info['terminal_observation']`).
"""
from __future__ import annotations

import numpy as np

from . import ENVS, MAX_EPISODE_STEPS
from .env import PlasticineEnv


class VecPlasticineEnv:
    def __init__(self, env_name, n_envs, sdf_loss=10, density_loss=10, contact_loss=1, soft_contact_loss=False,
                 max_episode_steps=MAX_EPISODE_STEPS, **engine_kwargs):
        spec = ENVS[env_name]
        self.envs = []
        for _ in range(int(n_envs)):
            env = PlasticineEnv(**spec, **engine_kwargs)
            env.taichi_env.loss.set_weights(sdf=sdf_loss, density=density_loss, contact=contact_loss, is_soft_contact=soft_contact_loss)
            self.envs.append(env)
        self.num_envs = len(self.envs)
        self.observation_space, self.action_space = self.envs[0].observation_space, self.envs[0].action_space
        self._max_episode_steps = int(max_episode_steps)
        self._elapsed = np.zeros(self.num_envs, dtype=np.int64)

    def reset(self):
        self._elapsed[:] = 0
        return np.stack([e.reset() for e in self.envs])

    def step(self, actions):
        actions = np.asarray(actions, dtype=np.float64)
        assert actions.shape[0] == self.num_envs
        for env, a in zip(self.envs, actions):          # phase 1: enqueue every env's substep graph on its own stream
            env.begin_step(a)
        obs, rew, done, infos = [], [], [], []
        for i, (env, a) in enumerate(zip(self.envs, actions)):      # phase 2: read back (syncs one stream at a time)
            o, r, d, info = env.finish_step(a)
            self._elapsed[i] += 1
            if self._elapsed[i] >= self._max_episode_steps:
                d = True
                info = dict(info, **{"TimeLimit.truncated": True, "terminal_observation": o})
                o = env.reset()
                self._elapsed[i] = 0
            obs.append(o); rew.append(r); done.append(d); infos.append(info)
        return np.stack(obs), np.asarray(rew), np.asarray(done), infos

    def close(self):
        for e in self.envs:
            e.taichi_env.engine.close()
