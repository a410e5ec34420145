"""Open-loop action-sequence optimisation: the loop of `plb/optimizer/solver.py:14-100` on the engine's own tape.

One iteration = reset the env to the start state in trajectory mode, run the horizon under the tape (step + loss after
every step), read the summed loss and d loss / d action, hand the gradient to the optimiser."""
from __future__ import annotations

import numpy as np

from ..config import CfgNode, make_cls_config
from .optim import Adam, Momentum, Optimizer

OPTIMS = {'Adam': Adam, 'Momentum': Momentum}


class Solver:
    def __init__(self, env, logger=None, cfg=None, **kwargs):
        self.env, self.logger = env, logger
        self.cfg = make_cls_config(self, cfg, **kwargs)
        self.optim_cfg = self.cfg.optim
        self.total_steps = 0

    @classmethod
    def default_config(cls):
        return CfgNode(dict(optim=Optimizer.default_config(), n_iters=100, softness=666., horizon=50, init_range=0.,
                            init_sampler='uniform'))

    @staticmethod
    def init_actions(env, cfg):
        if cfg.init_sampler != 'uniform':
            raise NotImplementedError(cfg.init_sampler)
        shape = (cfg.horizon, env.primitives.action_dim)
        return np.random.uniform(-cfg.init_range, cfg.init_range, size=shape)

    def forward(self, sim_state, action):
        """One fwd+bwd episode from `sim_state`: returns (summed loss, gradient [len(action), action_dim])."""
        env, log = self.env, self.logger
        if log is not None:
            log.reset()
        env.set_state(sim_state, self.cfg.softness, False)
        horizon = len(action)
        with env.tape(loss=env.loss.loss):
            for t, a in enumerate(action):
                env.step(a)
                self.total_steps += 1
                info = env.compute_loss()
                if log is not None:
                    log.step(None, None, info['reward'], None, t == horizon - 1, info)
        return env.loss.loss[None], env.primitives.get_grad(horizon)

    def solve(self, init_actions=None, callbacks=()):
        env = self.env
        actions = self.init_actions(env, self.cfg) if init_actions is None else init_actions
        optim = OPTIMS[self.optim_cfg.type](actions, self.optim_cfg)
        start = env.get_state()
        self.total_steps = 0
        best = (1e10, None)
        for _ in range(self.cfg.n_iters):
            self.params = actions.copy()
            loss, grad = self.forward(start['state'], actions)
            if loss < best[0]:
                best = (loss, actions.copy())
            actions = optim.step(grad)
            for cb in callbacks:
                cb(self, optim, loss, grad)
        env.set_state(**start)
        return best[1]


def solve_action(env, path=None, logger=None, args=None, **overrides):
    """`solve_action` (solver.py:86-100) minus the render-to-png tail (the renderer is out of scope)."""
    def opt(name, default):
        return getattr(args, name, overrides.get(name, default))
    env.reset()
    horizon = env._max_episode_steps
    solver = Solver(env.unwrapped.taichi_env, logger, None, n_iters=-(-opt('num_steps', 50 * 200) // horizon),
                    softness=opt('softness', 666.), horizon=horizon,
                    **{"optim.lr": opt('lr', 0.1), "optim.type": opt('optim', 'Adam'), "init_range": 0.0001})
    return solver.solve()
