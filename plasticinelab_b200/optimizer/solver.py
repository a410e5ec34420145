"""Open-loop action-sequence solver; same loop as `plb/optimizer/solver.py:14-100` with the engine's tape in place
of `ti.Tape`."""
import numpy as np

from ..config import CfgNode, make_cls_config
from .optim import Adam, Momentum, Optimizer

OPTIMS = {'Adam': Adam, 'Momentum': Momentum}


class Solver:
    def __init__(self, env, logger=None, cfg=None, **kwargs):
        self.cfg = make_cls_config(self, cfg, **kwargs)
        self.optim_cfg = self.cfg.optim
        self.env = env
        self.logger = logger

    def forward(self, sim_state, action):
        """One fwd+bwd episode: returns (summed loss, d loss / d action[T, A])."""
        env = self.env
        if self.logger is not None:
            self.logger.reset()
        env.set_state(sim_state, self.cfg.softness, False)
        with env.tape(loss=env.loss.loss):
            for i in range(len(action)):
                env.step(action[i])
                self.total_steps += 1
                loss_info = env.compute_loss()
                if self.logger is not None:
                    self.logger.step(None, None, loss_info['reward'], None, i == len(action) - 1, loss_info)
        loss = env.loss.loss[None]
        return loss, env.primitives.get_grad(len(action))

    def solve(self, init_actions=None, callbacks=()):
        env = self.env
        if init_actions is None:
            init_actions = self.init_actions(env, self.cfg)
        optim = OPTIMS[self.optim_cfg.type](init_actions, self.optim_cfg)
        env_state = env.get_state()
        self.total_steps = 0
        best_action, best_loss = None, 1e10
        actions = init_actions
        for _ in range(self.cfg.n_iters):
            self.params = actions.copy()
            loss, grad = self.forward(env_state['state'], actions)
            if loss < best_loss:
                best_loss, best_action = loss, actions.copy()
            actions = optim.step(grad)
            for callback in callbacks:
                callback(self, optim, loss, grad)
        env.set_state(**env_state)
        return best_action

    @staticmethod
    def init_actions(env, cfg):
        if cfg.init_sampler != 'uniform':
            raise NotImplementedError
        return np.random.uniform(-cfg.init_range, cfg.init_range, size=(cfg.horizon, env.primitives.action_dim))

    @classmethod
    def default_config(cls):
        cfg = CfgNode()
        cfg.optim = Optimizer.default_config()
        cfg.n_iters = 100
        cfg.softness = 666.
        cfg.horizon = 50
        cfg.init_range = 0.
        cfg.init_sampler = 'uniform'
        return cfg


def solve_action(env, path=None, logger=None, args=None, **overrides):
    """`solve_action` (solver.py:86-100) without the rendering/video tail (renderer is out of scope)."""
    env.reset()
    taichi_env = env.unwrapped.taichi_env
    T = env._max_episode_steps
    num_steps = getattr(args, 'num_steps', overrides.get('num_steps', 50 * 200))
    kw = {"optim.lr": getattr(args, 'lr', overrides.get('lr', 0.1)),
          "optim.type": getattr(args, 'optim', overrides.get('optim', 'Adam')), "init_range": 0.0001}
    solver = Solver(taichi_env, logger, None, n_iters=(num_steps + T - 1) // T,
                    softness=getattr(args, 'softness', overrides.get('softness', 666.)), horizon=T, **kw)
    return solver.solve()
