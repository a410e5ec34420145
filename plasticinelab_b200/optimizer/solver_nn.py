"""Policy-parameter optimisation through the simulator: the loop of `plb/optimizer/solver_nn.py:3-123` on the engine's tape.

One iteration = reset to the start state in trajectory mode, run the horizon under the tape with the policy choosing every
action from the observed state (`env.nn.set_action(i, S)`; `env.step()` without an action), read the summed loss and
d loss / d parameters (`env.nn.get_grad()`), hand the gradient to the optimiser (learning rate x 0.001, no bounds)."""
from __future__ import annotations

import numpy as np

from ..config import CfgNode, make_cls_config
from .optim import Optimizer
from .solver import OPTIMS


class SolverNN:
    def __init__(self, env, logger=None, cfg=None, **kwargs):
        self.cfg = make_cls_config(self, cfg, **kwargs)
        self.cfg.optim.lr *= 0.001
        self.cfg.optim.bounds = (-np.inf, np.inf)
        self.logger = logger
        self.optim_cfg = self.cfg.optim
        self.horizon = self.cfg.horizon
        self.env = env
        self.total_steps = 0

    @classmethod
    def default_config(cls):
        return CfgNode(dict(optim=Optimizer.default_config(), n_iters=100, softness=666., horizon=50, init_range=0.,
                            init_sampler='uniform'))

    def forward(self, sim_state, params):
        env, nn, log = self.env, self.env.nn, self.logger
        nn.set_params(params)
        env.set_state(sim_state, self.cfg.softness, False)
        if log is not None:
            log.reset()
        with env.tape(loss=env.loss.loss):
            for i in range(self.horizon):
                nn.set_action(i, env.simulator.substeps)
                env.step()
                self.total_steps += 1
                info = env.compute_loss()
                if log is not None:
                    log.step(None, None, info['reward'], None, i == self.horizon - 1, info)
        return env.loss.loss[None], nn.get_grad()

    def solve(self, callbacks=()):
        env = self.env
        assert hasattr(env, 'nn'), "nn must be an element of env .."
        params = env.nn.get_params()
        optim = OPTIMS[self.optim_cfg.type](env.nn.get_params(), self.optim_cfg)
        start = env.get_state()
        self.total_steps = 0
        best = (1e10, None)
        for _ in range(self.cfg.n_iters):
            self.params = params
            loss, grad = self.forward(start['state'], params)
            if loss < best[0]:
                best = (loss, params.copy())
            params = optim.step(grad)
            for cb in callbacks:
                cb(self, optim, loss, grad)
        env.set_state(**start)
        return best[1]


def init_mlp_params(inp_dim, oup_dim, hidden=(256, 256), seed=None):
    """Flattened W0, b0, W1, b1, ... with torch.nn.Linear's default initialisation, as `solve_nn` builds them
    (solver_nn.py:79-100: a torch MLP is created only to draw the initial weights)."""
    import torch
    if seed is not None:
        torch.manual_seed(seed)
    dims = (inp_dim,) + tuple(hidden) + (oup_dim,)
    out = []
    for i in range(len(dims) - 1):
        layer = torch.nn.Linear(dims[i], dims[i + 1])
        out += [layer.weight.data.double().numpy().reshape(-1), layer.bias.data.double().numpy().reshape(-1)]
    return np.concatenate(out)


def solve_nn(env, path, logger, args):
    """`solve_nn(env, path, logger, args)` (solver_nn.py:73-123) without the rendering loop (the renderer is out of scope)."""
    import os
    os.makedirs(path, exist_ok=True)
    T = env._max_episode_steps
    params = init_mlp_params(env.observation_space.shape[0], env.action_space.shape[0])
    env.reset()
    taichi_env = env.unwrapped.taichi_env
    solver = SolverNN(taichi_env, logger, None, n_iters=(args.num_steps + T - 1) // T, softness=args.softness, horizon=T,
                      **{"optim.lr": args.lr, "optim.type": args.optim, "init_range": 0.0001})
    taichi_env.nn.set_params(params)
    assert np.abs(taichi_env.nn.get_params() - params).max() < 1e-9
    params = solver.solve()
    taichi_env.nn.set_params(params)
    return params
