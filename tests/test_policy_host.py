"""CPU tests of the policy path's host side (engine/nn/mlp.py, optimizer/solver_nn.py) with a stand-in for the engine handle:
the dense layers' hand-written backward against torch autograd, parameter packing, the calls the policy makes on the C ABI."""
import ctypes as C

import numpy as np
import torch

from plasticinelab_b200.engine.nn.mlp import MLP
from plasticinelab_b200.optimizer.solver_nn import init_mlp_params


class _FakeEngine:
    """records the C-ABI calls of MLP and serves fixed observations / action gradients"""
    def __init__(self, n, ga, seed=0):
        rng = np.random.RandomState(seed)
        self.x, self.v, self.ga = rng.rand(n, 3), rng.randn(n, 3), ga
        self.calls = []

    def call(self, name, *args):
        self.calls.append(name)
        if name == "plb_gather_particles":
            slot, idx_p, n, xp, vp = args
            idx = np.ctypeslib.as_array(idx_p, shape=(n,))
            np.ctypeslib.as_array(xp, shape=(n * 3,))[:] = self.x[idx].reshape(-1)
            np.ctypeslib.as_array(vp, shape=(n * 3,))[:] = self.v[idx].reshape(-1)
        elif name == "plb_action_grad_step":
            step, S, outp = args
            np.ctypeslib.as_array(outp, shape=(len(self.ga),))[:] = self.ga
        elif name == "plb_scatter_adjoint":
            idx_p, n, gxp, gvp = args
            self.idx = np.ctypeslib.as_array(idx_p, shape=(n,)).copy()
            self.gx = np.ctypeslib.as_array(gxp, shape=(n, 3)).copy()
            self.gv = np.ctypeslib.as_array(gvp, shape=(n, 3)).copy()
        elif name == "plb_add_pose_adjoint":
            k, gp = args
            self.gpose = getattr(self, "gpose", {})
            self.gpose[k] = np.ctypeslib.as_array(gp, shape=(8,)).copy()


class _Prim:
    shape, state_dim, action_dim = 'Sphere', 7, 3

    def __init__(self, st):
        self.st = np.asarray(st, dtype=np.float64)

    def get_state(self, f):
        return self.st


class _Prims(list):
    state_dim, action_dim = 14, 6

    def set_action(self, s, n, a):
        self.last_action = np.array(a)


class _Sim:
    def __init__(self, engine, n):
        self.engine, self.n_particles, self.substeps = engine, n, 5


def test_policy_backward_matches_autograd_and_routes_observation_adjoints():
    n, A = 47, 6
    rng = np.random.RandomState(1)
    ga = rng.randn(A)
    eng = _FakeEngine(n, ga)
    prims = _Prims([_Prim([0.4, 0.5, 0.5, 1, 0, 0, 0]), _Prim([0.6, 0.5, 0.5, 0.8, 0.6, 0, 0])])
    mlp = MLP(_Sim(eng, n), prims, (16, 12), activation='relu', n_observed_particles=10)
    assert mlp.obs_step == 4 and mlp.obs_num == 11 and mlp.dims == (11 * 6 + 14, 16, 12, 6)
    params = np.concatenate([0.4 * rng.randn(sum(mlp.dims[i + 1] * mlp.dims[i] + mlp.dims[i + 1] for i in range(3))), [0.7]])
    mlp.set_params(params)
    assert mlp.velocity_weight == 0.7 and np.array_equal(mlp.get_params(), params[:-1])
    mlp.set_action(3, 5)
    obs = mlp.observe(3)
    assert np.array_equal(obs[:6], np.concatenate([eng.x[0], 0.7 * eng.v[0]])) and np.array_equal(obs[-7:], prims[1].st)
    assert np.abs(prims.last_action).max() <= 1.0
    mlp.zero_grad_for_tape()
    mlp.backward(3)
    # the same network in torch: loss = ga . clamp(net(obs))
    idx = np.arange(11) * 4
    x = torch.tensor(eng.x[idx], requires_grad=True)
    v = torch.tensor(eng.v[idx], requires_grad=True)
    pose = [torch.tensor(p.st, requires_grad=True) for p in prims]
    Ws = [torch.tensor(w, requires_grad=True) for w in mlp.W]
    bs = [torch.tensor(b, requires_grad=True) for b in mlp.b]
    h = torch.cat([torch.cat([x, v * 0.7], dim=1).reshape(-1)] + pose)
    for i in range(3):
        h = Ws[i] @ h + bs[i]
        if i != 2:
            h = torch.relu(h)
    (torch.clamp(h, -1, 1) * torch.tensor(ga)).sum().backward()
    ref = np.concatenate([t.grad.numpy().reshape(-1) for i in range(3) for t in (Ws[i], bs[i])])
    assert np.abs(mlp.get_grad() - ref).max() < 1e-12
    assert np.array_equal(eng.idx, idx) and np.abs(eng.gx - x.grad.numpy()).max() < 1e-12 and np.abs(eng.gv - v.grad.numpy()).max() < 1e-12
    for k in range(2):
        assert np.abs(eng.gpose[k][:7] - pose[k].grad.numpy()).max() < 1e-12 and eng.gpose[k][7] == 0.0
    assert eng.calls.count("plb_action_grad_step") == 1 and eng.calls.count("plb_scatter_adjoint") == 1


def test_init_mlp_params_layout():
    p = init_mlp_params(20, 4, hidden=(8, 8), seed=0)
    assert p.shape == (20 * 8 + 8 + 8 * 8 + 8 + 8 * 4 + 4,) and np.isfinite(p).all() and np.abs(p).max() < 1.0


def test_vec_env_two_phase_stepping_and_time_limit(monkeypatch):
    """envs/vec_env.py host logic with stand-in envs: every env's begin_step runs before the first finish_step (so the
    envs' work overlaps on the GPU), rewards/observations come back per env, episodes are cut and reset at the step limit."""
    from plasticinelab_b200.envs import vec_env
    log = []

    class FakeBox:
        shape = (2,)

    class FakeEnv:
        count = 0

        def __init__(self, **kw):
            self.i = FakeEnv.count
            FakeEnv.count += 1
            self.t = 0
            self.observation_space = self.action_space = FakeBox()
            self.taichi_env = type("T", (), {"loss": type("L", (), {"set_weights": staticmethod(lambda **k: None)})()})()

        def reset(self):
            self.t = 0
            log.append(("reset", self.i))
            return np.array([self.i, 0.0])

        def begin_step(self, a):
            log.append(("begin", self.i))

        def finish_step(self, a):
            self.t += 1
            log.append(("finish", self.i))
            return np.array([self.i, float(self.t)]), float(a[0]) + self.i, False, {"loss": 0.0}

    monkeypatch.setattr(vec_env, "PlasticineEnv", FakeEnv)
    vec = vec_env.VecPlasticineEnv("Move-v1", 3, max_episode_steps=2)
    obs = vec.reset()
    assert obs.shape == (3, 2)
    log.clear()
    obs, rew, done, infos = vec.step(np.array([[0.1, 0], [0.2, 0], [0.3, 0]]))
    assert [e[0] for e in log] == ["begin"] * 3 + ["finish"] * 3
    assert np.allclose(rew, [0.1, 1.2, 2.3]) and not done.any() and np.array_equal(obs[:, 1], [1, 1, 1])
    obs, rew, done, infos = vec.step(np.zeros((3, 2)))
    assert done.all() and np.array_equal(obs[:, 1], [0, 0, 0])                  # auto-reset: first observation of the new episode
    assert all(np.array_equal(i["terminal_observation"], [k, 2.0]) for k, i in enumerate(infos))
