"""GPU parity of the policy path: an episode driven by the state-feedback MLP under the tape -- engine
(plb_gather_particles / plb_scatter_adjoint / plb_action_grad_step / plb_add_pose_adjoint behind engine/nn/mlp.py) against the
float64 oracle's restatement (OracleEnv.rollout_policy, itself checked against finite differences on the CPU).

First green run on a B200: round 2 (gpurun_out/ab/pytest_policy.log); part of the default GPU suite since.  What can be checked
without a GPU is checked there too (tests/test_policy_host.py, the stepwise kinematics scan in tests/test_host_emulation.py,
the oracle in tests/test_oracle.py).  Tolerances: float64 1e-9 loss / 1e-6 gradient,
float32 1e-4 / 5e-2 (as for the action-gradient episodes in test_gpu_parity.py).
"""
import os

import numpy as np
import pytest

import plb_test_helpers as H
from test_gpu_parity import _episode_cfg, _target32
from oracle import plb_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_policy_episode_matches_oracle(dtype):
    from plasticinelab_b200.engine.taichi_env import TaichiEnv
    from plasticinelab_b200.engine.nn.mlp import MLP
    from plasticinelab_b200.optimizer.solver_nn import SolverNN
    cfg = _episode_cfg()
    env = TaichiEnv(cfg, dtype=dtype)
    env.nn = MLP(env.simulator, env.primitives, (16,), activation='relu', n_observed_particles=30)
    env.initialize()
    t32 = _target32(env)
    env.loss.load_target_density(grids=t32)
    env.loss.set_weights(10, 10, 1, False)
    rng = np.random.RandomState(3)
    dims = env.nn.dims
    params = 0.3 * rng.randn(sum(dims[i + 1] * dims[i] + dims[i + 1] for i in range(len(dims) - 1)))
    solver = SolverNN(env, None, None, n_iters=1, softness=666., horizon=3)
    loss, grad = solver.forward(env.get_state()['state'], params)
    oenv = O.OracleEnv(cfg, env.init_particles, t32, target_sdf=O.build_target_sdf_c(t32, 1 / 32), contact_grad='taichi')
    out = oenv.rollout_policy(params, 3, hidden=(16,), n_observed=30, softness=666.0)
    ltol, gtol = (1e-9, 1e-6) if dtype == 'float64' else (1e-4, 5e-2)
    H.record(f"policy[{dtype}]", loss=abs(loss - out['loss']) / abs(out['loss']), grad=H.relerr(grad, out['grad']))
    assert abs(loss - out['loss']) < ltol * abs(out['loss'])
    assert H.relerr(grad, out['grad']) < gtol
