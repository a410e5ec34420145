"""GPU test: every selectable kernel variant of the engine produces the same episode (loss, action gradient, final state).

The variants are chosen with environment switches read at plb_create (plb_engine.cu): chunked TMA-window kernels or per-thread
gathers, register caps of the fused particle kernels, SVD records, block list per env step or per substep, flush of the scatter
tiles, programmatic dependent launch, the forked grid pre-stage of the backward graphs, env-step re-sort.  The conservative configuration (everything serial, full tiles)
is compared against the float64 oracle by tests/test_gpu_parity.py; here all other variants are compared with it on the
same episode, in one process.  Tolerances: float64 1e-10 on the loss, 1e-7 on the gradient (summation order only); float32 1e-6 on the loss,
2e-5 on the gradient, 2e-6 on positions (about 10x what a B200 measured: <= 5e-9, 1.8e-6, 1.8e-7, gpurun_out/ab/pytest_variants.log).
"""
import os

import numpy as np
import pytest

import plb_test_helpers as H
from test_gpu_parity import _episode_cfg, _target32

pytestmark = pytest.mark.gpu

KEYS = ["PLB_BWD_OVERLAP", "PLB_FWD_MINB", "PLB_BWD_MINB", "PLB_FUSE", "PLB_GRID_BWD_V2", "PLB_SVD_STORE", "PLB_ENV_LIST", "PLB_TILE", "PLB_TILE_BWD",
        "PLB_TILE_FWD_MINB", "PLB_SVD_WARM", "PLB_FLUSH_MODE", "PLB_PDL", "PLB_WINDOW_FOLLOW", "PLB_RESORT"]
VARIANTS = {
    # everything serial and per substep: per-thread gathers, per-cell group flush, no SVD records, array-form grid adjoint, no PDL
    "conservative": dict(PLB_BWD_OVERLAP=0, PLB_FWD_MINB=5, PLB_BWD_MINB=3, PLB_GRID_BWD_V2=0, PLB_FLUSH_MODE=0, PLB_SVD_STORE=0, PLB_ENV_LIST=0,
                         PLB_TILE=0, PLB_PDL=0, PLB_WINDOW_FOLLOW=0, PLB_RESORT=0),
    "defaults": {},
    "unfused": dict(PLB_FUSE=0),
    "grid_bwd_arrays": dict(PLB_GRID_BWD_V2=0),
    "no_svd_store": dict(PLB_SVD_STORE=0),
    "svd_store_loose": dict(PLB_SVD_STORE=1, PLB_BWD_MINB=3),
    "substep_list": dict(PLB_ENV_LIST=0),
    # chunked TMA-window kernels (plb_tile.cuh) are the default since round 2; PLB_TILE=0 = the per-thread-gather kernels
    "no_tile": dict(PLB_TILE=0),
    "no_tile_tight": dict(PLB_TILE=0, PLB_FWD_MINB=6),
    "tile_no_svd_store": dict(PLB_SVD_STORE=0, PLB_BWD_MINB=3),
    "tile_serial_bwd": dict(PLB_BWD_OVERLAP=0),
    "tile_bwd": dict(PLB_TILE_BWD=1),                      # chunked backward kernels too (slower on a B200, kept selectable)
    "tile_bwd_no_svd_store": dict(PLB_TILE_BWD=1, PLB_SVD_STORE=0, PLB_BWD_MINB=3),
    "tile_fwd_96": dict(PLB_TILE_FWD_MINB=5),
    "svd_cold": dict(PLB_SVD_WARM=0),
    "no_pdl": dict(PLB_PDL=0),                             # env-step graphs without programmatic dependent launch edges
    "window_fixed": dict(PLB_WINDOW_FOLLOW=0),             # TMA windows fixed at the sort (default: re-centred on the material every env step)
    "flush_groups": dict(PLB_FLUSH_MODE=0),                # per-cell group flush in the per-warp kernels (default: runs of consecutive lanes)
    "no_resort": dict(PLB_RESORT=0),                       # (default: particles re-sorted at every env-step boundary, adjoint un-permuted on the way back)
    "no_resort_no_tile": dict(PLB_RESORT=0, PLB_TILE=0),
    "substep_list_no_svd_groups": dict(PLB_ENV_LIST=0, PLB_SVD_STORE=0, PLB_BWD_MINB=3, PLB_FLUSH_MODE=0),
}


def _run(monkeypatch, env_vars, dtype):
    from plasticinelab_b200.engine.taichi_env import TaichiEnv
    from plasticinelab_b200.optimizer.solver import Solver
    for k in KEYS:
        monkeypatch.delenv(k, raising=False)
    for k, v in env_vars.items():
        monkeypatch.setenv(k, str(v))
    cfg = _episode_cfg()
    env = TaichiEnv(cfg, dtype=dtype)
    env.initialize()
    env.loss.load_target_density(grids=_target32(env))
    env.loss.set_weights(10, 10, 1, False)
    actions = np.random.RandomState(1).uniform(-1, 1, (3, 6))
    solver = Solver(env, None, None, n_iters=1, softness=666., horizon=3)
    solver.total_steps = 0
    loss, grad = solver.forward(env.get_state()['state'], actions)
    x = env.simulator.get_state(env.simulator.cur)[0]
    env.engine.close()
    return loss, grad, x


@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_kernel_variants_agree(monkeypatch, dtype):
    ref = _run(monkeypatch, VARIANTS["conservative"], dtype)
    ltol, gtol, xtol = (1e-10, 1e-7, 1e-10) if dtype == 'float64' else (1e-6, 2e-5, 2e-6)
    bad = []
    for name, env_vars in VARIANTS.items():
        if name == "conservative":
            continue
        loss, grad, x = _run(monkeypatch, env_vars, dtype)
        dl, dg, dx = abs(loss - ref[0]) / abs(ref[0]), H.relerr(grad, ref[1]), float(np.abs(x - ref[2]).max())
        print(f"[variants {dtype}] {name:12s} loss {dl:.2e} grad {dg:.2e} x {dx:.2e}")
        if not (dl < ltol and dg < gtol and dx < xtol):
            bad.append((name, dl, dg, dx))
    assert not bad, bad


def test_vec_env_matches_sequential_envs():
    """K envs stepped as 'enqueue all, read all' (envs/vec_env.py) give what K separately stepped envs give (float64: 1e-9)."""
    from plasticinelab_b200.envs import make
    from plasticinelab_b200.envs.vec_env import VecPlasticineEnv
    K, T = 3, 4
    rng = np.random.RandomState(0)
    vec = VecPlasticineEnv('Move-v1', K, dtype='float64')
    acts = rng.uniform(-1, 1, (T, K, vec.action_space.shape[0]))
    obs0 = vec.reset()
    out = [vec.step(acts[t]) for t in range(T)]
    single = make('Move-v1', dtype='float64')
    for k in range(K):
        o = single.reset()
        assert np.abs(o - obs0[k]).max() < 1e-12
        for t in range(T):
            o, r, d, info = single.step(acts[t, k])
            assert np.abs(o - out[t][0][k]).max() < 1e-9 and abs(r - out[t][1][k]) < 1e-9 * max(1.0, abs(r))
    vec.close()
