"""CPU tests of the warp-level scatter kernels (plasticinelab_b200/csrc/plb_warp.cuh).

The shared-memory tile scatter, its flush (runs of consecutive lanes and per-cell groups) and the thread-level kernels the __global__
wrappers call are executed on 32 lock-stepped host threads per warp (tests/host/warp_emul.hpp) and compared with the
sequential direct-scatter bodies, which tests/test_host_emulation.py pins against the float64 oracle.  Differences come
from the summation order only: 1e-12 relative in float64, 3e-4 in float32 (the SVD adjoint amplifies the 1e-7 differences of the gathered grid adjoint).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from plasticinelab_b200 import _capi
import plb_test_helpers as H

D = _capi.dptr
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["grid_in(p2g)", "block flags", "grid_in(g2p+p2g)", "frame 1", "F of frame 2", "g_out(substep 1)", "partial x-adjoint 1",
         "g_out(substep 0)", "dF[1]", "partial x-adjoint 0", "loss mass grid", "g2p.grad: stored successor frame vs recomputed gather"]


@pytest.fixture(scope="module")
def wemul_lib():
    src = os.path.join(ROOT, "tests", "host", "warp_emul.cpp")
    out = os.path.join(ROOT, "tests", "host", "libplb_wemul.so")
    csrc = os.path.join(ROOT, "plasticinelab_b200", "csrc")
    deps = [src, os.path.join(ROOT, "tests", "host", "warp_emul.hpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.isfile(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-x", "c++", src, "-o", out])
    return C.CDLL(out)


def _state(n, seed, box, sort, n_grid):
    """particles in a small box (several per cell); optionally in the engine's order (sorted by base cell)"""
    x, v, Cm, F = H.random_state(n, seed, box[0], box[1])
    if sort:
        b = (x * n_grid - 0.5).astype(np.int64)
        order = np.lexsort((b[:, 2], b[:, 1], b[:, 0]))
        x, v, Cm, F = x[order], v[order], Cm[order], F[order]
    return tuple(np.ascontiguousarray(a) for a in (x, v, Cm, F))


PRIMS = [dict(shape='Sphere', radius=0.05, init_pos=(0.47, 0.5, 0.5), friction=0.9, action=dict(dim=3, scale=(0.01,) * 3))]


# flush_mode 3 = runs of consecutive lanes (the device default), 0 = per-cell groups; two_phase = the 128-register backward form (needs the SVD records)
@pytest.mark.parametrize('two_phase,flush_mode,svd_store', [(0, 3, 0), (0, 0, 0), (0, 3, 1), (1, 3, 1), (1, 0, 1)])
@pytest.mark.parametrize('dtype,tol', [('float64', 1e-12), ('float32', 3e-4)])
@pytest.mark.parametrize('n,box,sort,stored_next', [(100, (0.40, 0.52), True, 1),     # ~2 particles per cell, ragged last warp
                                                    (70, (0.45, 0.50), True, 1),      # one or two cells per warp
                                                    (64, (0.30, 0.70), False, 0),     # unsorted: 32 one-lane groups per warp
                                                    (40, (0.02, 0.12), True, 0)])     # at the domain corner (clamps, boundary)
def test_warp_scatter_kernels_match_direct_bodies(wemul_lib, two_phase, flush_mode, svd_store, dtype, tol, n, box, sort, stored_next):
    cfg = H.small_cfg(PRIMS, n_particles=n, yield_stress=30.0)
    conf, parr, _ = H.c_setup(cfg, n, dtype)
    x, v, Cm, F = _state(n, 3, box, sort, conf.n_grid)
    gx, gv, gC, gF = (np.ascontiguousarray(a) for a in H.random_adjoint(n, 5))
    pose0 = H.pose_array([[0.47, 0.5, 0.5, 1, 0, 0, 0]])
    pose1 = H.pose_array([[0.4702, 0.4999, 0.5001, 1, 0, 0, 0]])
    out = np.full(16, np.nan)
    k = wemul_lib.wemul_check(conf.dtype, two_phase, C.byref(conf), parr, C.c_double(666.0), D(x), D(v), D(F), D(Cm), D(pose0), D(pose1),
                              D(gx), D(gv), D(gF), D(gC), stored_next, flush_mode, svd_store, D(out))
    assert k == len(NAMES)
    for name, dev in zip(NAMES, out[:k]):
        assert dev <= tol, f"{name}: deviation {dev:.3e} (two_phase={two_phase}, flush_mode={flush_mode}, svd_store={svd_store}, {dtype})"


@pytest.mark.parametrize('dtype,tol', [('float64', 1e-12), ('float32', 2e-5)])
@pytest.mark.parametrize('prims', ['spheres', 'capsule', 'box', 'chopsticks'])
def test_grid_adjoint_register_form_matches_array_form(wemul_lib, dtype, tol, prims):
    """k_grid_bwd_sparse_v2's per-thread function (pose gradients of one primitive at a time in registers, float warp
    reduction) against grid_bwd_body (PoseGrad arrays), on the grid a random particle state scatters."""
    from test_host_emulation import PRIM_SETS, _poses
    from oracle import plb_oracle as O
    n = 200
    cfg = H.small_cfg(PRIM_SETS[prims], n_particles=n, yield_stress=30.0, ground_friction=1.5)
    conf, parr, _ = H.c_setup(cfg, n, dtype)
    x, v, Cm, F = (np.ascontiguousarray(a) for a in H.random_state(n, 1, 0.35, 0.65))
    osim = O.OracleSim(dict(cfg.SIMULATOR), [dict(p) for p in cfg.PRIMITIVES])
    pose0, pose1 = _poses(osim, 1)
    gout = np.ascontiguousarray(np.random.RandomState(9).randn(conf.n_grid ** 3, 4))
    out = np.full(8, np.nan)
    k = wemul_lib.wemul_check_grid_bwd(conf.dtype, C.byref(conf), parr, C.c_double(666.0), D(x), D(v), D(F), D(Cm), D(pose0), D(pose1), D(gout), D(out))
    assert k == 4
    assert out[3] > 0, "no node took the contact branch: the test scene does not exercise the pose gradients"
    assert out[0] <= tol, f"g_in deviates by {out[0]:.3e}"
    assert out[1] <= (1e-10 if dtype == 'float64' else 1e-4), f"pose gradients deviate by {out[1]:.3e}"
    assert out[2] == 0.0, "grid_in / g_out not cleared"
