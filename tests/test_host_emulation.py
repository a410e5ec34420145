"""CPU tests: the kernel bodies (run through tests/host/emul.cpp) against the float64 oracle.

These cover the math that ships in the CUDA kernels -- forward substep, hand-derived adjoint, loss adjoint, primitive
kinematics -- on the build box, which has no GPU.  Tolerances: float64 instantiation 1e-9 relative (different but
equivalent operation order and SVD algorithm), float32 instantiation 2e-4 relative on one substep.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import plb_oracle as O
from plasticinelab_b200 import _capi
import plb_test_helpers as H

D = _capi.dptr

PRIM_SETS = {
    'none': [],
    'spheres': [dict(shape='Sphere', radius=0.08, init_pos=(0.45, 0.5, 0.5), friction=0.9, action=dict(dim=3, scale=(0.01,) * 3)),
                dict(shape='Sphere', radius=0.06, init_pos=(0.6, 0.45, 0.55), friction=0.5, action=dict(dim=3, scale=(0.01,) * 3))],
    'capsule': [dict(shape='Capsule', h=0.12, r=0.05, init_pos=(0.5, 0.5, 0.5), init_rot=(0.9, 0.1, 0.3, -0.2), friction=0.7,
                     action=dict(dim=6, scale=(0.01,) * 6))],
    'cylinder': [dict(shape='Cylinder', h=0.12, r=0.08, init_pos=(0.5, 0.45, 0.5), init_rot=(0.8, 0.3, -0.1, 0.2), friction=0.9)],
    'torus': [dict(shape='Torus', tx=0.12, ty=0.05, init_pos=(0.5, 0.5, 0.5), init_rot=(0.7, 0.2, 0.1, 0.6), friction=0.9,
                   action=dict(dim=3, scale=(0.004,) * 3))],
    'chopsticks': [dict(shape='Chopsticks', h=0.2, r=0.04, init_pos=(0.5, 0.55, 0.5), init_rot=(0.9, 0.2, 0.1, 0.1), init_gap=0.1,
                        friction=10., action=dict(dim=7, scale=(0.02, 0.02, 0.02, 0.04, 0.04, 0.04, 0.02)))],
    'box': [dict(shape='Box', size=(0.1, 0.07, 0.09), init_pos=(0.5, 0.5, 0.5), init_rot=(0.9, 0.1, -0.3, 0.2), friction=0.9,
                 action=dict(dim=6, scale=(0.01,) * 6))],
}


def _poses(osim, seed):
    """pose f from the cfg (rotation normalised), pose f+1 = small random rigid motion of it."""
    rng = np.random.RandomState(77 + seed)
    p0, p1 = [], []
    for p in osim.prims:
        s = p.init_state().numpy().copy()
        s[3:7] /= np.linalg.norm(s[3:7])
        s1 = s.copy()
        s1[0:3] += 2e-4 * rng.randn(3)
        q = s1[3:7] + 2e-3 * rng.randn(4)
        s1[3:7] = q / np.linalg.norm(q)
        if len(s) == 8:
            s1[7] = s[7] - 1e-4
        p0.append(s); p1.append(s1)
    return H.pose_array(p0), H.pose_array(p1)


def _run_pair(emul_lib, name, dtype, seed, softness, gf, n=300):
    cfg = H.small_cfg(PRIM_SETS[name], n_particles=n, ground_friction=gf, yield_stress=30.0)
    osim = O.OracleSim(dict(cfg.SIMULATOR), [dict(p) for p in cfg.PRIMITIVES])
    osim.set_materials(n)
    osim.set_softness(softness)
    lo, hi = (0.02, 0.3) if seed % 2 else (0.3, 0.7)      # odd seeds sit on the floor / walls
    x, v, Cm, F = H.random_state(n, seed, lo, hi)
    pose0, pose1 = _poses(osim, seed)
    conf, parr, _ = H.c_setup(cfg, n, dtype)
    return cfg, osim, (x, v, Cm, F), (pose0, pose1), conf, parr


@pytest.mark.parametrize('name,softness,gf', [('none', 0.0, 1.5), ('spheres', 666.0, 1.5), ('spheres', 0.0, 0.0),
                                              ('capsule', 666.0, 100.0), ('cylinder', 666.0, 0.3), ('torus', 666.0, 100.0),
                                              ('chopsticks', 666.0, 0.0), ('box', 666.0, 1.5)])
@pytest.mark.parametrize('seed', [0, 1])
def test_substep_forward_and_adjoint_f64(emul_lib, name, softness, gf, seed):
    n = 300
    cfg, osim, (x, v, Cm, F), (pose0, pose1), conf, parr = _run_pair(emul_lib, name, 'float64', seed, softness, gf, n)
    st = tuple(torch.as_tensor(a) for a in (x, v, Cm, F))
    pf, pf1 = H.oracle_prim_states(osim, pose0), H.oracle_prim_states(osim, pose1)
    (ox, ov, oC, oF), (gvi, gm, gvo) = osim.substep(st, pf, pf1, return_grid=True)
    xo, vo, Fo, Co = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
    G = conf.n_grid ** 3
    gin, gout = np.zeros((G, 4)), np.zeros((G, 4))
    emul_lib.emul_substep_fwd(conf.dtype, C.byref(conf), parr, C.c_double(softness), D(x), D(v), D(F), D(Cm), D(pose0), D(pose1),
                              D(xo), D(vo), D(Fo), D(Co), D(gin), D(gout))
    assert H.relerr(gin[:, :3], gvi.numpy()) < 1e-10 and H.relerr(gin[:, 3], gm.numpy()) < 1e-12
    assert H.relerr(gout[:, :3], gvo.numpy()) < 1e-9
    assert H.relerr(xo, ox.numpy()) < 1e-12 and H.relerr(vo, ov.numpy()) < 1e-9
    assert H.relerr(Co, oC.numpy()) < 1e-9 and H.relerr(Fo, oF.numpy()) < 1e-9
    # adjoint
    adj = H.random_adjoint(n, seed)            # gx, gv, gC, gF order below
    gxn, gvn, gCn, gFn = adj
    o_adj, o_g0, o_g1 = osim.substep_vjp(st, pf, pf1, tuple(torch.as_tensor(a) for a in (gxn, gvn, gCn, gFn)))
    gx, gv, gF, gC = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
    P = max(len(osim.prims), 1)
    gp0, gp1 = np.zeros((P, 8)), np.zeros((P, 8))
    emul_lib.emul_substep_bwd(conf.dtype, C.byref(conf), parr, C.c_double(softness), D(x), D(v), D(F), D(Cm), D(pose0), D(pose1),
                              D(gxn), D(gvn), D(gFn), D(gCn), D(gx), D(gv), D(gF), D(gC), D(gp0), D(gp1))
    assert H.relerr(gx, o_adj[0].numpy()) < 1e-8
    assert H.relerr(gv, o_adj[1].numpy()) < 1e-9
    assert H.relerr(gC, o_adj[2].numpy()) < 1e-8
    assert H.relerr(gF, o_adj[3].numpy()) < 1e-8
    for k, p in enumerate(osim.prims):
        d = p.state_dim
        ref0, ref1 = o_g0[k].numpy(), o_g1[k].numpy()
        scale = max(np.abs(ref0).max(), np.abs(ref1).max(), 1e-12)
        assert np.abs(gp0[k, :d] - ref0).max() < 1e-7 * scale + 1e-12, (k, gp0[k], ref0)
        assert np.abs(gp1[k, :d] - ref1).max() < 1e-7 * scale + 1e-12, (k, gp1[k], ref1)


@pytest.mark.parametrize('name,softness', [('spheres', 666.0), ('torus', 666.0)])
def test_substep_f32_close_to_f64_oracle(emul_lib, name, softness):
    n = 300
    cfg, osim, (x, v, Cm, F), (pose0, pose1), conf, parr = _run_pair(emul_lib, name, 'float32', 0, softness, 1.5, n)
    st = tuple(torch.as_tensor(a) for a in (x, v, Cm, F))
    pf, pf1 = H.oracle_prim_states(osim, pose0), H.oracle_prim_states(osim, pose1)
    ox, ov, oC, oF = osim.substep(st, pf, pf1)
    xo, vo, Fo, Co = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
    emul_lib.emul_substep_fwd(conf.dtype, C.byref(conf), parr, C.c_double(softness), D(x), D(v), D(F), D(Cm), D(pose0), D(pose1),
                              D(xo), D(vo), D(Fo), D(Co), None, None)
    assert H.relerr(xo, ox.numpy()) < 1e-6 and H.relerr(vo, ov.numpy()) < 2e-4
    assert H.relerr(Co, oC.numpy()) < 2e-4 and H.relerr(Fo, oF.numpy()) < 1e-5
    gxn, gvn, gCn, gFn = H.random_adjoint(n, 0)
    o_adj, _, _ = osim.substep_vjp(st, pf, pf1, tuple(torch.as_tensor(a) for a in (gxn, gvn, gCn, gFn)))
    gx, gv, gF, gC = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
    gp0, gp1 = np.zeros((8, 8)), np.zeros((8, 8))
    emul_lib.emul_substep_bwd(conf.dtype, C.byref(conf), parr, C.c_double(softness), D(x), D(v), D(F), D(Cm), D(pose0), D(pose1),
                              D(gxn), D(gvn), D(gFn), D(gCn), D(gx), D(gv), D(gF), D(gC), D(gp0), D(gp1))
    for a, b in zip((gx, gv, gC, gF), o_adj):
        assert H.relerr(a, b.numpy()) < 2e-3


def test_svd_convention_and_accuracy(emul_lib):
    rng = np.random.RandomState(3)
    for dtype, tol in ((_capi.PLB_F64, 1e-13), (_capi.PLB_F32, 3e-6)):
        for i in range(200):
            F = np.eye(3) + 0.3 * rng.randn(3, 3) if i % 3 else rng.randn(3, 3)
            if i == 7:
                F = np.eye(3)
            U, s, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
            emul_lib.emul_svd(dtype, D(np.ascontiguousarray(F)), D(U), D(s), D(V))
            assert abs(np.linalg.det(U) - 1) < 10 * tol and abs(np.linalg.det(V) - 1) < 10 * tol
            assert np.abs(U @ np.diag(s) @ V.T - F).max() < tol * max(1.0, np.abs(F).max()) * 10
            assert s[0] >= s[1] >= abs(s[2]) - 10 * tol
            assert np.sign(s[2]) == np.sign(np.linalg.det(F)) or abs(np.linalg.det(F)) < 1e-6


def test_svd_warm_start_chain_stays_accurate(emul_lib):
    """Warm-started Jacobi (V of the previous substep as the starting point, plb_svd.cuh): a chain of 2000 slowly drifting
    matrices, each decomposition started from the previous V, stays an SVD to working precision (V is re-orthonormalised at
    every start, so nothing accumulates), keeps the convention, and U V^T / U S V^T agree with the cold start."""
    rng = np.random.RandomState(8)
    for dtype, tol in ((_capi.PLB_F64, 1e-13), (_capi.PLB_F32, 3e-6)):
        F = np.eye(3) + 0.2 * rng.randn(3, 3)
        U, s, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
        emul_lib.emul_svd(dtype, D(np.ascontiguousarray(F)), D(U), D(s), D(V))
        for i in range(2000):
            F = (np.eye(3) + 2e-3 * rng.randn(3, 3)) @ F
            if i == 1000:
                F = np.eye(3) + 1e-9 * rng.randn(3, 3)          # (nearly) degenerate: any V is a valid start
            W = V.copy()
            emul_lib.emul_svd_warm(dtype, D(np.ascontiguousarray(F)), D(W), D(U), D(s), D(V))
            assert abs(np.linalg.det(U) - 1) < 10 * tol and abs(np.linalg.det(V) - 1) < 10 * tol
            assert np.abs(V.T @ V - np.eye(3)).max() < 10 * tol and np.abs(U.T @ U - np.eye(3)).max() < 10 * tol
            assert np.abs(U @ np.diag(s) @ V.T - F).max() < tol * max(1.0, np.abs(F).max()) * 10
            assert s[0] >= s[1] >= abs(s[2]) - 10 * tol
        Uc, sc, Vc = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
        emul_lib.emul_svd(dtype, D(np.ascontiguousarray(F)), D(Uc), D(sc), D(Vc))
        assert np.abs(U @ V.T - Uc @ Vc.T).max() < 100 * tol and np.abs(s - sc).max() < 100 * tol


@pytest.mark.parametrize('fscale,ys', [(0.004, 30.0), (0.1, 1e9), (0.0, 50.0)])
def test_substep_adjoint_mixed_and_elastic_f64(emul_lib, fscale, ys):
    """Return-mapping branch coverage: ~half of the particles yield / none yield / exactly F = I (degenerate SVD)."""
    n = 300
    cfg = H.small_cfg(PRIM_SETS['spheres'], n_particles=n, ground_friction=1.5, yield_stress=ys)
    osim = O.OracleSim(dict(cfg.SIMULATOR), [dict(p) for p in cfg.PRIMITIVES])
    osim.set_materials(n)
    osim.set_softness(666.0)
    x, v, Cm, F = H.random_state(n, 5, 0.3, 0.7)
    rng = np.random.RandomState(11)
    F = np.eye(3)[None] + fscale * rng.randn(n, 3, 3)
    if fscale == 0.0:
        Cm = np.zeros_like(Cm)
    pose0, pose1 = _poses(osim, 0)
    conf, parr, _ = H.c_setup(cfg, n, 'float64')
    st = tuple(torch.as_tensor(a) for a in (x, v, Cm, F))
    pf, pf1 = H.oracle_prim_states(osim, pose0), H.oracle_prim_states(osim, pose1)
    ox, ov, oC, oF = osim.substep(st, pf, pf1)
    if fscale == 0.004:
        frac = float((torch.linalg.norm(oF - (torch.eye(3) + osim.dt * st[2]) @ st[3], dim=(1, 2)) > 1e-12).double().mean())
        assert 0.2 < frac < 0.8, frac
    xo, vo, Fo, Co = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
    emul_lib.emul_substep_fwd(conf.dtype, C.byref(conf), parr, C.c_double(666.0), D(x), D(v), D(F), D(Cm), D(pose0), D(pose1),
                              D(xo), D(vo), D(Fo), D(Co), None, None)
    assert H.relerr(Fo, oF.numpy()) < 1e-10 and H.relerr(vo, ov.numpy()) < 1e-9
    gxn, gvn, gCn, gFn = H.random_adjoint(n, 5)
    o_adj, _, _ = osim.substep_vjp(st, pf, pf1, tuple(torch.as_tensor(a) for a in (gxn, gvn, gCn, gFn)))
    gx, gv, gF, gC = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
    gp0, gp1 = np.zeros((8, 8)), np.zeros((8, 8))
    emul_lib.emul_substep_bwd(conf.dtype, C.byref(conf), parr, C.c_double(666.0), D(x), D(v), D(F), D(Cm), D(pose0), D(pose1),
                              D(gxn), D(gvn), D(gFn), D(gCn), D(gx), D(gv), D(gF), D(gC), D(gp0), D(gp1))
    for a, b in zip((gx, gv, gC, gF), o_adj):
        assert H.relerr(a, b.numpy()) < 1e-7


@pytest.mark.parametrize('name', ['capsule', 'chopsticks', 'rollingpin', 'spheres'])
def test_kinematics_forward_and_adjoint(emul_lib, name):
    """forward_kinematics + its adjoint in the engine's C++ (csrc/plb_kinematics.hpp) against torch autograd through the
    oracle's `fk` (base class: left-multiplied rotation; Chopsticks: right-multiplied + gap; RollingPin: custom)."""
    sets = dict(PRIM_SETS)
    sets['rollingpin'] = [dict(shape='RollingPin', h=0.3, r=0.03, init_pos=(0.5, 0.3, 0.5), init_rot=(0.707, 0.707, 0., 0.), friction=0.9,
                               lower_bound=(0., 0.05, 0.), action=dict(dim=3, scale=(0.7, 0.005, 0.005)))]
    cfg = H.small_cfg(sets[name], n_particles=10)
    osim = O.OracleSim(dict(cfg.SIMULATOR), [dict(p) for p in cfg.PRIMITIVES])
    rng = np.random.RandomState(12)
    for k, prim in enumerate(osim.prims):
        desc = _capi.primitive_desc(dict(cfg.PRIMITIVES[k]))
        st = prim.init_state().numpy().copy()
        st[3:7] = st[3:7] / np.linalg.norm(st[3:7])
        for trial in range(3):
            v = 0.01 * rng.randn(3)
            w = 0.02 * rng.randn(3) if trial < 2 else np.zeros(3)          # trial 2: |w| <= 1e-9 branch
            gv = 0.003 * rng.rand()
            if trial == 1:
                st[1] = float(prim.lower[1]) + 1e-4                        # lower clamp active on y
                v[1] = -abs(v[1]) - 1e-3
            st8 = np.zeros(8); st8[:len(st)] = st
            out = np.zeros(8)
            emul_lib.emul_fk(C.byref(desc), D(st8), D(np.ascontiguousarray(v)), D(np.ascontiguousarray(w)), C.c_double(gv), D(out))
            ts = torch.as_tensor(st.copy()).requires_grad_(True)
            tv = torch.as_tensor(v.copy()).requires_grad_(True)
            tw = torch.as_tensor(w.copy()).requires_grad_(True)
            tg = torch.tensor(gv, dtype=torch.float64, requires_grad=True)
            ref = prim.fk(ts, tv, tw, tg)
            assert np.abs(out[:len(st)] - ref.detach().numpy()).max() < 1e-13
            gout = rng.randn(len(st))
            grads = torch.autograd.grad(ref, [ts, tv, tw, tg], grad_outputs=torch.as_tensor(gout), allow_unused=True)
            grads = [np.zeros_like(x.detach().numpy()) if g is None else g.numpy() for g, x in zip(grads, (ts, tv, tw, tg))]
            g8 = np.zeros(8); g8[:len(st)] = gout
            gst, gvel, gw, ggv = np.zeros(8), np.zeros(3), np.zeros(3), C.c_double(0.0)
            emul_lib.emul_fk_bwd(C.byref(desc), D(st8), D(np.ascontiguousarray(v)), D(np.ascontiguousarray(w)), C.c_double(gv), D(g8),
                                 D(gst), D(gvel), D(gw), C.byref(ggv))
            assert np.abs(gst[:len(st)] - grads[0]).max() < 1e-12, (name, trial, gst, grads[0])
            assert np.abs(gvel - grads[1]).max() < 1e-12
            if name != 'rollingpin':
                assert np.abs(gw - grads[2]).max() < 1e-12
            if name == 'chopsticks':
                assert abs(ggv.value - float(grads[3])) < 1e-12


@pytest.mark.parametrize('mode', ['hard_taichi', 'hard_argmin', 'soft'])
def test_loss_value_and_adjoint(emul_lib, mode):
    """Per-step loss (density + target-SDF + contact) and its adjoint wrt particle positions and primitive poses
    (plb/engine/losses/loss.py:116-153,186-237), hard contact in both gradient conventions and the soft minimum."""
    n = 500
    prims = PRIM_SETS['spheres'] + [dict(shape='Capsule', h=0.1, r=0.03, init_pos=(0.45, 0.62, 0.5), init_rot=(0.9, 0.1, 0.3, 0.2),
                                         friction=0.9, action=dict(dim=6, scale=(0.01,) * 6))]
    cfg = H.small_cfg(prims, n_particles=n)
    osim = O.OracleSim(dict(cfg.SIMULATOR), [dict(p) for p in cfg.PRIMITIVES])
    osim.set_materials(n)
    rng = np.random.RandomState(21)
    x = rng.uniform(0.3, 0.7, (n, 3))
    G = osim.n_grid
    target = np.zeros((G, G, G)); target[10:18, 10:20, 12:20] = rng.rand(8, 10, 8) * 1e-3
    tsdf = rng.rand(G, G, G)
    soft = mode == 'soft'
    oloss = O.OracleLoss(osim, target, (3.0, 7.0, 2.0), soft_contact=soft, contact_grad='taichi' if mode != 'hard_argmin' else 'argmin',
                         target_sdf=tsdf)
    poses = [p.init_state().numpy().copy() for p in osim.prims]
    for s in poses:
        s[3:7] /= np.linalg.norm(s[3:7])
    pf = [torch.as_tensor(s) for s in poses]
    val = oloss.value(torch.as_tensor(x), pf)
    ogx, ogp = oloss.vjp(torch.as_tensor(x), pf)
    conf, parr, _ = H.c_setup(cfg, n, 'float64')
    out4, gx, gp = np.zeros(4), np.zeros((n, 3)), np.zeros((len(poses), 8))
    emul_lib.emul_loss(conf.dtype, C.byref(conf), parr, D(np.ascontiguousarray(x)), D(H.pose_array(poses)), D(target), D(tsdf),
                       C.c_double(3.0), C.c_double(7.0), C.c_double(2.0), int(mode != 'hard_argmin'), D(out4), D(gx), D(gp), int(soft))
    assert abs(out4[0] - val['loss']) < 1e-12 * abs(val['loss']) and abs(out4[1] - val['contact_loss']) < 1e-14 + 1e-12 * val['contact_loss']
    assert H.relerr(gx, ogx.numpy()) < 1e-10
    for k in range(len(poses)):
        ref = ogp[k].numpy()
        assert np.abs(gp[k, :7] - ref).max() < 1e-9 * max(np.abs(ref).max(), 1e-12) + 1e-13, (k, gp[k], ref)


def test_stepwise_action_gradient_scan_equals_whole_episode_scan(emul_lib):
    """plb_action_grad_step (policy path: one env step at a time, carried pose adjoint, injected observation adjoints) against
    plb_get_action_grad's whole-episode scan on random pose adjoints.  Exact up to summation order (1e-12)."""
    rng = np.random.RandomState(4)
    names = ['spheres', 'capsule', 'chopsticks']
    prims = [dict(p) for n in names for p in PRIM_SETS[n]]
    descs = [_capi.primitive_desc(p) for p in prims]
    parr = (_capi.PrimitiveDesc * len(descs))(*descs)
    n_steps, S, MAXP = 4, 5, 8
    nf = n_steps * S
    traj = np.zeros((nf + 1, MAXP, 8)); vel = np.zeros((nf + 1, MAXP, 8))
    for k, d in enumerate(descs):
        st = np.array(list(d.init_state), dtype=np.float64)
        st[3:7] /= np.linalg.norm(st[3:7])
        traj[0, k] = st
        for f in range(nf):
            vel[f, k, :3] = 2e-3 * rng.randn(3); vel[f, k, 3:6] = 5e-3 * rng.randn(3); vel[f, k, 6] = 1e-3 * rng.rand()
            out = np.zeros(8)
            emul_lib.emul_fk(C.byref(d), D(np.ascontiguousarray(traj[f, k])), D(np.ascontiguousarray(vel[f, k, :3])),
                             D(np.ascontiguousarray(vel[f, k, 3:6])), C.c_double(vel[f, k, 6]), D(out))
            traj[f + 1, k] = out
    g = np.zeros((nf + 1, MAXP, 8)); g[:, :len(descs)] = rng.randn(nf + 1, len(descs), 8)
    inject = np.ascontiguousarray(rng.randn(n_steps, len(descs), 8))
    emul_lib.emul_action_grad.restype = C.c_int
    for inj in (None, inject):
        res = []
        for chunked in (0, 1):
            A = sum(d.action_dim for d in descs)
            out = np.zeros((n_steps, A))
            r = emul_lib.emul_action_grad(parr, len(descs), D(traj), D(vel), D(g), n_steps, S, chunked, D(inj), D(out))
            assert r == A
            res.append(out)
        assert np.abs(res[0]).max() > 0
        assert np.abs(res[0] - res[1]).max() < 1e-12 * max(1.0, np.abs(res[0]).max())
