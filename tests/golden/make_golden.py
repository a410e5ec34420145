#!/usr/bin/env python
"""Generates tests/golden/*.npz from the float64 oracle (oracle/plb_oracle.py).

The reference itself (Taichi 0.7.14) cannot run in this container, so these vectors pin the ORACLE (and through it the CUDA
engine) against regressions; the oracle in turn is pinned to the reference by the Move-v1 loss anchor and by finite
differences (tests/test_oracle.py).  Inputs are fully seeded; re-running this script must reproduce the files bit for bit
up to BLAS summation order.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402
from oracle import plb_oracle as O  # noqa: E402
from plasticinelab_b200.config import load_dict  # noqa: E402
from plasticinelab_b200.engine.shapes import Shapes  # noqa: E402
from plasticinelab_b200.envs.scene import load_target, load_variants  # noqa: E402


def episode_case():
    """Two spheres squeezing a 600-particle ball at quality 0.5 (32^3, 9 substeps/step), 3 env steps, hard contact loss."""
    tree = dict(SIMULATOR=dict(quality=0.5, yield_stress=200.0, max_steps=64),
                SHAPES=[dict(shape='sphere', radius=0.1, init_pos=(0.5, 0.5, 0.5), n_particles=600)],
                PRIMITIVES=[dict(shape='Sphere', radius=0.04, init_pos=(0.38, 0.5, 0.5), friction=0.9, action=dict(dim=3, scale=(0.01,) * 3)),
                            dict(shape='Sphere', radius=0.04, init_pos=(0.62, 0.5, 0.5), friction=0.9, action=dict(dim=3, scale=(0.01,) * 3))])
    cfg = load_dict(tree)
    x0, _ = Shapes(cfg.SHAPES).get()
    t = load_target('Move3D-v1').reshape(32, 2, 32, 2, 32, 2).sum((1, 3, 5))
    t = t * (len(x0) * (1 / 32 * 0.5) ** 2 / t.sum())
    sdf = O.build_target_sdf_c(t, 1 / 32)
    actions = np.random.RandomState(1).uniform(-1, 1, (3, 6))
    out = {}
    for mode in ('taichi', 'argmin'):
        env = O.OracleEnv(cfg, x0, t, target_sdf=sdf, contact_grad=mode)
        r = env.rollout(actions, softness=666.0)
        out[f'grad_{mode}'] = r['grad']
        out['loss'] = np.array(r['loss'])
        out['per_step'] = np.array([[p['loss'], p['contact_loss'], p['density_loss'], p['sdf_loss'], p['iou']] for p in r['per_step']])
        out['final_x'] = r['final_state'][0].numpy()
        out['final_F'] = r['final_state'][3].numpy()
    out['actions'] = actions
    out['target32'] = t
    np.savez_compressed(os.path.join(HERE, 'episode_two_spheres_q0.5.npz'), **out)
    print('episode case: loss', float(out['loss']))


def move_v1_prefix():
    """Stock Move-v1 (10k particles, 64^3): frame-0 loss terms and the state after the first env step (19 substeps)."""
    cfg = load_variants('move.yml', 1)
    x0, _ = Shapes(cfg.SHAPES).get()
    t = load_target(cfg.ENV.loss.target_path)
    sdf = O.build_target_sdf_c(t, 1 / 64)
    env = O.OracleEnv(cfg, x0, t, target_sdf=sdf)
    info0 = env.loss.value(env.initial_state()[0], env.initial_prims())
    a = np.random.RandomState(0).uniform(-1, 1, (1, 6))
    r = env.rollout(a, softness=666.0, with_grad=False)
    sel = np.arange(0, len(x0), 97)
    np.savez_compressed(os.path.join(HERE, 'move_v1_first_step.npz'), action=a, sel=sel, x=r['final_state'][0].numpy()[sel],
                        v=r['final_state'][1].numpy()[sel], F=r['final_state'][3].numpy()[sel],
                        frame0=np.array([info0['loss'], info0['contact_loss'], info0['density_loss'], info0['sdf_loss']]),
                        step_loss=np.array(r['loss']), sdf_sum=np.array(sdf.sum()), sdf_max=np.array(sdf.max()))
    print('move-v1 prefix: frame0', info0, 'step loss', r['loss'])


if __name__ == '__main__':
    entry.build_oracle()
    episode_case()
    move_v1_prefix()
