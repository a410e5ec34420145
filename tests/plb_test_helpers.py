"""Shared helpers for the parity tests (inputs per SURVEY.md 8(d) 'kernel-level parity inputs')."""
import ctypes as C

import numpy as np
import torch

from plasticinelab_b200 import _capi
from plasticinelab_b200.config import load_dict


def small_cfg(prims, quality=0.5, n_particles=300, **sim):
    tree = dict(SIMULATOR=dict(quality=quality, **sim), PRIMITIVES=prims,
                SHAPES=[dict(shape='box', width=(0.2, 0.2, 0.2), init_pos=(0.5, 0.5, 0.5), n_particles=n_particles)])
    return load_dict(tree)


def random_state(n, seed, lo=0.2, hi=0.8):
    rng = np.random.RandomState(seed)
    x = rng.uniform(lo, hi, (n, 3))
    v = 0.5 * rng.randn(n, 3)
    C_ = 5.0 * rng.randn(n, 3, 3)
    F = np.eye(3)[None] + 0.1 * rng.randn(n, 3, 3)
    bad = np.linalg.det(F) <= 0.2
    while bad.any():
        F[bad] = np.eye(3)[None] + 0.1 * rng.randn(int(bad.sum()), 3, 3)
        bad = np.linalg.det(F) <= 0.2
    return x, v, C_, F


def random_adjoint(n, seed):
    rng = np.random.RandomState(1000 + seed)
    return rng.randn(n, 3), rng.randn(n, 3), rng.randn(n, 3, 3), rng.randn(n, 3, 3)


def pose_array(states):
    out = np.zeros((len(states), 8))
    for k, s in enumerate(states):
        s = np.asarray(s, dtype=np.float64)
        out[k, :len(s)] = s
    return out


def oracle_prim_states(osim, poses):
    return [torch.as_tensor(poses[k, :p.state_dim].copy()) for k, p in enumerate(osim.prims)]


def c_setup(cfg, n, dtype='float64', **kw):
    descs = [_capi.primitive_desc(dict(p)) for p in cfg.PRIMITIVES]
    conf = _capi.make_config(dict(cfg.SIMULATOR), n, len(descs), dtype=dtype, **kw)
    arr = (_capi.PrimitiveDesc * max(len(descs), 1))(*descs)
    return conf, arr, descs


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def record(name, **vals):
    """Append measured parity errors to $PLB_PARITY_LOG (one JSON line per call): the tolerances in the GPU tests are set
    to about 3x what this log showed on a B200 (profiles/r2_parity_measured.md)."""
    import json
    import os
    path = os.environ.get("PLB_PARITY_LOG")
    if path:
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "a") as f:
            f.write(json.dumps(dict(test=name, **{k: float(v) for k, v in vals.items()})) + "\n")
