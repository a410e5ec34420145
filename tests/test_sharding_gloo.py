"""CPU, world_size 2, gloo: the slab decomposition protocol (partition, zone planes, partial-sum exchange) reproduces the
single-process P2G grid wherever a rank's particles look.  The per-rank scatter is the kernel body run by the host
emulation (tests/host/emul.cpp)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _p2g_grid(lib, cfg, x, v, Cm, F):
    import plb_test_helpers as H
    from plasticinelab_b200 import _capi
    D = _capi.dptr
    n = len(x)
    conf, parr, _ = H.c_setup(cfg, n, 'float64')
    G = conf.n_grid ** 3
    gin, gout = np.zeros((G, 4)), np.zeros((G, 4))
    xo, vo, Fo, Co = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
    pose = np.zeros((1, 8))
    lib.emul_substep_fwd(conf.dtype, C.byref(conf), parr, C.c_double(0.0), D(x), D(v), D(F), D(Cm), D(pose), D(pose), D(xo), D(vo), D(Fo), D(Co),
                         D(gin), D(gout))
    return gin.reshape(conf.n_grid, conf.n_grid, conf.n_grid, 4), gout.reshape(conf.n_grid, conf.n_grid, conf.n_grid, 4)


def _worker(rank, world, port, so_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import plb_test_helpers as H
    from plasticinelab_b200.engine import sharding
    lib = C.CDLL(so_path)
    n, n_grid, w = 1500, 32, 4
    cfg = H.small_cfg([], n_particles=n)
    x, v, Cm, F = H.random_state(n, 0, 0.2, 0.8)
    full_in, _ = _p2g_grid(lib, cfg, x, v, Cm, F)
    bounds = sharding.slab_bounds(x[:, 0], n_grid, world, w)
    idx = sharding.owned_index(x[:, 0], n_grid, bounds, rank)
    assert sharding.check_margin(x[idx, 0], n_grid, bounds, rank, w)
    mine, _ = _p2g_grid(lib, cfg, *(np.ascontiguousarray(a[idx]) for a in (x, v, Cm, F)))
    for side in (0, 1):
        z = sharding.zone(bounds, rank, side, w)
        if z is None:
            continue
        peer = rank - 1 if side == 0 else rank + 1
        send = torch.from_numpy(np.ascontiguousarray(mine[z[0]:z[1]]))
        recv = torch.zeros_like(send)
        if rank < peer:
            dist.send(send, peer); dist.recv(recv, peer)
        else:
            dist.recv(recv, peer); dist.send(send, peer)
        mine[z[0]:z[1]] += recv.numpy()
    # every node a local particle touches must now carry the global sum
    planes = sharding.base_plane(x[idx, 0], n_grid)
    lo, hi = planes.min(), planes.max() + 3
    touched = np.zeros((n_grid,) * 3, bool)
    b = (x[idx] * n_grid - 0.5).astype(int)
    for di in range(3):
        for dj in range(3):
            for dk in range(3):
                touched[b[:, 0] + di, b[:, 1] + dj, b[:, 2] + dk] = True
    err = np.abs(mine[touched] - full_in[touched]).max()
    assert err < 1e-14, (rank, err)
    # the two ranks' particle sets are disjoint and complete
    counts = torch.tensor([len(idx)])
    dist.all_reduce(counts)
    assert int(counts.item()) == n
    dist.barrier()
    dist.destroy_process_group()


def test_slab_protocol_world2_gloo(emul_lib):
    so_path = os.path.join(ROOT, 'tests', 'host', 'libplb_emul.so')
    mp.spawn(_worker, args=(2, 29533, so_path), nprocs=2, join=True)


def test_slab_bounds_properties():
    sys.path.insert(0, ROOT)
    from plasticinelab_b200.engine import sharding
    rng = np.random.RandomState(1)
    for world in (2, 4, 8):
        x = rng.uniform(0.05, 0.95, 200000)
        b = sharding.slab_bounds(x, 512, world, 8)
        assert b[0] == 0 and b[-1] == 512 and all(v % 4 == 0 for v in b)
        assert all(b[i + 1] - b[i] >= 16 for i in range(1, world - 1))
        sizes = [len(sharding.owned_index(x, 512, b, r)) for r in range(world)]
        assert sum(sizes) == len(x) and max(sizes) < 1.2 * len(x) / world
