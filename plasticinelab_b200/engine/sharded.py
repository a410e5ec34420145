"""One rank of a slab-decomposed env: `TaichiEnv` surface for the local particles + halo exchange over NCCL.

No reference counterpart (the reference is single-device).  Per substep the rank scatters its particles,
exchanges the partial sums of the boundary zones with its two neighbours (`torch.distributed` send/recv straight out of /
into engine memory), finishes the substep; the backward pass mirrors it with the adjoint of `grid_out`; loss scalars,
the contact minimum and the pose gradients are all-reduced.  See include/plb_b200.h ("multi-GPU slab decomposition").
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from .. import _capi
from . import sharding
from .shapes import Shapes
from .taichi_env import TaichiEnv


class _DevView:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (uint8 / float64, 1-D)."""

    def __init__(self, ptr, nbytes, typestr="|u1", itemsize=1):
        self.__cuda_array_interface__ = {"shape": (int(nbytes) // itemsize,), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _tensor(ptr, nbytes, device, f64=False):
    v = _DevView(ptr, nbytes, "<f8", 8) if f64 else _DevView(ptr, nbytes)
    return torch.as_tensor(v, device=device)


class ShardedEnv:
    def __init__(self, cfg, rank=None, world=None, dtype="float32", device=None, halo_w=8, group=None, peer=None, materials=None):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.group = group
        self.halo_w = halo_w
        dev_index = torch.cuda.current_device() if device is None else device
        self.device = torch.device("cuda", dev_index)
        k = _capi.sim_constants(dict(cfg.SIMULATOR))
        self.n_grid = k["n_grid"]
        full_x, _ = Shapes(cfg.SHAPES).get()
        self.n_global = len(full_x)
        self.bounds = sharding.slab_bounds(full_x[:, 0], self.n_grid, self.world, halo_w)
        self.index = sharding.owned_index(full_x[:, 0], self.n_grid, self.bounds, self.rank)
        assert len(self.index) > 0, f"rank {self.rank} owns no particles (bounds {self.bounds})"
        self.env = TaichiEnv(cfg, dtype=dtype, device=dev_index, particle_index=self.index)
        self.env.initialize()
        self.env.set_copy(False)
        if materials is not None:          # callable: positions (n, 3) of this rank's particles -> (mu, lam, yield_stress) arrays
            self.env.simulator.set_materials(*materials(self.env.init_particles))
        self.engine = eng = self.env.engine
        self.S = self.env.simulator.substeps
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.sides = [s for s in (0, 1) if sharding.zone(self.bounds, self.rank, s, halo_w) is not None]
        eng.call("plb_slab_configure", lo, hi, halo_w, int(0 in self.sides), int(1 in self.sides))
        self.buf = {}
        for which in range(3):
            for side in self.sides:
                for d in (0, 1):
                    ptr, nb = C.c_void_p(), C.c_longlong()
                    eng.call("plb_slab_buffer", which, side, d, C.byref(ptr), C.byref(nb))
                    self.buf[(which, side, d)] = _tensor(ptr.value, nb.value, self.device)
        ptr, nb = C.c_void_p(), C.c_longlong()
        eng.call("plb_device_buffer", 0, C.byref(ptr), C.byref(nb))
        self.acc = _tensor(ptr.value, nb.value, self.device, f64=True)
        eng.call("plb_device_buffer", 1, C.byref(ptr), C.byref(nb))
        self.prim_grad = _tensor(ptr.value, nb.value, self.device, f64=True)
        self.records = []
        self.cur = 0
        self.direct = False
        self.peer = (os.environ.get("PLB_SLAB_PEER", "1") != "0") if peer is None else bool(peer)
        if self.peer and self.world > 1:
            self._setup_peer()
        else:
            self.peer = False

    def _setup_peer(self):
        """Exchange the CUDA-IPC handles of the halo inboxes: afterwards the engine does the per-substep halo itself
        (P2P stores over NVLink, in-stream flags) and env steps run as CUDA graphs."""
        eng = self.engine
        mine = {}
        for side in self.sides:
            buf = C.create_string_buffer(64)
            eng.call("plb_slab_ipc_export", side, buf)
            mine[side] = buf.raw
        gathered = [None] * self.world
        dist.all_gather_object(gathered, mine, group=self.group)
        for side in self.sides:
            peer = self.rank - 1 if side == 0 else self.rank + 1
            h = gathered[peer][1 - side]          # the neighbour's inbox that faces me
            eng.call("plb_slab_ipc_import", side, C.create_string_buffer(h, 64))
        # direct halo (opt-in, PLB_SLAB_DIRECT=1): map the neighbours' scatter targets too (grid_in / adjoint of grid_out, even and
        # odd substeps).  Measured slower than the pushed zone blocks on 2 B200s (4.91e9 vs 5.84e9 particle-substeps/s at slab1m,
        # profiles/r2c_multi2_timeline.txt): with 8-plane zones 57 % of a 28-plane slab's scatter is duplicated as remote REDs.
        self.direct = os.environ.get("PLB_SLAB_DIRECT", "0") != "0" and os.environ.get("PLB_BWD_OVERLAP", "1") != "0"
        if self.direct:
            grids = []
            for which in range(4):
                buf = C.create_string_buffer(64)
                eng.call("plb_slab_ipc_export_grid", which, buf)
                grids.append(buf.raw)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, grids, group=self.group)
            for side in self.sides:
                peer = self.rank - 1 if side == 0 else self.rank + 1
                for which in range(4):
                    eng.call("plb_slab_ipc_import_grid", side, which, C.create_string_buffer(gathered[peer][which], 64))
        dist.barrier(group=self.group)

    def close(self):
        """Collective: unmap the neighbours' inboxes on every rank BEFORE any rank frees its own (CUDA IPC rule), then destroy."""
        if self.engine is None:
            return
        if self.peer:
            self.engine.call("plb_slab_ipc_close")
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        self.engine.close()
        self.engine = None

    # ---- communication
    def _exchange(self, which):
        if not self.sides:
            return
        ops = []
        for side in self.sides:
            peer = self.rank - 1 if side == 0 else self.rank + 1
            ops.append(dist.P2POp(dist.isend, self.buf[(which, side, 0)], peer, group=self.group))
            ops.append(dist.P2POp(dist.irecv, self.buf[(which, side, 1)], peer, group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def _allreduce_acc(self):
        if self.world == 1:
            return
        dist.all_reduce(self.acc[0:4], op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(self.acc[4:5], op=dist.ReduceOp.MAX, group=self.group)
        if self.env.loss.soft_contact_loss:      # soft minimum: both partial sums add up
            dist.all_reduce(self.acc[8:24], op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(self.acc[8:16], op=dist.ReduceOp.MIN, group=self.group)

    # ---- episode (trajectory mode), same call pattern as TaichiEnv under the tape
    def begin_episode(self, softness=666.0):
        env = self.env
        env.simulator.cur = 0
        for p in env.primitives:
            p.set_state(0, p.init_state)
        env.primitives.set_softness(softness)
        self.engine.call("plb_zero_grads")
        self.records = []
        self.cur = 0

    def step(self, action):
        eng, S, start = self.engine, self.S, self.cur
        self.env.primitives.set_action(start // S, S, action)
        eng.call("plb_kinematics", start, S)
        if self.peer:
            eng.call("plb_step_fwd", start, start, S)            # halo inside the engine (peer memory), one graph launch
        else:
            for s in range(start, start + S):
                eng.call("plb_slab_fwd_p2g", s, s + 1)
                self._exchange(0)
                eng.call("plb_slab_fwd_finish", s, s + 1, s)
        self.records.append(("step", start, S))
        self.cur = start + S

    def _loss_terms(self, f):
        self.engine.call("plb_slab_loss_begin", f)
        self._exchange(2)
        self.engine.call("plb_slab_loss_reduce", f, f)
        self._allreduce_acc()

    def compute_loss(self, sync=False):
        f = self.cur
        self._loss_terms(f)
        out = np.zeros(8) if sync else None
        self.engine.call("plb_slab_loss_finish", f, f, 0, _capi.dptr(out))
        self.records.append(("loss", f))
        return out

    def backward(self):
        eng = self.engine
        for rec in reversed(self.records):
            if rec[0] == "loss":
                self._loss_terms(rec[1])
                eng.call("plb_slab_loss_finish", rec[1], rec[1], 1, None)
            else:
                _, start, n = rec
                if self.peer:
                    eng.call("plb_step_bwd", start, start, n)
                else:
                    for s in reversed(range(start, start + n)):
                        eng.call("plb_slab_bwd_begin", s, s)
                        self._exchange(1)
                        eng.call("plb_slab_bwd_finish", s, s)
        if self.world > 1:
            dist.all_reduce(self.prim_grad, op=dist.ReduceOp.SUM, group=self.group)
        n_steps = sum(1 for r in self.records if r[0] == "step")
        return self.env.primitives.get_grad(n_steps, self.S)

    def loss_value(self):
        return self.env.loss.loss[None]

    def local_x(self, f):
        return self.env.simulator.get_x(f)

    def margin_ok(self, f):
        return sharding.check_margin(self.local_x(f)[:, 0], self.n_grid, self.bounds, self.rank, self.halo_w)
