"""CPU float64 ORACLE for the differentiable MPM hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, in plain torch/numpy float64 on the CPU, what the reference
(hzaskywalker/PlasticineLab @ ac1a2f7b) computes on the path SURVEY.md section 8
names.  It is the checker for the CUDA engine and never part of the product:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it.

What it follows (reference file:line):
  substep forward      plb/engine/mpm_simulator.py:82-90,124-141,157-184,189-242,245-257
  SVD adjoint          plb/engine/mpm_simulator.py:97-115,143-151   (literal formula + clamp)
  primitives           plb/engine/primitive/primive_base.py:57-121,184-192
                       plb/engine/primitive/primitives.py:8-257
                       plb/engine/primitive/utils.py:3-47
  loss                 plb/engine/losses/loss.py:81-106,116-162,186-254,269-298
  episode / tape       plb/optimizer/solver.py:31-44, plb/engine/taichi_env.py:78-106

Adjoints: the reference relies on Taichi 0.7.x compiler-generated reverse mode.
Here every adjoint is obtained from torch.autograd on the forward restatement,
ONE SUBSTEP AT A TIME (state at frame s + adjoint at frame s+1 -> adjoint at
frame s), which is exactly the structure of `substep_grad`
(mpm_simulator.py:260-278).  Three places deliberately do NOT use the true
derivative but the reference's behaviour:
  * SVD backward uses the reference formula with 1/clamp(s_j - s_i, +-1e-6).
  * max/min route the gradient by strict comparison, ties to the 2nd operand
    (Taichi autodiff of BinaryOp max/min).
  * the hard contact loss differentiates `ti.atomic_min` as if it were
    `atomic_add` (Taichi 0.7 MakeAdjoint::visit(AtomicOpStmt)): every particle
    with sdf>0 receives min_dist.grad.  `contact_grad='argmin'` switches to the
    mathematically true sub-gradient instead.

PARITY PIN STATUS: the Taichi runtime is a third-party dependency that is not
vendored in /root/reference and not installable here (taichi 0.7.14, commit
58feee37, per plb/optimizer/long_term_gradient.ipynb cell 1).  The forward path
is pinned to the only recorded output of the reference (Move-v1 summed loss,
notebook cell 3; see tests/test_oracle_anchor.py).  Gradients are pinned by
central finite differences of this oracle's own forward in float64 -- at the
Taichi boundary they are "parity unpinned".

`ti.svd`: Taichi's 3x3 SVD (McAdams/Sifakis) returns det(U)=det(V)=+1 with the
sign carried by the smallest singular value.  The oracle calls LAPACK and then
moves the signs to that convention; U V^T and U f(S) V^T are invariant to the
remaining gauge freedom whenever singular values are distinct.
"""
from __future__ import annotations

import math

import numpy as np
import torch

DT = torch.float64


# --------------------------------------------------------------------------------------
# Taichi-flavoured primitives of the autodiff
# --------------------------------------------------------------------------------------
def tmax(a, b):
    """ti.max(a, b): value max; gradient to a iff b < a, else to b."""
    a, b = torch.broadcast_tensors(torch.as_tensor(a, dtype=DT), torch.as_tensor(b, dtype=DT))
    return torch.where(b < a, a, b)


def tmin(a, b):
    """ti.min(a, b): value min; gradient to a iff a < b, else to b."""
    a, b = torch.broadcast_tensors(torch.as_tensor(a, dtype=DT), torch.as_tensor(b, dtype=DT))
    return torch.where(a < b, a, b)


def _length(x, eps):
    return torch.sqrt((x * x).sum(-1) + eps)


def qrot(rot, v):
    """plb/engine/primitive/utils.py:7-13"""
    qvec = rot[..., 1:4].expand(v.shape[:-1] + (3,))
    uv = torch.cross(qvec, v, dim=-1)
    uuv = torch.cross(qvec, uv, dim=-1)
    return v + 2 * (rot[..., 0:1] * uv + uuv)


def qmul(q, r):
    """plb/engine/primitive/utils.py:19-27 (terms = r outer q), renormalised."""
    t = r[:, None] * q[None, :]
    w = t[0, 0] - t[1, 1] - t[2, 2] - t[3, 3]
    x = t[0, 1] + t[1, 0] - t[2, 3] + t[3, 2]
    y = t[0, 2] + t[1, 3] + t[2, 0] - t[3, 1]
    z = t[0, 3] - t[1, 2] + t[2, 1] + t[3, 0]
    out = torch.stack([w, x, y, z])
    return out / torch.sqrt((out * out).sum())


def w2quat(w):
    """plb/engine/primitive/utils.py:29-41.  For |w| <= 1e-9 the identity is returned and no
    gradient reaches w (the oracle defines the Taichi 0*inf case as 0)."""
    n2 = float((w * w).sum().detach())
    if math.sqrt(n2) > 1e-9:
        n = torch.sqrt((w * w).sum())
        v = (w / n) * torch.sin(n / 2)
        return torch.cat([torch.cos(n / 2)[None], v])
    return torch.tensor([1.0, 0.0, 0.0, 0.0], dtype=DT) + 0.0 * w.sum().detach()


def inv_trans(pos, position, rotation):
    """plb/engine/primitive/utils.py:43-47"""
    inv = torch.stack([rotation[0], -rotation[1], -rotation[2], -rotation[3]])
    inv = inv / torch.sqrt((inv * inv).sum())
    return qrot(inv, pos - position)


# --------------------------------------------------------------------------------------
# SVD with the reference's backward
# --------------------------------------------------------------------------------------
# UNVERIFIED Taichi behaviour (SURVEY.md 8c caveat 1): `ti.svd(A)` without `dt` may run its float32 routine on the float64 data.
# SVD_INTERNAL_F32 = True makes the oracle decompose in float32 and cast back, which is how large the effect of that would be
# (tests/test_oracle.py::test_f32_internal_svd_sensitivity); the default is the exact float64 decomposition.
SVD_INTERNAL_F32 = False


def svd_sifakis_convention(A: torch.Tensor):
    """LAPACK SVD moved to det(U)=det(V)=+1, |sigma| descending, sign on the last one."""
    if SVD_INTERNAL_F32 and A.dtype == torch.float64:
        U, S, Vh = torch.linalg.svd(A.float())
        U, S, Vh = U.double(), S.double(), Vh.double()
    else:
        U, S, Vh = torch.linalg.svd(A)
    V = Vh.transpose(-1, -2).clone()
    U = U.clone()
    S = S.clone()
    du = torch.linalg.det(U) < 0
    U[du, :, 2] = -U[du, :, 2]
    S[du, 2] = -S[du, 2]
    dv = torch.linalg.det(V) < 0
    V[dv, :, 2] = -V[dv, :, 2]
    S[dv, 2] = -S[dv, 2]
    return U, S, V


def _clamp_ref(a):
    """mpm_simulator.py:143-151"""
    return torch.where(a >= 0, torch.clamp(a, min=1e-6), torch.clamp(a, max=-1e-6))


class RefSVD(torch.autograd.Function):
    """U, sig (as a diagonal matrix, like ti.svd), V with `backward_svd` (mpm_simulator.py:97-115)."""

    @staticmethod
    def forward(ctx, A):
        U, S, V = svd_sifakis_convention(A)
        sig = torch.diag_embed(S)
        ctx.save_for_backward(U, sig, V)
        return U, sig, V

    @staticmethod
    def backward(ctx, gu, gsigma, gv):
        u, sig, v = ctx.saved_tensors
        vt = v.transpose(-1, -2)
        ut = u.transpose(-1, -2)
        sigma_term = u @ gsigma @ vt
        s = torch.diagonal(sig, dim1=-2, dim2=-1) ** 2
        diff = s[:, None, :] - s[:, :, None]          # [i, j] = s[j] - s[i]
        Fm = 1.0 / _clamp_ref(diff)
        eye = torch.eye(3, dtype=torch.bool)
        Fm = torch.where(eye, torch.zeros_like(Fm), Fm)
        u_term = u @ ((Fm * (ut @ gu - gu.transpose(-1, -2) @ u)) @ sig) @ vt
        v_term = u @ (sig @ ((Fm * (vt @ gv - gv.transpose(-1, -2) @ v)) @ vt))
        return u_term + v_term + sigma_term


# --------------------------------------------------------------------------------------
# Rigid primitives
# --------------------------------------------------------------------------------------
def _plen(x):
    """primitives.py:8-10 (eps 1e-14)"""
    return _length(x, 1e-14)


def _pnormalize(x):
    return x / _plen(x)[..., None]


class Prim:
    """One rigid manipulator.  Kinematic state per frame: position(3), rotation(4) [, gap]."""
    state_dim = 7

    def __init__(self, cfg: dict):
        self.cfg = cfg
        self.shape = cfg['shape']
        self.friction = float(cfg.get('friction', 0.9))
        self.init_pos = tuple(float(v) for v in cfg.get('init_pos', (0.3, 0.3, 0.3)))
        self.init_rot = tuple(float(v) for v in cfg.get('init_rot', (1.0, 0.0, 0.0, 0.0)))
        self.lower = torch.tensor([float(v) for v in cfg.get('lower_bound', (0.0, 0.0, 0.0))], dtype=DT)
        self.upper = torch.tensor([float(v) for v in cfg.get('upper_bound', (1.0, 1.0, 1.0))], dtype=DT)
        action = cfg.get('action', None) or {}
        self.action_dim = int(action.get('dim', 0) or 0)
        self.action_scale = torch.tensor([float(v) for v in action.get('scale', ())], dtype=DT)
        self.softness = 0.0

    # ---- kinematic state helpers: state = tensor(7 or 8)
    def init_state(self):
        return torch.tensor(self.init_pos + self.init_rot, dtype=DT)

    def velocities(self, act):
        """set_velocity (primive_base.py:184-192): per-substep v (3), w (3) [, gap_vel]; caller divides by S."""
        v = act[0:3] * self.action_scale[0:3]
        if self.action_dim > 3:
            w = act[3:6] * self.action_scale[3:6]
        else:
            w = torch.zeros(3, dtype=DT)
        return v, w, None

    def fk(self, st, v, w, gv):
        """forward_kinematics (primive_base.py:117-121): base class left-multiplies the rotation."""
        pos = tmax(tmin(st[0:3] + v, self.upper), self.lower)
        rot = qmul(w2quat(w), st[3:7])
        return torch.cat([pos, rot])

    # ---- geometry (world-space query points p: (M,3); state st: (7|8,))
    def sdf(self, st, p):
        return self._sdf(st, inv_trans(p, st[0:3], st[3:7]))

    def normal(self, st, p):
        return qrot(st[3:7], self._normal(st, inv_trans(p, st[0:3], st[3:7])))

    def collider_v(self, st, st1, p, dt):
        """primive_base.py:82-89"""
        rel = inv_trans(p, st[0:3], st[3:7])
        new_pos = qrot(st1[3:7], rel) + st1[0:3]
        return (new_pos - p) / dt

    def collide(self, st, st1, p, v_out, dt):
        """primive_base.py:91-115, evaluated only on the rows where the branch is taken."""
        with torch.no_grad():
            dist0 = self.sdf(st, p)
            infl0 = torch.clamp(torch.exp(-dist0 * self.softness), max=1.0)
            take = ((self.softness > 0) & (infl0 > 0.1)) | (dist0 <= 0)
        if not bool(take.any()):
            return v_out
        idx = take.nonzero()[:, 0]
        ps = p[idx]
        vo = v_out[idx]
        dist = self.sdf(st, ps)
        influence = tmin(torch.exp(-dist * self.softness), 1.0)
        D = self.normal(st, ps)
        cv = self.collider_v(st, st1, ps, dt)
        iv = vo - cv
        nc = (iv * D).sum(-1)
        vt = iv - tmin(nc, 0.0)[:, None] * D
        vtn = _length(vt, 1e-8)                       # primitive/utils.py:3-5
        vtf = vt / vtn[:, None] * tmax(0.0, vtn + nc * self.friction)[:, None]
        with torch.no_grad():
            flag = ((nc < 0) & (torch.sqrt((vt * vt).sum(-1)) > 1e-30)).to(DT)[:, None]
        vt = vtf * flag + vt * (1 - flag)
        new = cv + iv * (1 - influence)[:, None] + vt * influence[:, None]
        return v_out.index_copy(0, idx, new)


class Sphere(Prim):
    """primitives.py:17-34: sdf/normal in world frame, rotation ignored."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.radius = float(cfg.get('radius', 1.0))

    def sdf(self, st, p):
        return _plen(p - st[0:3]) - self.radius

    def normal(self, st, p):
        return _pnormalize(p - st[0:3])


class Capsule(Prim):
    """primitives.py:36-60"""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.h = float(cfg.get('h', 0.06))
        self.r = float(cfg.get('r', 0.03))

    def _p2(self, q):
        y = q[:, 1] + self.h / 2
        y = y - tmin(tmax(y, 0.0), self.h)
        return torch.stack([q[:, 0], y, q[:, 2]], -1)

    def _sdf(self, st, q):
        return _plen(self._p2(q)) - self.r

    def _normal(self, st, q):
        return _pnormalize(self._p2(q))


class RollingPin(Capsule):
    """primitives.py:63-80"""

    def fk(self, st, v, w, gv):
        dw, dth, dy = v[0], v[1], v[2]
        rot = st[3:7]
        y_dir = qrot(rot, torch.tensor([[0.0, -1.0, 0.0]], dtype=DT))[0]
        x_dir = torch.cross(torch.tensor([0.0, 1.0, 0.0], dtype=DT), y_dir, dim=0) * dw * 0.03
        x_dir = torch.stack([x_dir[0], dy, x_dir[2]])
        z = torch.zeros((), dtype=DT)
        rot1 = qmul(w2quat(torch.stack([z, -dth, z])), qmul(rot, w2quat(torch.stack([z, dw, z]))))
        pos = tmax(tmin(st[0:3] + x_dir, self.upper), self.lower)
        return torch.cat([pos, rot1])


class Chopsticks(Capsule):
    """primitives.py:83-155 (state has an 8th entry: the gap)."""
    state_dim = 8

    def __init__(self, cfg):
        super().__init__(cfg)
        self.minimal_gap = float(cfg.get('minimal_gap', 0.06))
        self.init_gap = float(cfg.get('init_gap', 0.06))
        assert self.action_dim == 7

    def init_state(self):
        return torch.tensor(self.init_pos + self.init_rot + (self.init_gap,), dtype=DT)

    def velocities(self, act):
        v = act[0:3] * self.action_scale[0:3]
        w = act[3:6] * self.action_scale[3:6]
        return v, w, act[6] * self.action_scale[6]

    def fk(self, st, v, w, gv):
        gap = tmax(st[7] - gv, self.minimal_gap)
        pos = tmax(tmin(st[0:3] + v, self.upper), self.lower)
        rot = qmul(st[3:7], w2quat(w))                # right-multiply (primitives.py:96)
        return torch.cat([pos, rot, gap[None]])

    def _ab(self, st, q):
        delta = torch.stack([st[7] / 2, torch.zeros((), dtype=DT), torch.zeros((), dtype=DT)])
        p = q - torch.tensor([0.0, -self.h / 2, 0.0], dtype=DT)
        return p - delta, p + delta

    def _sdf(self, st, q):
        pa, pb = self._ab(st, q)
        return tmin(Capsule._sdf(self, st, pa), Capsule._sdf(self, st, pb))

    def _normal(self, st, q):
        pa, pb = self._ab(st, q)
        with torch.no_grad():
            m = (Capsule._sdf(self, st, pa) <= Capsule._sdf(self, st, pb)).to(DT)[:, None]
        return m * Capsule._normal(self, st, pa) + (1 - m) * Capsule._normal(self, st, pb)


class Cylinder(Prim):
    """primitives.py:158-190 (note: cfg.h is the radial half-extent, cfg.r the axial one)."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.h = float(cfg.get('h', 0.2))
        self.r = float(cfg.get('r', 0.1))

    def _sdf(self, st, q):
        l = _plen(torch.stack([q[:, 0], q[:, 2]], -1))
        d = torch.abs(torch.stack([l, q[:, 1]], -1)) - torch.tensor([self.h, self.r], dtype=DT)
        return tmin(tmax(d[:, 0], d[:, 1]), 0.0) + _plen(tmax(d, 0.0))

    def _normal(self, st, q):
        p = torch.stack([q[:, 0], q[:, 2]], -1)
        l = _plen(p)
        d = torch.stack([l, torch.abs(q[:, 1])], -1) - torch.tensor([self.h, self.r], dtype=DT)
        with torch.no_grad():
            f = (d[:, 0] > d[:, 1]).to(DT)
            inside = (torch.maximum(d[:, 0], d[:, 1]) <= 0.0).to(DT)
            sgn = (q[:, 1] >= 0).to(DT) * 2 - 1
        n2 = tmax(d, 0.0) + inside[:, None] * torch.stack([f, 1 - f], -1)
        n2_ = n2 / _plen(n2)[:, None]
        p2 = p / l[:, None]
        n3 = torch.stack([p2[:, 0] * n2_[:, 0], n2_[:, 1] * sgn, p2[:, 1] * n2_[:, 0]], -1)
        return _pnormalize(n3)


class Torus(Prim):
    """primitives.py:193-221"""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.tx = float(cfg.get('tx', 0.2))
        self.ty = float(cfg.get('ty', 0.1))

    def _sdf(self, st, q):
        l = _plen(torch.stack([q[:, 0], q[:, 2]], -1))
        return _plen(torch.stack([l - self.tx, q[:, 1]], -1)) - self.ty

    def _normal(self, st, q):
        x = torch.stack([q[:, 0], q[:, 2]], -1)
        l = _plen(x)
        qq = torch.stack([l - self.tx, q[:, 1]], -1)
        n2 = qq / _plen(qq)[:, None]
        x2 = x / l[:, None]
        n3 = torch.stack([x2[:, 0] * n2[:, 0], n2[:, 1], x2[:, 1] * n2[:, 0]], -1)
        return _pnormalize(n3)


class Box(Prim):
    """primitives.py:224-256 (central-difference normal, d = 1e-4)."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.size = torch.tensor([float(v) for v in cfg.get('size', (0.1, 0.1, 0.1))], dtype=DT)

    def _sdf(self, st, q):
        d = torch.abs(q) - self.size
        out = _plen(tmax(d, 0.0))
        return out + tmin(tmax(d[:, 0], tmax(d[:, 1], d[:, 2])), 0.0)

    def _normal(self, st, q):
        d = 1e-4
        cols = []
        for i in range(3):
            e = torch.zeros(3, dtype=DT)
            e[i] = d
            cols.append((0.5 / d) * (self._sdf(st, q + e) - self._sdf(st, q - e)))
        n = torch.stack(cols, -1)
        return n / _length(n, 1e-14)[:, None]


PRIM_TYPES = dict(Sphere=Sphere, Capsule=Capsule, RollingPin=RollingPin, Chopsticks=Chopsticks,
                  Cylinder=Cylinder, Torus=Torus, Box=Box)


# --------------------------------------------------------------------------------------
# Simulator
# --------------------------------------------------------------------------------------
class OracleSim:
    """State is a tuple of torch float64 tensors (x[N,3], v[N,3], C[N,3,3], F[N,3,3])."""

    def __init__(self, sim_cfg: dict, prim_cfgs=(), n_particles=None):
        c = sim_cfg
        assert int(c.get('dim', 3)) == 3
        quality = float(c.get('quality', 1)) * 0.5
        self.n_grid = int(128 * quality)
        self.dx, self.inv_dx = 1.0 / self.n_grid, float(self.n_grid)
        self.dt = 0.5e-4 / quality
        self.p_vol = (self.dx * 0.5) ** 2
        self.p_mass = self.p_vol * 1
        E, nu = float(c.get('E', 5e3)), float(c.get('nu', 0.2))
        self.mu0, self.lam0 = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
        self.yield0 = float(c.get('yield_stress', 50.0))
        self.ground_friction = float(c.get('ground_friction', 1.5))
        g = c.get('gravity', (0, -1, 0))
        g = eval(g) if isinstance(g, str) else g
        self.gravity = torch.tensor([float(v) for v in g], dtype=DT)
        self.substeps = int(2e-3 // self.dt)
        self.prims = [PRIM_TYPES[p['shape']](p) for p in prim_cfgs]
        self.n_particles = n_particles
        self.mu = self.lam = self.yield_stress = None      # per-particle, set by set_materials

        offs = torch.tensor([[i, j, k] for i in range(3) for j in range(3) for k in range(3)])
        self.offs = offs                                   # (27,3) int64, same order as ti.ndrange(3,3,3)
        self.offs_f = offs.to(DT)

    def set_materials(self, n, mu=None, lam=None, yield_stress=None):
        self.n_particles = n
        self.mu = torch.full((n,), self.mu0, dtype=DT) if mu is None else torch.as_tensor(mu, dtype=DT)
        self.lam = torch.full((n,), self.lam0, dtype=DT) if lam is None else torch.as_tensor(lam, dtype=DT)
        self.yield_stress = (torch.full((n,), self.yield0, dtype=DT) if yield_stress is None
                             else torch.as_tensor(yield_stress, dtype=DT))

    def set_softness(self, s):
        for p in self.prims:
            p.softness = float(s)

    # ---------------- stencil helpers
    def _stencil(self, x):
        base = (x * self.inv_dx - 0.5).detach().to(torch.int64)       # C-style truncation (cast(int))
        fx = x * self.inv_dx - base.to(DT)
        w = torch.stack([0.5 * (1.5 - fx) ** 2, 0.75 - (fx - 1.0) ** 2, 0.5 * (fx - 0.5) ** 2], 1)  # (N,3off,3dim)
        o = self.offs
        weight = w[:, o[:, 0], 0] * w[:, o[:, 1], 1] * w[:, o[:, 2], 2]                              # (N,27)
        node = base[:, None, :] + o[None, :, :]                                                     # (N,27,3)
        n = self.n_grid
        lin = (node[..., 0] * n + node[..., 1]) * n + node[..., 2]
        return base, fx, weight, lin

    # ---------------- forward pieces (mpm_simulator.py:82-242)
    def p2g(self, x, v, C, F):
        N = x.shape[0]
        I3 = torch.eye(3, dtype=DT)
        F_tmp = (I3 + self.dt * C) @ F
        U, sig, V = RefSVD.apply(F_tmp)
        # compute_von_mises (:124-141)
        sd = torch.diagonal(sig, dim1=-2, dim2=-1)
        sc = tmax(sd, 0.05)
        eps = torch.log(sc)
        eps_hat = eps - eps.sum(-1, keepdim=True) / 3
        eps_norm = torch.sqrt((eps_hat * eps_hat).sum(-1) + 1e-8)
        dgamma = eps_norm - self.yield_stress / (2 * self.mu)
        with torch.no_grad():
            yidx = (dgamma > 0).nonzero()[:, 0]          # the yield branch is evaluated on yielding rows only
        new_F = F_tmp
        if len(yidx) > 0:
            eps_new = eps[yidx] - (dgamma[yidx] / eps_norm[yidx])[:, None] * eps_hat[yidx]
            F_yield = U[yidx] @ torch.diag_embed(torch.exp(eps_new)) @ V[yidx].transpose(-1, -2)
            new_F = F_tmp.index_copy(0, yidx, F_yield)
        # stress (:166-174)
        J = torch.linalg.det(new_F)
        r = U @ V.transpose(-1, -2)
        stress = 2 * self.mu[:, None, None] * (new_F - r) @ new_F.transpose(-1, -2) \
            + I3 * (self.lam * J * (J - 1))[:, None, None]
        stress = (-self.dt * self.p_vol * 4 * self.inv_dx * self.inv_dx) * stress
        affine = stress + self.p_mass * C
        base, fx, weight, lin = self._stencil(x)
        dpos = (self.offs_f[None] - fx[:, None, :]) * self.dx                      # (N,27,3)
        mv = weight[..., None] * (self.p_mass * v[:, None, :] + torch.einsum('nij,nkj->nki', affine, dpos))
        G = self.n_grid ** 3
        grid_v_in = torch.zeros(G, 3, dtype=DT).index_add(0, lin.reshape(-1), mv.reshape(-1, 3))
        grid_m = torch.zeros(G, dtype=DT).index_add(0, lin.reshape(-1), (weight * self.p_mass).reshape(-1))
        return new_F, grid_v_in, grid_m

    def grid_op(self, grid_v_in, grid_m, prim_f, prim_f1):
        """:189-221 on the active nodes only (inactive nodes keep v_out = 0)."""
        n = self.n_grid
        with torch.no_grad():
            idx = (grid_m > 1e-12).nonzero()[:, 0]
        m = grid_m[idx]
        I = torch.stack([idx // (n * n), (idx // n) % n, idx % n], -1)
        If = I.to(DT)
        v_out = (1 / m)[:, None] * grid_v_in[idx]
        v_out = v_out + self.dt * self.gravity * 30
        pos = If * self.dx
        for k, prim in enumerate(self.prims):
            v_out = prim.collide(prim_f[k], prim_f1[k], pos, v_out, self.dt)
        bound = 3
        gf = self.ground_friction
        for d in range(3):
            c1 = (I[:, d] < bound) & (v_out[:, d].detach() < 0)
            if d != 1 or gf == 0:
                col = torch.where(c1, torch.zeros_like(v_out[:, d]), v_out[:, d])
                v_out = torch.cat([v_out[:, :d], col[:, None], v_out[:, d + 1:]], 1)
            elif gf < 10:
                normal = torch.zeros(3, dtype=DT)
                normal[d] = 1.0
                lin_ = (v_out * normal).sum(-1) + 1e-30
                vit = v_out - lin_[:, None] * normal - If * 1e-30
                lit = torch.sqrt((vit * vit).sum(-1) + 1e-8)
                vn = tmax(1.0 + gf * lin_ / lit, 0.0)[:, None] * (vit + If * 1e-30)
                vn = torch.stack([vn[:, 0], torch.zeros_like(vn[:, 1]), vn[:, 2]], -1)
                v_out = torch.where(c1[:, None], vn, v_out)
            else:
                v_out = torch.where(c1[:, None], torch.zeros_like(v_out), v_out)
            c2 = (I[:, d] > n - bound) & (v_out[:, d].detach() > 0)
            col = torch.where(c2, torch.zeros_like(v_out[:, d]), v_out[:, d])
            v_out = torch.cat([v_out[:, :d], col[:, None], v_out[:, d + 1:]], 1)
        return torch.zeros(n ** 3, 3, dtype=DT).index_copy(0, idx, v_out)

    def g2p(self, x, grid_v_out):
        base, fx, weight, lin = self._stencil(x)
        dpos = self.offs_f[None] - fx[:, None, :]
        gv = grid_v_out[lin]                                                        # (N,27,3)
        new_v = (weight[..., None] * gv).sum(1)
        new_C = 4 * self.inv_dx * torch.einsum('nk,nki,nkj->nij', weight, gv, dpos)
        new_x = tmax(tmin(x + self.dt * new_v, 1.0 - 3 * self.dx), 0.0)
        return new_x, new_v, new_C

    def substep(self, state, prim_f, prim_f1, return_grid=False):
        x, v, C, F = state
        new_F, gvi, gm = self.p2g(x, v, C, F)
        gvo = self.grid_op(gvi, gm, prim_f, prim_f1)
        nx, nv, nC = self.g2p(x, gvo)
        if return_grid:
            return (nx, nv, nC, new_F), (gvi, gm, gvo)
        return nx, nv, nC, new_F

    def substep_vjp(self, state, prim_f, prim_f1, adj_next):
        """Adjoint of one substep: returns (adj_state[s], [gprim_f], [gprim_f1])."""
        leaves = [t.detach().clone().requires_grad_(True) for t in state]
        pf = [t.detach().clone().requires_grad_(True) for t in prim_f]
        pf1 = [t.detach().clone().requires_grad_(True) for t in prim_f1]
        out = self.substep(tuple(leaves), pf, pf1)
        inputs = leaves + pf + pf1
        grads = torch.autograd.grad(out, inputs, grad_outputs=list(adj_next), allow_unused=True)
        grads = [torch.zeros_like(i) if g is None else g for g, i in zip(grads, inputs)]
        P = len(pf)
        return tuple(grads[:4]), grads[4:4 + P], grads[4 + P:]

    # ---------------- loss (losses/loss.py)
    def grid_mass(self, x):
        """compute_grid_m_kernel, mpm_simulator.py:382-392"""
        base, fx, weight, lin = self._stencil(x)
        return torch.zeros(self.n_grid ** 3, dtype=DT).index_add(0, lin.reshape(-1), (weight * self.p_mass).reshape(-1))


class OracleLoss:
    def __init__(self, sim: OracleSim, target_density: np.ndarray, weights=(10.0, 10.0, 1.0), soft_contact=False,
                 contact_grad='taichi', target_sdf: np.ndarray | None = None):
        self.sim = sim
        self.sdf_w, self.density_w, self.contact_w = [float(w) for w in weights]
        self.soft_contact = bool(soft_contact)
        self.contact_grad = contact_grad
        self.movable = [k for k, p in enumerate(sim.prims) if p.action_dim > 0]      # loss.py:20-24
        self.target_density = torch.as_tensor(np.asarray(target_density, dtype=np.float64)).reshape(-1)
        if target_sdf is None:
            target_sdf = build_target_sdf(np.asarray(target_density, dtype=np.float64), sim.dx)
        self.target_sdf = torch.as_tensor(target_sdf).reshape(-1)
        with torch.no_grad():
            self.target_iou = float(self.iou(self.target_density))

    def iou(self, grid_m):
        """iou_kernel loss.py:239-254"""
        ma, mb = grid_m.max(), self.target_density.max()
        ma = torch.clamp(ma, min=0.0)
        mb = torch.clamp(mb, min=0.0)
        I = (grid_m * self.target_density).sum() / ma / mb
        U = grid_m.sum() / ma + self.target_density.sum() / mb
        return I / (U - I)

    def terms(self, x, prim_f, surrogate=False):
        """(density, sdf, contact, grid_m).  With surrogate=True `contact` is a scalar whose autograd
        gradient equals what the reference propagates (see module docstring); its value is meaningless."""
        sim = self.sim
        gm = sim.grid_mass(x)
        density = torch.abs(gm - self.target_density).sum()
        sdf = (self.target_sdf * gm).sum()
        contact = torch.zeros((), dtype=DT)
        for k in self.movable:
            d = tmax(sim.prims[k].sdf(prim_f[k], x), 0.0)
            if self.soft_contact:
                sw = 1 / (1 + d * d * 10000)
                md = (d * sw / sw.sum()).sum()
                contact = contact + md ** 2
            else:
                md_val = torch.clamp(d.detach().min(), max=100000.0)
                if not surrogate:
                    contact = contact + md_val ** 2
                elif self.contact_grad == 'taichi':
                    contact = contact + (2 * md_val) * tmax(d, 0.0).sum()
                else:
                    contact = contact + (2 * md_val) * d[torch.argmin(d.detach())]
        return density, sdf, contact, gm

    def value(self, x, prim_f):
        with torch.no_grad():
            d, s, c, gm = self.terms(x, prim_f)
            total = self.contact_w * c + self.density_w * d + self.sdf_w * s
            return dict(loss=float(total), contact_loss=float(c), density_loss=float(d), sdf_loss=float(s),
                        iou=float(self.iou(gm)), target_iou=self.target_iou)

    def vjp(self, x, prim_f):
        """d(total step loss)/d(x, prim_f) with seed 1 (Tape sets loss.grad = 1)."""
        xl = x.detach().clone().requires_grad_(True)
        pf = [t.detach().clone().requires_grad_(True) for t in prim_f]
        d, s, c, _ = self.terms(xl, pf, surrogate=not self.soft_contact)
        total = self.contact_w * c + self.density_w * d + self.sdf_w * s
        grads = torch.autograd.grad(total, [xl] + pf, allow_unused=True)
        grads = [torch.zeros_like(i) if g is None else g for g, i in zip(grads, [xl] + pf)]
        return grads[0], grads[1:]


def build_target_sdf_c(density: np.ndarray, dx: float) -> np.ndarray:
    """Same sweep in plain C (oracle/oracle_c.c, built by __graft_entry__.build_oracle); bit-identical to the numpy
    statement below (tests/test_oracle.py) and ~50x faster."""
    import ctypes
    import os
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_build', 'liboracle_c.so')
    lib = ctypes.CDLL(so)
    d = np.ascontiguousarray(density, dtype=np.float64)
    out = np.zeros_like(d)
    rc = lib.oracle_build_target_sdf(int(d.shape[0]), ctypes.c_double(dx), d.ctypes.data_as(ctypes.c_void_p),
                                     out.ctypes.data_as(ctypes.c_void_p))
    assert rc > 0
    return out


def build_target_sdf(density: np.ndarray, dx: float, inf: float = 1000.0) -> np.ndarray:
    """Loss.update_target (loss.py:81-106): 2*n_grid Jacobi sweeps, each node looking at the 6^3-1
    neighbours with offsets in [-3, 3) in lexicographic order, strict '<' updates.  The sweep is a pure
    function of the previous sweep's (sdf, nearest_point) pair, so iteration stops early at a fixed point."""
    n = density.shape[0]
    ii = np.arange(n)
    gp = np.stack(np.meshgrid(ii, ii, ii, indexing='ij'), -1).astype(np.float64) * dx
    inside = density > 1e-4
    sdf_copy = np.full((n, n, n), inf)
    near_copy = np.zeros((n, n, n, 3))
    offsets = [(a, b, c) for a in range(-3, 3) for b in range(-3, 3) for c in range(-3, 3) if (a, b, c) != (0, 0, 0)]
    for _ in range(2 * n):
        sdf = np.full((n, n, n), inf)
        near = near_copy.copy()          # nearest_point[I] persists unless overwritten (field semantics)
        for (a, b, c) in offsets:
            sl_dst = tuple(slice(max(0, -o), n - max(0, o)) for o in (a, b, c))
            sl_src = tuple(slice(max(0, o), n - max(0, -o)) for o in (a, b, c))
            src_sdf = sdf_copy[sl_src]
            src_near = near_copy[sl_src]
            diff = gp[sl_dst] - src_near
            dist = np.sqrt((diff * diff).sum(-1) + 1e-8)
            upd = (src_sdf < inf) & (dist < sdf[sl_dst]) & (~inside[sl_dst])
            sdf[sl_dst] = np.where(upd, dist, sdf[sl_dst])
            near[sl_dst] = np.where(upd[..., None], src_near, near[sl_dst])
        sdf[inside] = 0.0
        near[inside] = gp[inside]
        if np.array_equal(sdf, sdf_copy) and np.array_equal(near, near_copy):
            break
        sdf_copy, near_copy = sdf, near
    return sdf_copy


# --------------------------------------------------------------------------------------
# Episode driver = TaichiEnv.step / compute_loss under ti.Tape (solver.py:31-44)
# --------------------------------------------------------------------------------------
class OracleEnv:
    def __init__(self, cfg: dict, init_particles: np.ndarray, target_density: np.ndarray | None = None,
                 loss_weights=(10.0, 10.0, 1.0), soft_contact=False, contact_grad='taichi', target_sdf=None,
                 materials=None):
        self.sim = OracleSim(dict(cfg['SIMULATOR']), [dict(p) for p in cfg['PRIMITIVES']])
        n = len(init_particles)
        self.sim.set_materials(n, **(materials or {}))
        self.n = n
        self.init_x = torch.as_tensor(np.asarray(init_particles, dtype=np.float64))
        self.loss = None
        if target_density is not None:
            self.loss = OracleLoss(self.sim, target_density, loss_weights, soft_contact, contact_grad, target_sdf)
        self.action_dims = [0]
        for p in self.sim.prims:
            self.action_dims.append(self.action_dims[-1] + p.action_dim)

    def initial_state(self):
        n = self.n
        I3 = torch.eye(3, dtype=DT).expand(n, 3, 3).clone()
        return (self.init_x.clone(), torch.zeros(n, 3, dtype=DT), torch.zeros(n, 3, 3, dtype=DT), I3)

    def initial_prims(self):
        return [p.init_state() for p in self.sim.prims]

    def trajectory(self, prim0, actions: torch.Tensor):
        """Primitive frames 0..T*S as a differentiable function of the (clipped) actions."""
        S = self.sim.substeps
        frames = [list(prim0)]
        for a in actions:
            a = torch.clamp(a.detach(), -1, 1) + (a - a.detach())      # host-side clip, no_grad_set_action
            for _ in range(S):
                cur = frames[-1]
                nxt = []
                for k, p in enumerate(self.sim.prims):
                    if p.action_dim > 0:
                        v, w, gv = p.velocities(a[self.action_dims[k]:self.action_dims[k + 1]])
                        v, w = v / S, w / S
                        gv = None if gv is None else gv / S
                    else:
                        v, w, gv = torch.zeros(3, dtype=DT), torch.zeros(3, dtype=DT), torch.zeros((), dtype=DT)
                    nxt.append(p.fk(cur[k], v, w, gv))
                frames.append(nxt)
        return frames

    def rollout(self, actions, softness=666.0, state=None, prim0=None, with_grad=True, keep_states=False,
                loss_every_step=True):
        """Forward T env steps (+ loss after each) and, optionally, the action gradient.
        Returns dict(loss, per_step, grad[T,A], final_state, states?)."""
        sim = self.sim
        sim.set_softness(softness)
        S = sim.substeps
        acts = torch.as_tensor(np.asarray(actions, dtype=np.float64)).clone().requires_grad_(with_grad)
        prim0 = self.initial_prims() if prim0 is None else [torch.as_tensor(p, dtype=DT) for p in prim0]
        frames = self.trajectory(prim0, acts)
        fr_d = [[t.detach() for t in f] for f in frames]
        state = self.initial_state() if state is None else tuple(torch.as_tensor(t, dtype=DT) for t in state)
        states = [state]
        per_step = []
        total = 0.0
        T = len(acts)
        with torch.no_grad():
            for i in range(T):
                for s in range(i * S, (i + 1) * S):
                    state = sim.substep(state, fr_d[s], fr_d[s + 1])
                    states.append(state)
                if self.loss is not None and loss_every_step:
                    info = self.loss.value(state[0], fr_d[(i + 1) * S])
                    per_step.append(info)
                    total += info['loss']
        out = dict(loss=total, per_step=per_step, final_state=state, frames=fr_d)
        if keep_states:
            out['states'] = states
        if not with_grad:
            return out
        # ---------------- backward (Tape replay)
        n = self.n
        adj = (torch.zeros(n, 3, dtype=DT), torch.zeros(n, 3, dtype=DT), torch.zeros(n, 3, 3, dtype=DT),
               torch.zeros(n, 3, 3, dtype=DT))
        gfr = [[torch.zeros_like(t) for t in f] for f in fr_d]
        for i in reversed(range(T)):
            f_end = (i + 1) * S
            if self.loss is not None and loss_every_step:
                gx, gp = self.loss.vjp(states[f_end][0], fr_d[f_end])
                adj = (adj[0] + gx, adj[1], adj[2], adj[3])
                for k, g in enumerate(gp):
                    gfr[f_end][k] = gfr[f_end][k] + g
            for s in reversed(range(i * S, f_end)):
                adj, g0, g1 = sim.substep_vjp(states[s], fr_d[s], fr_d[s + 1], adj)
                for k in range(len(g0)):
                    gfr[s][k] = gfr[s][k] + g0[k]
                    gfr[s + 1][k] = gfr[s + 1][k] + g1[k]
        flat_out = [t for f in frames[1:] for t in f]
        flat_g = [t for f in gfr[1:] for t in f]
        keep = [(o, g) for o, g in zip(flat_out, flat_g) if o.requires_grad]
        if keep:
            (gact,) = torch.autograd.grad([o for o, _ in keep], [acts], grad_outputs=[g for _, g in keep],
                                          allow_unused=True)
            gact = torch.zeros_like(acts) if gact is None else gact
        else:
            gact = torch.zeros_like(acts)
        out['grad'] = gact.detach().numpy()
        out['adj0'] = adj
        out['prim_grads'] = gfr
        return out

    # ------------------------------------------------------------------ policy path (plb/engine/nn/mlp.py, solver_nn.py:33-43)
    def policy_obs_index(self, n_observed=200):
        step = self.n // n_observed
        return np.arange(self.n // step) * step

    @staticmethod
    def policy_unpack(params, dims):
        """flattened W0, b0, W1, b1, ... (+ optional velocity_weight) -> lists of tensors (mlp.py:167-183)"""
        params = torch.as_tensor(np.asarray(params, dtype=np.float64))
        Ws, bs, o = [], [], 0
        for i in range(len(dims) - 1):
            n = dims[i + 1] * dims[i]
            Ws.append(params[o:o + n].reshape(dims[i + 1], dims[i])); o += n
            bs.append(params[o:o + dims[i + 1]]); o += dims[i + 1]
        vw = float(params[o]) if len(params) - o == 1 else 1.0
        assert len(params) - o in (0, 1)
        return Ws, bs, vw

    def rollout_policy(self, params, horizon, hidden=(256, 256), activation='relu', n_observed=200, softness=666.0,
                       with_grad=True):
        """Episode driven by the reference's MLP policy: at env step t the observation is x, v * velocity_weight of every
        (N // n_observed)-th particle and position + rotation of every primitive at frame t*S (mlp.py:63-89); the action is
        the network output clamped to [-1, 1] (mlp.py:91-103).  Returns dict(loss, grad[len(params)] wrt W/b, actions).
        Backward = the tape replay: per env step (descending) loss adjoint, S substep adjoints, then the kinematics chain of
        that step (autograd through `fk`), the clamp, the network, and the observation adjoint into x, v of the observed
        particles and into the primitive pose of frame t*S."""
        sim = self.sim
        sim.set_softness(softness)
        S = sim.substeps
        idx = torch.as_tensor(self.policy_obs_index(n_observed))
        n_obs, P = len(idx), len(sim.prims)
        A = self.action_dims[-1]
        dims = (n_obs * 6 + 7 * P,) + tuple(hidden) + (A,)
        Ws, bs, vw = self.policy_unpack(params, dims)

        def net(obs, Wl, bl):
            h = obs
            for i in range(len(Wl)):
                h = Wl[i] @ h + bl[i]
                if i != len(Wl) - 1:
                    # Taichi routes max(z, 0) by the strict comparison 0 < z; torch.relu's subgradient at 0 is 0 as well
                    h = torch.relu(h) if activation == 'relu' else (torch.tanh(h) if activation == 'tanh' else h)
            # max(min(h, 1), -1) with strict routing == clamp (its gradient is 1 only strictly inside up to the measure-zero ends)
            return torch.where((h < 1.0) & (-1.0 < h), h, h.detach().clamp(-1.0, 1.0))

        def step_frames(pose0, a):
            """frames t*S+1 .. (t+1)*S as a function of the pose at t*S and the (already clamped) action"""
            out, cur = [], list(pose0)
            for _ in range(S):
                nxt = []
                for k, p in enumerate(sim.prims):
                    if p.action_dim > 0:
                        v, w, gv = p.velocities(a[self.action_dims[k]:self.action_dims[k + 1]])
                        v, w = v / S, w / S
                        gv = None if gv is None else gv / S
                    else:
                        v, w, gv = torch.zeros(3, dtype=DT), torch.zeros(3, dtype=DT), torch.zeros((), dtype=DT)
                    nxt.append(p.fk(cur[k], v, w, gv))
                out.append(nxt)
                cur = nxt
            return out

        def observation(state, pose):
            return torch.cat([torch.cat([state[0][idx], state[1][idx] * vw], dim=1).reshape(-1)] + [q[:7] for q in pose])

        # ---------------- forward
        state = self.initial_state()
        frames = [[q.detach() for q in self.initial_prims()]]
        states, actions, total = [state], [], 0.0
        with torch.no_grad():
            for t in range(horizon):
                a = net(observation(state, frames[t * S]), Ws, bs)
                actions.append(a)
                frames += [[q.detach() for q in f] for f in step_frames(frames[t * S], a)]
                for s_ in range(t * S, (t + 1) * S):
                    state = sim.substep(state, frames[s_], frames[s_ + 1])
                    states.append(state)
                if self.loss is not None:
                    total += self.loss.value(state[0], frames[(t + 1) * S])['loss']
        out = dict(loss=total, actions=torch.stack(actions).numpy(), final_state=state)
        if not with_grad:
            return out
        # ---------------- backward
        n = self.n
        adj = (torch.zeros(n, 3, dtype=DT), torch.zeros(n, 3, dtype=DT), torch.zeros(n, 3, 3, dtype=DT), torch.zeros(n, 3, 3, dtype=DT))
        gfr = [[torch.zeros_like(q) for q in f] for f in frames]
        gW = [torch.zeros_like(w) for w in Ws]
        gb = [torch.zeros_like(b) for b in bs]
        for t in reversed(range(horizon)):
            f_end = (t + 1) * S
            if self.loss is not None:
                gx, gp = self.loss.vjp(states[f_end][0], frames[f_end])
                adj = (adj[0] + gx, adj[1], adj[2], adj[3])
                for k, g in enumerate(gp):
                    gfr[f_end][k] = gfr[f_end][k] + g
            for s_ in reversed(range(t * S, f_end)):
                adj, g0, g1 = sim.substep_vjp(states[s_], frames[s_], frames[s_ + 1], adj)
                for k in range(P):
                    gfr[s_][k] = gfr[s_][k] + g0[k]
                    gfr[s_ + 1][k] = gfr[s_ + 1][k] + g1[k]
            # kinematics chain of env step t: pose(t*S), action -> frames t*S+1 .. f_end
            pose0 = [q.clone().requires_grad_(True) for q in frames[t * S]]
            x_obs = states[t * S][0][idx].clone().requires_grad_(True)
            v_obs = states[t * S][1][idx].clone().requires_grad_(True)
            Wl = [w.clone().requires_grad_(True) for w in Ws]
            bl = [b.clone().requires_grad_(True) for b in bs]
            obs = torch.cat([torch.cat([x_obs, v_obs * vw], dim=1).reshape(-1)] + [q[:7] for q in pose0])
            a = net(obs, Wl, bl)
            fr = step_frames(pose0, a)
            outs = [q for f in fr for q in f]
            gouts = [g for f in gfr[t * S + 1: f_end + 1] for g in f]
            keep = [(o, g) for o, g in zip(outs, gouts) if o.requires_grad]
            inputs = pose0 + [x_obs, v_obs] + Wl + bl
            grads = torch.autograd.grad([o for o, _ in keep], inputs, grad_outputs=[g for _, g in keep], allow_unused=True)
            grads = [torch.zeros_like(i) if g is None else g for g, i in zip(grads, inputs)]
            for k in range(P):
                gfr[t * S][k] = gfr[t * S][k] + grads[k]
            gx_obs, gv_obs = grads[P], grads[P + 1]
            ax, av = adj[0].clone(), adj[1].clone()
            ax[idx] += gx_obs
            av[idx] += gv_obs
            adj = (ax, av, adj[2], adj[3])
            for i in range(len(Ws)):
                gW[i] += grads[P + 2 + i]
                gb[i] += grads[P + 2 + len(Ws) + i]
        out['grad'] = np.concatenate([a_.numpy().reshape(-1) for i in range(len(Ws)) for a_ in (gW[i], gb[i])])
        return out

