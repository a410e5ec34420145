"""State-feedback MLP policy differentiated through the simulator: the mirror of `plb/engine/nn/mlp.py:12-183`.

The reference writes the network as Taichi kernels so that `ti.Tape` carries gradients from the loss through the simulator
into the weights.  Here the network itself is a few dense layers in float64 numpy on the host (a 1207 x 256 x 256 x A
network is ~0.4 MFLOP per env step, nothing next to 39 substeps of 10^5 particles); what needs the engine is its coupling to the
simulator, which goes through four C-ABI calls:

  observation   plb_gather_particles(frame t*S, every (N // 200)-th particle)  +  primitive poses of that frame
                (`input_particles` / `input_primitives`, mlp.py:63-89)
  action        clamp to [-1, 1] -> plb_set_action(t, S)  (`set_action` kernel + `set_velocity`, mlp.py:91-103,145-155)
  backward      when the tape's reverse sweep stands at frame t*S:  plb_action_grad_step(t) gives d loss / d action_t;
                the clamp routes it like Taichi's max/min (strict comparisons); the dense layers are differentiated by hand;
                the observation adjoint goes back with plb_scatter_adjoint (x, v of the observed particles) and
                plb_add_pose_adjoint (position, rotation of every primitive).

Same public surface: `MLP(simulator, primitives, hidden_dims, activation='relu', n_observed_particles=200)`, `set_action(s,
n_substeps)`, `get_params()`, `set_params(p)` (an optional trailing scalar is `velocity_weight`), `get_grad()` -- parameters
and gradients flattened as W0, b0, W1, b1, ... with W[i] of shape (dims[i+1], dims[i]).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _capi
from ..tape import active_tape


class MLP:
    def __init__(self, simulator, primitives, hidden_dims, activation='relu', n_observed_particles=200):
        for p in primitives:
            if p.shape == 'Chopsticks':
                raise AssertionError("Chopstick is not supported now..")          # mlp.py:28-29
        if activation not in ('relu', 'tanh', None):
            raise ValueError(activation)
        self.simulator, self.primitives, self.activation = simulator, primitives, activation
        self.engine = simulator.engine
        n = simulator.n_particles
        self.n_observed_particles = n_observed_particles
        self.obs_step = n // n_observed_particles
        self.obs_num = n // self.obs_step
        self.obs_index = np.ascontiguousarray(np.arange(self.obs_num, dtype=np.int32) * self.obs_step)
        self.substeps = simulator.substeps
        self.dims = (self.obs_num * 6 + primitives.state_dim,) + tuple(hidden_dims) + (primitives.action_dim,)
        self.n_layer = len(self.dims) - 1
        self.W = [np.zeros((self.dims[i + 1], self.dims[i])) for i in range(self.n_layer)]
        self.b = [np.zeros(self.dims[i + 1]) for i in range(self.n_layer)]
        self.W_grad = [np.zeros_like(w) for w in self.W]
        self.b_grad = [np.zeros_like(b) for b in self.b]
        self.velocity_weight = 1.0
        self._acts = {}                  # env step -> (inputs of every layer, pre-activations, raw output)

    # ------------------------------------------------------------------ parameters
    def get_params(self):
        return np.concatenate([a.reshape(-1) for i in range(self.n_layer) for a in (self.W[i], self.b[i])])

    def set_params(self, param):
        param = np.asarray(param, dtype=np.float64).reshape(-1)
        for i in range(self.n_layer):
            n = self.W[i].size
            self.W[i] = param[:n].reshape(self.W[i].shape).copy()
            param = param[n:]
            n = self.b[i].size
            self.b[i] = param[:n].copy()
            param = param[n:]
        if len(param) == 1:
            self.velocity_weight = float(param[-1])
        else:
            self.velocity_weight = 1.0
            assert len(param) == 0

    def get_grad(self):
        return np.concatenate([a.reshape(-1) for i in range(self.n_layer) for a in (self.W_grad[i], self.b_grad[i])])

    def zero_grad(self):
        for a in self.W_grad + self.b_grad:
            a[...] = 0.0
        self._acts.clear()

    def zero_grad_for_tape(self):
        """First use under a new tape: ti.Tape zeroes every .grad on entry (solver_nn.py:35)."""
        for a in self.W_grad + self.b_grad:
            a[...] = 0.0

    # ------------------------------------------------------------------ forward
    def observe(self, t):
        """hidden[0][t] of the reference: x, v * velocity_weight of the observed particles, then 7 numbers per primitive."""
        f = int(t) * self.substeps
        x, v = np.zeros((self.obs_num, 3)), np.zeros((self.obs_num, 3))
        self.engine.call("plb_gather_particles", f, self.obs_index.ctypes.data_as(C.POINTER(C.c_int)), self.obs_num,
                         _capi.dptr(x), _capi.dptr(v))
        prim = [p.get_state(f)[:7] for p in self.primitives]
        return np.concatenate([np.concatenate([x, v * self.velocity_weight], axis=1).reshape(-1)] + prim)

    def _act(self, z):
        if self.activation == 'relu':
            return np.maximum(z, 0.0)
        if self.activation == 'tanh':
            return np.tanh(z)
        return z

    def forward(self, obs):
        hs, zs = [np.asarray(obs, dtype=np.float64)], []
        for i in range(self.n_layer):
            z = self.W[i] @ hs[-1] + self.b[i]
            zs.append(z)
            hs.append(self._act(z) if i != self.n_layer - 1 else z)
        return hs, zs

    def set_action(self, s, n_substeps):
        """Observe frame s*S, run the network, write the clamped output as the action of env step s (mlp.py:145-155)."""
        assert n_substeps == self.substeps
        hs, zs = self.forward(self.observe(s))
        self._acts[int(s)] = (hs, zs)
        self.primitives.set_action(int(s), int(n_substeps), np.clip(hs[-1], -1.0, 1.0))
        tape = active_tape()
        if tape is not None:
            tape.record_policy(self, int(s))

    # ------------------------------------------------------------------ backward (called by the tape, steps descending)
    def backward(self, s):
        hs, zs = self._acts.pop(int(s))
        A = self.dims[-1]
        ga = np.zeros(max(A, 1))
        self.engine.call("plb_action_grad_step", int(s), self.substeps, _capi.dptr(ga))
        out = hs[-1]
        g = ga[:A] * ((out < 1.0) & (-1.0 < out))          # max(min(h, 1), -1): to h iff h < 1 and -1 < h
        for i in reversed(range(self.n_layer)):
            if i != self.n_layer - 1:
                if self.activation == 'relu':
                    g = g * (0.0 < zs[i])                   # max(z, 0): to z iff 0 < z
                elif self.activation == 'tanh':
                    g = g * (1.0 - np.tanh(zs[i]) ** 2)
            self.W_grad[i] += np.outer(g, hs[i])
            self.b_grad[i] += g
            g = self.W[i].T @ g
        gp = g[: self.obs_num * 6].reshape(self.obs_num, 6)
        gx = np.ascontiguousarray(gp[:, :3])
        gv = np.ascontiguousarray(gp[:, 3:] * self.velocity_weight)
        self.engine.call("plb_scatter_adjoint", self.obs_index.ctypes.data_as(C.POINTER(C.c_int)), self.obs_num,
                         _capi.dptr(gx), _capi.dptr(gv))
        base = self.obs_num * 6
        for k in range(len(self.primitives)):
            g8 = np.zeros(8)
            g8[:7] = g[base + k * 7: base + k * 7 + 7]
            self.engine.call("plb_add_pose_adjoint", k, _capi.dptr(g8))
