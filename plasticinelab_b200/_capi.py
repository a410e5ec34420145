"""ctypes binding of the engine's C ABI (`include/plb_b200.h`).

The shared library `libplb_b200.so` is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no
CPU fallback: if the library is missing or no CUDA device is present, creating an engine raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PLB_F32, PLB_F64 = 0, 1
PRIM_TYPE_ID = dict(Sphere=0, Capsule=1, RollingPin=2, Chopsticks=3, Cylinder=4, Torus=5, Box=6)
MAX_PRIMITIVES = 8
MAX_ACTION_DIM = 7

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLB_LIB") or os.path.join(_HERE, "libplb_b200.so")


class PrimitiveDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int),
        ("params", C.c_double * 4),
        ("friction", C.c_double),
        ("init_state", C.c_double * 8),
        ("lower_bound", C.c_double * 3),
        ("upper_bound", C.c_double * 3),
        ("action_dim", C.c_int),
        ("action_scale", C.c_double * MAX_ACTION_DIM),
        ("minimal_gap", C.c_double),
    ]


class Config(C.Structure):
    _fields_ = [
        ("dtype", C.c_int),
        ("n_particles", C.c_int),
        ("n_grid", C.c_int),
        ("substeps", C.c_int),
        ("max_frames", C.c_int),
        ("max_prim_frames", C.c_int),
        ("dt", C.c_double), ("dx", C.c_double), ("p_vol", C.c_double), ("p_mass", C.c_double),
        ("E", C.c_double), ("nu", C.c_double),
        ("yield_stress", C.c_double),
        ("ground_friction", C.c_double),
        ("gravity", C.c_double * 3),
        ("n_primitives", C.c_int),
        ("device", C.c_int),
        ("kernel_variant", C.c_int),
    ]


def _tuple(v, n=None):
    if isinstance(v, str):
        v = eval(v)
    v = tuple(float(x) for x in v)
    assert n is None or len(v) == n, (v, n)
    return v


def primitive_desc(cfg: dict) -> PrimitiveDesc:
    """Per-primitive cfg (defaults as in primive_base.py:208-224 and primitives.py default_config)."""
    d = PrimitiveDesc()
    shape = cfg["shape"]
    d.type = PRIM_TYPE_ID[shape]
    p = [0.0] * 4
    if shape == "Sphere":
        p[0] = float(cfg.get("radius", 1.0))
    elif shape in ("Capsule", "RollingPin", "Chopsticks"):
        p[0], p[1] = float(cfg.get("h", 0.06)), float(cfg.get("r", 0.03))
    elif shape == "Cylinder":
        p[0], p[1] = float(cfg.get("h", 0.2)), float(cfg.get("r", 0.1))
    elif shape == "Torus":
        p[0], p[1] = float(cfg.get("tx", 0.2)), float(cfg.get("ty", 0.1))
    elif shape == "Box":
        p[0:3] = _tuple(cfg.get("size", (0.1, 0.1, 0.1)), 3)
    d.params = (C.c_double * 4)(*p)
    d.friction = float(cfg.get("friction", 0.9))
    st = _tuple(cfg.get("init_pos", (0.3, 0.3, 0.3)), 3) + _tuple(cfg.get("init_rot", (1.0, 0.0, 0.0, 0.0)), 4)
    st = st + (float(cfg.get("init_gap", 0.06)) if shape == "Chopsticks" else 0.0,)
    d.init_state = (C.c_double * 8)(*st)
    d.lower_bound = (C.c_double * 3)(*_tuple(cfg.get("lower_bound", (0.0, 0.0, 0.0)), 3))
    d.upper_bound = (C.c_double * 3)(*_tuple(cfg.get("upper_bound", (1.0, 1.0, 1.0)), 3))
    action = cfg.get("action") or {}
    d.action_dim = int(action.get("dim", 0) or 0)
    scale = list(_tuple(action.get("scale", ()))) if d.action_dim > 0 else []
    assert len(scale) >= d.action_dim and d.action_dim <= MAX_ACTION_DIM
    d.action_scale = (C.c_double * MAX_ACTION_DIM)(*(scale + [0.0] * (MAX_ACTION_DIM - len(scale))))
    d.minimal_gap = float(cfg.get("minimal_gap", 0.06))
    return d


def sim_constants(sim_cfg: dict) -> dict:
    """Derived constants of MPMSimulator.__init__ (plb/engine/mpm_simulator.py:14-34), 3-D."""
    assert int(sim_cfg.get("dim", 3)) == 3, "only the 3-D simulator is implemented"
    quality = float(sim_cfg.get("quality", 1)) * 0.5
    n_grid = int(128 * quality)
    dx = 1.0 / n_grid
    dt = 0.5e-4 / quality
    p_vol = (dx * 0.5) ** 2
    g = sim_cfg.get("gravity", (0, -1, 0))
    return dict(n_grid=n_grid, dx=dx, inv_dx=float(n_grid), dt=dt, p_vol=p_vol, p_rho=1, p_mass=p_vol * 1,
                substeps=int(2e-3 // dt), E=float(sim_cfg.get("E", 5e3)), nu=float(sim_cfg.get("nu", 0.2)),
                yield_stress=float(sim_cfg.get("yield_stress", 50.0)),
                ground_friction=float(sim_cfg.get("ground_friction", 1.5)), gravity=_tuple(g, 3))


def make_config(sim_cfg: dict, n_particles: int, n_primitives: int, dtype="float32", max_frames=None,
                max_prim_frames=None, device=0, kernel_variant=0) -> Config:
    k = sim_constants(sim_cfg)
    c = Config()
    c.dtype = PLB_F64 if str(dtype) in ("float64", "f64", "double") else PLB_F32
    c.n_particles = int(n_particles)
    c.n_grid = k["n_grid"]
    c.substeps = k["substeps"]
    c.max_frames = int(max_frames if max_frames is not None else sim_cfg.get("max_steps", 1024))
    c.max_prim_frames = int(max_prim_frames if max_prim_frames is not None else max(c.max_frames, 2))
    c.dt, c.dx, c.p_vol, c.p_mass = k["dt"], k["dx"], k["p_vol"], k["p_mass"]
    c.E, c.nu = k["E"], k["nu"]
    c.yield_stress = k["yield_stress"]
    c.ground_friction = k["ground_friction"]
    c.gravity = (C.c_double * 3)(*k["gravity"])
    c.n_primitives = int(n_primitives)
    c.device = int(device)
    c.kernel_variant = int(kernel_variant)
    return c


def dptr(a):
    """float64 C-contiguous numpy array (or None) -> double*"""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


_D = C.POINTER(C.c_double)
_SIGNATURES = {
    "plb_create": ([C.POINTER(Config), C.POINTER(PrimitiveDesc), C.POINTER(C.c_void_p)], C.c_int),
    "plb_destroy": ([C.c_void_p], C.c_int),
    "plb_last_error": ([C.c_void_p], C.c_char_p),
    "plb_set_stream": ([C.c_void_p, C.c_void_p], C.c_int),
    "plb_synchronize": ([C.c_void_p], C.c_int),
    "plb_abi_version": ([], C.c_int),
    "plb_set_materials": ([C.c_void_p, _D, _D, _D], C.c_int),
    "plb_set_frame": ([C.c_void_p, C.c_int, _D, _D, _D, _D], C.c_int),
    "plb_get_frame": ([C.c_void_p, C.c_int, _D, _D, _D, _D], C.c_int),
    "plb_copy_frame": ([C.c_void_p, C.c_int, C.c_int], C.c_int),
    "plb_sort_particles": ([C.c_void_p, C.c_int], C.c_int),
    "plb_frame_device_ptr": ([C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_int)], C.c_int),
    "plb_set_primitive_state": ([C.c_void_p, C.c_int, C.c_int, _D], C.c_int),
    "plb_get_primitive_state": ([C.c_void_p, C.c_int, C.c_int, _D], C.c_int),
    "plb_copy_primitive_frame": ([C.c_void_p, C.c_int, C.c_int], C.c_int),
    "plb_set_softness": ([C.c_void_p, C.c_double], C.c_int),
    "plb_set_action": ([C.c_void_p, C.c_int, C.c_int, _D, C.c_int], C.c_int),
    "plb_kinematics": ([C.c_void_p, C.c_int, C.c_int], C.c_int),
    "plb_substep_fwd": ([C.c_void_p, C.c_int, C.c_int, C.c_int], C.c_int),
    "plb_step_fwd": ([C.c_void_p, C.c_int, C.c_int, C.c_int], C.c_int),
    "plb_substep_bwd": ([C.c_void_p, C.c_int, C.c_int], C.c_int),
    "plb_step_bwd": ([C.c_void_p, C.c_int, C.c_int, C.c_int], C.c_int),
    "plb_zero_grads": ([C.c_void_p], C.c_int),
    "plb_set_adjoint": ([C.c_void_p, _D, _D, _D, _D], C.c_int),
    "plb_get_adjoint": ([C.c_void_p, _D, _D, _D, _D], C.c_int),
    "plb_get_primitive_grads": ([C.c_void_p, C.c_int, C.c_int, _D], C.c_int),
    "plb_get_action_grad": ([C.c_void_p, C.c_int, C.c_int, _D], C.c_int),
    "plb_action_grad_step": ([C.c_void_p, C.c_int, C.c_int, _D], C.c_int),
    "plb_add_pose_adjoint": ([C.c_void_p, C.c_int, _D], C.c_int),
    "plb_gather_particles": ([C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, _D, _D], C.c_int),
    "plb_scatter_adjoint": ([C.c_void_p, C.POINTER(C.c_int), C.c_int, _D, _D], C.c_int),
    "plb_set_target": ([C.c_void_p, _D, _D], C.c_int),
    "plb_get_target_sdf": ([C.c_void_p, _D], C.c_int),
    "plb_set_loss_weights": ([C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int], C.c_int),
    "plb_loss_fwd": ([C.c_void_p, C.c_int, C.c_int, _D], C.c_int),
    "plb_loss_bwd": ([C.c_void_p, C.c_int, C.c_int], C.c_int),
    "plb_get_loss": ([C.c_void_p, _D], C.c_int),
    "plb_clear_loss": ([C.c_void_p], C.c_int),
    "plb_debug_get_grid": ([C.c_void_p, _D, _D], C.c_int),
    "plb_launch_count": ([C.c_void_p], C.c_longlong),
    "plb_slab_configure": ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
    "plb_slab_buffer": ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)], C.c_int),
    "plb_slab_fwd_p2g": ([C.c_void_p, C.c_int, C.c_int], C.c_int),
    "plb_slab_fwd_finish": ([C.c_void_p, C.c_int, C.c_int, C.c_int], C.c_int),
    "plb_slab_bwd_begin": ([C.c_void_p, C.c_int, C.c_int], C.c_int),
    "plb_slab_bwd_finish": ([C.c_void_p, C.c_int, C.c_int], C.c_int),
    "plb_slab_loss_begin": ([C.c_void_p, C.c_int], C.c_int),
    "plb_slab_loss_reduce": ([C.c_void_p, C.c_int, C.c_int], C.c_int),
    "plb_slab_loss_finish": ([C.c_void_p, C.c_int, C.c_int, C.c_int, _D], C.c_int),
    "plb_device_buffer": ([C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)], C.c_int),
    "plb_slab_ipc_export": ([C.c_void_p, C.c_int, C.c_void_p], C.c_int),
    "plb_slab_ipc_import": ([C.c_void_p, C.c_int, C.c_void_p], C.c_int),
    "plb_slab_ipc_close": ([C.c_void_p], C.c_int),
    "plb_slab_ipc_export_grid": ([C.c_void_p, C.c_int, C.c_void_p], C.c_int),
    "plb_slab_ipc_import_grid": ([C.c_void_p, C.c_int, C.c_int, C.c_void_p], C.c_int),
    "plb_profile_enable": ([C.c_void_p, C.c_int], C.c_int),
    "plb_profile_read": ([C.c_void_p, C.c_int, _D, C.POINTER(C.c_longlong)], C.c_int),
    "plb_kernel_name": ([C.c_int], C.c_char_p),
    "plb_count_active": ([C.c_void_p, C.c_int, C.POINTER(C.c_longlong)], C.c_int),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load_library(path: str | None = None):
    """dlopen the engine library and type its entry points.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise RuntimeError(f"{p} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()'). "
                           "There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    if path is None:
        _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


class Engine:
    """Thin RAII wrapper: one engine handle, errors turned into exceptions."""

    def __init__(self, config: Config, prim_descs):
        self.lib = load_library()
        self.config = config
        n = len(prim_descs)
        arr = (PrimitiveDesc * max(n, 1))(*prim_descs)
        self._prims = arr
        h = C.c_void_p()
        rc = self.lib.plb_create(C.byref(config), arr, C.byref(h))
        if rc != 0:
            raise EngineError(f"plb_create failed ({rc}): {self.lib.plb_last_error(None).decode()}")
        self.h = h

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != 0:
            raise EngineError(f"{name} failed ({rc}): {self.lib.plb_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.plb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
