"""The two `taichi` names PlasticineLab's solver touches, mapped onto this engine.

The reference's `plb/optimizer/solver.py:1,36` does `import taichi as ti` and `with ti.Tape(loss=env.loss.loss):`.
`install()` registers this module as `taichi` (only if the real one is absent), so code written against the reference
that just needs `ti.Tape` / `ti.init` keeps working with a `plasticinelab_b200` env:

    import plasticinelab_b200.ti_shim as shim; shim.install()
    import taichi as ti
    with ti.Tape(loss=taichi_env.loss.loss): ...

Nothing else of Taichi is emulated (no fields, no kernels): the engine is hand-written CUDA behind the C ABI.
"""
from __future__ import annotations

import sys

from .engine.tape import Tape  # noqa: F401  (ti.Tape(loss=env.loss.loss))

gpu = "cuda"
cpu = "cpu"
f32, f64 = "float32", "float64"


def init(*args, **kwargs):
    """`ti.init(arch=ti.gpu, ...)` (plb/engine/taichi_env.py:6): nothing to do, the CUDA context is created per engine."""
    return None


def install():
    if "taichi" not in sys.modules:
        sys.modules["taichi"] = sys.modules[__name__]
    return sys.modules["taichi"]
