"""B200-native differentiable MPM engine behind PlasticineLab's `TaichiEnv` / `plb.envs` / `plb.optimizer` surface.

Host orchestration in Python; all compute in hand-written sm_100a CUDA kernels reached through the C ABI declared in
`include/plb_b200.h` (built in-tree as `plasticinelab_b200/libplb_b200.so`).  No CPU fallback.
"""
__version__ = "0.1.0"
