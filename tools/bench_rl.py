#!/usr/bin/env python
"""Forward-only (RL, copy mode) throughput of the stock Move-v1 task -- BASELINE.json configs[0] (10k particles, 64^3) --
for K envs stepped together on one GPU (envs/vec_env.py).  Prints one JSON line per K: env steps/s and particle-substeps/s.
Usage: tools/bench_rl.py [--envs 1,4,16,64] [--steps 50] [--dtype float32]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="1,4,16,64")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--task", default="Move-v1")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as entry
    entry.build()
    from plasticinelab_b200.envs.vec_env import VecPlasticineEnv
    for K in [int(k) for k in args.envs.split(",")]:
        vec = VecPlasticineEnv(args.task, K, dtype=args.dtype)
        sim = vec.envs[0].taichi_env.simulator
        rng = np.random.RandomState(0)
        acts = rng.uniform(-1, 1, (args.steps, K, vec.action_space.shape[0]))
        vec.reset()
        for t in range(5):
            vec.step(acts[t])
        vec.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(args.steps):
            vec.step(acts[t])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"metric": "forward-only env steps/s (gym surface, loss + observation read back every step)", "task": args.task,
                          "n_envs": K, "n_particles": sim.n_particles, "substeps_per_env_step": sim.substeps, "dtype": args.dtype,
                          "env_steps_per_s": K * args.steps / dt, "particle_substeps_per_s": K * args.steps * sim.substeps * sim.n_particles / dt,
                          "ms_per_vector_step": 1e3 * dt / args.steps}), flush=True)
        vec.close()
        del vec


if __name__ == "__main__":
    main()
