#!/usr/bin/env python
"""Benchmark of the hot path: differentiable MPM substeps, forward + backward, particle-substeps / s.

  python bench.py --gpus N --steps K --warmup W [--workload slab1m] [--impl reference]

One "step" = one fwd+bwd episode of the workload (H env steps x S substeps forward with the loss after every env
step, then the full adjoint back to the action gradient) = N_particles * H * S particle-substeps.

ONE workload for every N (weak scaling): `slab1m`, 1M particles and 28 planes of a 256^3 grid per GPU.  At N = 1 it runs on
the regular single-GPU path (it is a 1M-particle single-GPU configuration), at N > 1 the same bar, N times as long, is
slab-decomposed.  At N = 1 the line additionally carries `also`: the north-star roofline size (`move1m`, 1M / 128^3),
BASELINE configs[1] (`move100k`) and configs[2] (`rope1m`) measured in the same process with the same rules.

  value     whole-job particle-substeps/s with the particle state already resident in HBM (device-timed, CUDA events);
  e2e       the same episode through the public API (`Solver.forward`): host float64 state -> device, per-env-step loss
            read-back, action gradient read-back;
  roofline  the dominant kernel OF THE TIMED PATH (the fused particle kernels the env-step graphs replay): algorithmic
            bytes per launch / mean CUDA-event duration of that kernel, against MEASURED_PEAKS.json.  The per-kernel times
            come from one more episode of the same workload in which the engine launches the graph's kernel sequence
            one by one with an event pair around each kernel (plb_profile_enable);
  parity    float32 (benchmarked) against the float64 engine on the same episode: action-gradient relative error, loss
            relative error, final positions (the float64 engine agrees with the float64 oracle to 1e-9, tests/);
  cpu_baseline  the float64 C/OpenMP restatement of the reference's Taichi kernels (Taichi itself is not installable here
            nor on the GPU box, gpurun_out/probe/taichi_probe.txt) timed on this box's host cores on a bounded sample.

--impl reference times that restatement alone (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-substeps/sec fwd+bwd"
UNIT = "particle-substeps/s"
DEFAULT_WORKLOAD = "slab1m"

WORKLOADS = {
    # BASELINE.json configs[1]
    "move100k": dict(scene="move.yml", n=100_000, quality=2, horizon=50,
                     desc="Move-v1 geometry, 100k particles, 128^3 grid, 50 env steps x 39 substeps, fwd+bwd action gradient"),
    # north_star roofline target size
    "move1m": dict(scene="move.yml", n=1_000_000, quality=2, horizon=10,
                   desc="Move-v1 geometry, 1M particles, 128^3 grid, 10 env steps x 39 substeps, fwd+bwd"),
    # the full 50-step episode at 1M particles: 1951 frames would need 187 GB, so env-step checkpointing (91 frames, 8.7 GB)
    "move1m_ckpt": dict(scene="move.yml", n=1_000_000, quality=2, horizon=50, checkpoint=True,
                        desc="Move-v1 geometry, 1M particles, 128^3 grid, 50 env steps x 39 substeps, fwd+bwd with env-step "
                             "checkpointing (one extra forward pass; a fwd+bwd substep still counts once)"),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case), here as fwd+bwd
    "move10k": dict(scene="move.yml", n=10_000, quality=1, horizon=50,
                    desc="Move-v1 stock, 10k particles, 64^3 grid, 50 env steps x 19 substeps, fwd+bwd"),
    # a translating body (every particle starts with velocity (2, 0, 2): ~0.5 cells per env step, 5 cells over the episode):
    # exercises the TMA windows that follow the material (k_chunk_origins; PLB_WINDOW_FOLLOW=0 pins them to the sort-time blocks)
    "fly1m": dict(scene="move.yml", n=1_000_000, quality=2, horizon=10, v0=(2.0, 0.0, 2.0),
                  desc="Move-v1 geometry with initial velocity (2, 0, 2), 1M particles, 128^3 grid, 10 env steps x 39 substeps, fwd+bwd"),
    # weak scaling (north_star / BASELINE configs[4]-style): an elastic-plastic bar along the slab axis, 1M particles and
    # 0.109 of the domain (28 planes of 256) per GPU; --gpus N decomposes the N-times-longer bar into N slabs
    # (slab runs: 4-plane halo zones -- the material moves less than a plane per env step here; every env step checks that all
    #  stencils stay inside the slab +- halo, k_check_margin)
    "slab1m": dict(scene="slab", n=1_000_000, quality=4, horizon=2, halo_w=4,
                   desc="bar (0.109*N) x 0.1 x 0.1 on the ground pressed by two spheres, 1M particles per GPU, 256^3 grid, "
                        "2 env steps x 79 substeps, fwd+bwd"),
    # BASELINE.json configs[2]
    "rope1m": dict(scene="rope.yml", n=1_000_000, quality=4, horizon=2,
                   desc="Rope-v1 geometry, 1M particles, 256^3 grid, 2 env steps x 79 substeps, fwd+bwd"),
    # BASELINE.json configs[3]: Torus-v1 geometry (sticky ground), two materials split at x = 0.5, 1M particles per GPU
    # (4M on 4 GPUs), 256^3.  The box is 0.3 wide along the slab axis: 4 slabs of ~19 planes would be thinner than two
    # 8-plane halos, so the slab run uses 4-plane halos (the material moves < 1 plane per env step here).
    "torus4m": dict(scene="torus.yml", n=1_000_000, quality=4, horizon=1, materials="split", scale_n=True, halo_w=4,
                    desc="Torus-v1 geometry (box 0.3 x 0.1 x 0.3, Torus primitive, sticky ground), per-particle mu/lam/yield "
                         "(E 3e3 / yield 50 for x < 0.5, E 6.5e3 / yield 200 otherwise), 1M particles per GPU, 256^3 grid, "
                         "1 env step x 79 substeps, fwd+bwd"),
    # BASELINE.json configs[4]: synthetic elastic block, 2M particles and 64 planes of 512^3 per GPU, no primitives
    "block16m": dict(scene="block", n=2_000_000, quality=8, horizon=1,
                     desc="elastic block (yield 1e9, no primitives) (0.1125*N) x 0.2 x 0.2, 2M particles per GPU, 512^3 grid, "
                          "1 env step x 159 substeps, fwd+bwd"),
}

# algorithmic bytes per launch of each kernel: (bytes per particle, bytes per active node), float32 scalars.
# g2p_p2g / p2g_bwd_g2p_bwd are the fused particle kernels the env-step graphs replay.
KERNEL_BYTES = {
    "p2g": (132, 16), "grid_fwd": (0, 28), "g2p": (72, 12), "p2g_recompute": (0, 16), "grid_fwd_recompute": (0, 28),
    "g2p_bwd": (84, 24), "grid_bwd": (0, 44), "p2g_bwd": (240, 16), "g2p_p2g": (144, 28), "p2g_bwd_g2p_bwd": (204, 40),
}

# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the fused kernels from `ncu --set full` captures of
# this bench at the named workload (profiles/r2_*_ncu_full_summary.md); `roofline.traffic` is null for anything else.
NCU_TRAFFIC = {("slab1m", "p2g_bwd_g2p_bwd"): 370.6e6, ("slab1m", "g2p_p2g"): 212.1e6}
NCU_TRAFFIC_SOURCE = {"slab1m": "profiles/r2j_slab1m_{bwd_warp,fwd_chunk}_ncu_full_raw.csv.gz (dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                                "above the algorithmic figure because the SVD store adds 84 B written / 84 B read per particle and substep)"}


def build_cfg(w, world=1):
    from plasticinelab_b200.envs.scene import load_variants
    from plasticinelab_b200 import _capi
    from plasticinelab_b200.config import load_dict
    if w["scene"] == "slab":
        L = 0.109375 * world          # 28 planes of 256 per GPU: slab boundaries fall on 4-plane block boundaries
        tree = dict(SIMULATOR=dict(quality=w["quality"], yield_stress=50.0, ground_friction=0.3),
                    SHAPES=[dict(shape="box", width=(L, 0.1, 0.1), init_pos=(0.5, 0.06, 0.5), n_particles=w["n"] * world)],
                    PRIMITIVES=[dict(shape="Sphere", radius=0.03, init_pos=(0.5 - 0.3 * L, 0.145, 0.5), friction=0.9,
                                     action=dict(dim=3, scale=(0.01, 0.01, 0.01))),
                                dict(shape="Sphere", radius=0.03, init_pos=(0.5 + 0.3 * L, 0.145, 0.5), friction=0.9,
                                     action=dict(dim=3, scale=(0.01, 0.01, 0.01)))])
        cfg = load_dict(tree)
        cfg.ENV.loss.target_path = "envs/assets/Rope3D-v1.npy"
    elif w["scene"] == "block":
        L = 0.1125 * world            # 57.6 planes of 512 per GPU (BASELINE configs[4]: 0.9 of the domain on 8 GPUs), along the slab axis
        tree = dict(SIMULATOR=dict(quality=w["quality"], yield_stress=1e9),
                    SHAPES=[dict(shape="box", width=(L, 0.2, 0.2), init_pos=(0.5, 0.2, 0.5), n_particles=w["n"] * world)],
                    PRIMITIVES=[])
        cfg = load_dict(tree)
        cfg.ENV.loss.target_path = "envs/assets/Rope3D-v1.npy"
    else:
        cfg = load_variants(w["scene"], 1)
        cfg.SIMULATOR.quality = w["quality"]
        cfg.SHAPES[0]["n_particles"] = w["n"] * (world if w.get("scale_n") else 1)
    S = _capi.sim_constants(dict(cfg.SIMULATOR))["substeps"]
    cfg.SIMULATOR.max_steps = (S + 1 + w["horizon"] + 2) if w.get("checkpoint") else (w["horizon"] * S + 2)
    return cfg, S


def split_materials(x0):
    """BASELINE configs[3] 'multi-material': E 3e3 / yield 50 for x < 0.5, E 6.5e3 / yield 200 otherwise (nu 0.2).
    (SURVEY.md 8d suggests E 2e4 for the stiff half "e.g."; at 256^3 the reference's fixed dt = 2.5e-5 puts its P-wave at
    0.95 cells per substep, beyond the usual stability limit of an explicit MPM step (the 200k-particle instance of this scene
    left its env-step block list within one env step, gpurun_out/multi4), so the stiff half stays at 0.54 cells per substep.)"""
    stiff = x0[:, 0] >= 0.5
    E = np.where(stiff, 6.5e3, 3e3)
    return E / 2.4, E * 0.2 / (1.2 * 0.6), np.where(stiff, 200.0, 50.0)


def actions_for(w, A, horizon=None):
    a = np.random.RandomState(0).uniform(-0.01, 0.01, (horizon or w["horizon"], max(A, 1)))[:, :A]
    if w["scene"] == "slab":
        a[:, 1::3] = -0.5          # press the spheres into the bar
    if w["scene"] == "torus.yml":
        a[:, 1] = -0.5             # lower the torus onto the box
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                p = [t.strip() for t in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def oracle_sample(cfg, w, S, n_sub, threads):
    """Bounded CPU sample: n_sub fwd+bwd substeps of the same scene on the host cores.
    Sphere-only scenes use the plain-C/OpenMP port of the reference kernels (oracle/mpm_oracle.c, dense grid sweeps and
    atomics like Taichi's x64 backend); other scenes fall back to the torch-CPU oracle.  Returns (value, seconds, kind)."""
    import torch
    import __graft_entry__ as entry
    from oracle.plb_oracle import OracleEnv
    from plasticinelab_b200.engine.shapes import Shapes
    entry.build_oracle()
    torch.set_num_threads(threads)
    x0, _ = Shapes(cfg.SHAPES).get()
    oenv = OracleEnv(cfg, x0, None)
    sim = oenv.sim
    sim.set_softness(666.0)
    A = oenv.action_dims[-1]
    acts = torch.as_tensor(actions_for(w, A)[:1])
    frames = oenv.trajectory(oenv.initial_prims(), acts)
    fr = [[t.detach() for t in f] for f in frames]
    state = oenv.initial_state()
    if all(p.shape == "Sphere" for p in sim.prims):
        from oracle.c_port import CPort
        port = CPort(sim, len(x0), 666.0)
        poses = [port.poses([t.numpy() for t in f]) for f in fr[:n_sub + 1]]
        st = tuple(t.numpy() for t in state)
        port.substep_fwd(st, poses[0], poses[1])          # warm the threads / page in the grid
        # "all the host threads it can use": more OpenMP threads are not always faster on a dual-socket host (atomics on
        # a shared dense grid), so time one fwd+bwd substep at a few thread counts and keep the fastest
        best = (None, 1e30)
        ones = tuple(np.ones_like(a) for a in st)
        for nt in sorted({threads, max(threads // 2, 1), max(threads // 4, 1), max(threads // 8, 1), min(threads, 16), min(threads, 8)}, reverse=True):
            port.lib.oc_set_threads(int(nt))
            t1 = time.perf_counter()
            port.substep_fwd(st, poses[0], poses[1])
            port.substep_bwd(st, poses[0], poses[1], ones)
            el = time.perf_counter() - t1
            if el < best[1]:
                best = (nt, el)
        port.lib.oc_set_threads(int(best[0]))
        port.threads = int(best[0])
        t0 = time.perf_counter()
        states = [st]
        for s in range(n_sub):
            st = port.substep_fwd(st, poses[s], poses[s + 1])
            states.append(st)
        adj = tuple(np.ones_like(a) for a in st)
        for s in reversed(range(n_sub)):
            adj, _, _ = port.substep_bwd(states[s], poses[s], poses[s + 1], adj)
        dt = time.perf_counter() - t0
        return (len(x0) * n_sub / dt, dt,
                f"C/OpenMP float64 port of the reference kernels, {port.threads} of {threads} threads (fastest of a thread-count sweep "
                f"on one fwd+bwd substep); substeps only: no per-env-step loss evaluation, adjoint seeded with ones", port.threads)
    t0 = time.perf_counter()
    states = [state]
    with torch.no_grad():
        for s in range(n_sub):
            state = sim.substep(state, fr[s], fr[s + 1])
            states.append(state)
    adj = tuple(torch.ones_like(t) for t in state)
    for s in reversed(range(n_sub)):
        adj, _, _ = sim.substep_vjp(states[s], fr[s], fr[s + 1], adj)
    dt = time.perf_counter() - t0
    return len(x0) * n_sub / dt, dt, f"float64 torch-CPU oracle, {threads} threads; substeps only, no loss evaluation", threads


def run_reference(args, w, rank):
    if rank != 0:
        return
    cfg, S = build_cfg(w)
    threads = os.cpu_count() or 1
    n_sub = max(1, int(os.environ.get("PLB_BENCH_CPU_SUBSTEPS", "4" if w["n"] >= 1_000_000 else "20")))
    vals, times = [], []
    kind, used = "", threads
    for i in range(args.warmup + args.steps):
        v, dt, kind, used = oracle_sample(cfg, w, S, n_sub, threads)
        if i >= args.warmup:
            vals.append(v); times.append(dt)
    value = float(np.mean(vals))
    sample = f"{n_sub} fwd+bwd substeps of the workload scene (one GPU's share) per step ({kind})"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": w["desc"], "n_particles": w["n"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class Job:
    """One workload on this rank: env, episodes (device-resident / end-to-end), timing, roofline, parity."""

    def __init__(self, args, wname, rank, world, local_rank, dist):
        import torch
        from plasticinelab_b200 import _capi
        from plasticinelab_b200.engine.taichi_env import TaichiEnv
        from plasticinelab_b200.optimizer.solver import Solver
        self.torch, self.capi, self.dist = torch, _capi, dist
        self.args, self.wname, self.w = args, wname, WORKLOADS[wname]
        self.rank, self.world, self.local_rank = rank, world, local_rank
        w = self.w
        self.slab = world > 1
        self.cfg, self.S = build_cfg(w, world)
        self.senv = None
        if self.slab:
            from plasticinelab_b200.engine.sharded import ShardedEnv
            self.senv = ShardedEnv(self.cfg, dtype=args.dtype, device=local_rank, halo_w=int(os.environ.get("PLB_BENCH_HALO_W", w.get("halo_w", 8))),
                                   materials=split_materials if w.get("materials") else None)
            self.env = self.senv.env
        else:
            self.env = TaichiEnv(self.cfg, dtype=args.dtype, device=local_rank, max_prim_frames=w["horizon"] * self.S + 2)
            self.env.initialize()
            if w.get("materials"):
                self.env.simulator.set_materials(*split_materials(self.env.init_particles))
        env = self.env
        env.loss.set_weights(10, 10, 1, False)
        if w.get("v0") and not self.slab:
            st = [np.array(a) for a in env.get_state()["state"]]
            st[1][:] = np.asarray(w["v0"])
            env.set_state(st, 666.0, False)
        self.eng = env.engine
        self.ckpt = None
        if w.get("checkpoint"):
            from plasticinelab_b200.engine.checkpoint import CheckpointedEpisode
            self.ckpt = CheckpointedEpisode(env, w["horizon"])
        self.eng.call("plb_set_stream", C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.N, self.H = env.n_particles, w["horizon"]
        self.N_global = self.senv.n_global if self.slab else self.N
        self.A = env.primitives.action_dim
        self.actions = actions_for(w, self.A)
        self.pinned_actions = torch.from_numpy(np.ascontiguousarray(self.actions)).pin_memory()
        # e2e inputs live in pinned host memory (particle state as float64 like the reference's get_state(), actions)
        self.host_state = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in env.get_state()["state"]]
        self.solver = Solver(env, None, None, n_iters=1, softness=666.0, horizon=self.H)
        self.solver.total_steps = 0
        self.grad_out = np.zeros((self.H, max(self.A, 1)))
        self.last_loss = None

    def close(self):
        if self.senv is not None:
            self.senv.close()          # (collective)
        else:
            self.eng.close()

    # ---- episodes
    def episode_slab(self, load_state=False):
        senv, env = self.senv, self.env
        if load_state:
            env.simulator.set_state(0, self.host_state)
        senv.begin_episode(666.0)
        a = self.pinned_actions.numpy()
        for i in range(self.H):
            senv.step(a[i])
            senv.compute_loss(sync=load_state)
        self.grad_out[:, :self.A] = senv.backward()

    def episode_device(self):
        """state resident in HBM (frame 0), no host read-back except the final action gradient"""
        env, eng, S, H, A, D = self.env, self.eng, self.S, self.H, self.A, self.capi.dptr
        if self.slab:
            return self.episode_slab(False)
        if self.ckpt is not None:
            self.grad_out[:, :A] = self.ckpt.forward_backward(self.pinned_actions.numpy())[1]
            return
        env.simulator.cur = 0
        env._is_copy = False
        for p in env.primitives:
            p.set_state(0, p.init_state)
        eng.call("plb_zero_grads")
        a = self.pinned_actions.numpy()
        for i in range(H):
            if A:
                eng.call("plb_set_action", i, S, D(np.ascontiguousarray(a[i])), A)
            eng.call("plb_kinematics", i * S, S)
            eng.call("plb_step_fwd", i * S, i * S, S)
            eng.call("plb_loss_fwd", (i + 1) * S, (i + 1) * S, None)
        for i in reversed(range(H)):
            eng.call("plb_loss_bwd", (i + 1) * S, (i + 1) * S)
            eng.call("plb_step_bwd", i * S, i * S, S)
        eng.call("plb_get_action_grad", H, S, D(self.grad_out))

    def episode_e2e(self):
        if self.slab:
            return self.episode_slab(True)
        if self.ckpt is not None:
            self.env.set_state(self.host_state, 666.0, False)
            self.last_loss, g = self.ckpt.forward_backward(self.pinned_actions.numpy(), sync_losses=True)
            self.grad_out[:, :self.A] = g
            return
        self.last_loss, g = self.solver.forward(self.host_state, self.pinned_actions.numpy())
        self.grad_out[:, :self.A] = g

    def timed(self, fn, k):
        torch, dist = self.torch, self.dist
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms

    # ---- float32 (benchmarked) vs float64 engine on the same episode
    def parity(self):
        """Runs on one GPU.  The float64 trajectory must fit beside nothing else: the float32 engine is closed first."""
        from plasticinelab_b200.engine.taichi_env import TaichiEnv
        from plasticinelab_b200.optimizer.solver import Solver
        w, S, N = self.w, self.S, self.N
        # frames of the float64 run that fit in ~60 GB (state 192 B + SVD store 168 B per particle and frame)
        Hp = max(1, min(self.H, int(60e9 / (N * 360.0) / S), 5 if self.ckpt is not None else self.H))
        acts = self.actions[:Hp]
        out = {}
        for dtype in (self.args.dtype, "float64"):
            cfg, _ = build_cfg(dict(w, horizon=Hp, checkpoint=False))
            env = TaichiEnv(cfg, dtype=dtype, device=self.local_rank, max_prim_frames=Hp * S + 2)
            env.initialize()
            if w.get("materials"):
                env.simulator.set_materials(*split_materials(env.init_particles))
            env.loss.set_weights(10, 10, 1, False)
            solver = Solver(env, None, None, n_iters=1, softness=666.0, horizon=Hp)
            solver.total_steps = 0
            st = [np.array(a) for a in env.get_state()["state"]]
            if w.get("v0"):
                st[1][:] = np.asarray(w["v0"])
            loss, grad = solver.forward(st, acts)
            x = env.simulator.get_x(env.simulator.cur)
            out[dtype] = (loss, np.array(grad), x)
            env.engine.close()
            del env, solver
        (l32, g32, x32), (l64, g64, x64) = out[self.args.dtype], out["float64"]
        gn = float(np.linalg.norm(g64))
        return {"what": f"{self.args.dtype} engine (benchmarked) vs float64 engine, same scene / state / actions, through Solver.forward",
                "grad_rel_err_f32_vs_f64": float(np.linalg.norm(g32 - g64) / max(gn, 1e-300)) if self.A else None,
                "grad_max_abs_err": float(np.abs(g32 - g64).max()) if self.A else None, "grad_norm": gn,
                "loss_rel_err": float(abs(l32 - l64) / abs(l64)), "loss": float(l64),
                "max_abs_dx": float(np.abs(x32 - x64).max()), "dx_cells": float(np.abs(x32 - x64).max() * self.env.simulator.n_grid),
                "n_substeps": int(Hp * S), "env_steps": int(Hp), "n_particles": int(N),
                "chain": "float64 engine vs float64 oracle: <= 1e-9 relative (tests/test_gpu_parity.py, small scenes); oracle pinned on the "
                         "reference's loss anchors, unpinned at the Taichi boundary for gradients (DESIGN.md 2)"}

    def run(self, primary=True):
        torch, dist, args, w = self.torch, self.dist, self.args, self.w
        env, eng, S, H, N, A = self.env, self.eng, self.S, self.H, self.N, self.A
        rank, world = self.rank, self.world
        steps = args.steps if primary else min(args.steps, 2)
        env.primitives.set_softness(666.0)
        for _ in range(max(args.warmup, 3)):
            self.episode_device()
        torch.cuda.synchronize()
        sampler = ClockSampler(self.local_rank)
        if rank == 0:
            sampler.start()
        l0 = eng.lib.plb_launch_count(eng.h)
        ms_dev = self.timed(self.episode_device, steps)
        launches = eng.lib.plb_launch_count(eng.h) - l0
        clocks = sampler.stop() if rank == 0 else None
        # per-kernel device times: one more episode of the same workload; the engine launches the kernel sequence of its
        # env-step graphs one by one with a CUDA-event pair around every kernel (same kernels, same order, same streams)
        eng.call("plb_profile_enable", 1)
        self.episode_device()
        kms = np.zeros(16)
        kcnt = (C.c_longlong * 16)()
        nk = eng.lib.plb_profile_read(eng.h, 16, self.capi.dptr(kms), kcnt)
        eng.call("plb_profile_enable", 0)

        # end-to-end through the public API
        self.episode_e2e()
        ms_e2e = self.timed(self.episode_e2e, steps)

        units_per_step = self.N_global * H * S
        value = units_per_step * steps / (ms_dev * 1e-3)
        e2e_value = units_per_step * steps / (ms_e2e * 1e-3)
        if self.slab:
            torch.cuda.synchronize()
        if rank != 0:
            self.close()
            return None

        # ---- roofline of the dominant kernel of the timed path + of the fused substep
        na = C.c_longlong()
        eng.call("plb_count_active", (S // 2) if self.ckpt is not None else (H // 2) * S, C.byref(na))
        n_active = int(na.value)
        peak, peak_src = measured_peak_gbs()
        names = [eng.lib.plb_kernel_name(i).decode() for i in range(nk)]
        per_kernel = {names[i]: {"launches": int(kcnt[i]), "total_ms": float(kms[i]), "avg_us": 1e3 * float(kms[i]) / max(int(kcnt[i]), 1)}
                      for i in range(nk) if kcnt[i] > 0}
        sub = {k: v for k, v in per_kernel.items() if k in KERNEL_BYTES}
        dom = max(sub, key=lambda k: sub[k]["total_ms"])
        bpp, bpn = KERNEL_BYTES[dom]
        sc = 2 if args.dtype == "float64" else 1
        alg_bytes = sc * (bpp * N + bpn * n_active)
        achieved = alg_bytes / (sub[dom]["avg_us"] * 1e-6) / 1e9
        # whole job: all ranks' particles and (approximately) all ranks' active nodes against world x the per-GPU peak
        fused_bytes = sc * (504 * self.N_global + 168 * n_active * world)
        fused_gbs = fused_bytes * (H * S * steps) / (ms_dev * 1e-3) / 1e9
        peak_job = peak * world
        kernel_ms_total = sum(v["total_ms"] for v in per_kernel.values())
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": NCU_TRAFFIC.get((self.wname, dom)) if args.dtype == "float32" else None,
                    "traffic_source": NCU_TRAFFIC_SOURCE.get(self.wname, "no ncu --set full capture at this workload (null)"),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                    "algorithmic_bytes_formula": f"{bpp} B x N particles + {bpn} B x N_active nodes (DESIGN.md 4)",
                    "avg_launch_us": sub[dom]["avg_us"], "share_of_kernel_time": sub[dom]["total_ms"] / max(kernel_ms_total, 1e-9),
                    "n_active_nodes": n_active,
                    "timing": "CUDA events around every kernel of one extra episode run inside bench.py right after the timed region: "
                              "the engine launches the kernel sequence of its env-step graphs one by one (same fused kernels, order and "
                              "streams); `value` itself is timed with the graphs replayed",
                    "fused_substep": {"algorithmic_bytes": fused_bytes, "achieved": fused_gbs, "frac": fused_gbs / peak_job, "peak": peak_job,
                                      "formula": "(504 N + 168 N_active) B per fwd+bwd substep (SURVEY.md 8d) x substeps / device time of `value`"},
                    "kernels": per_kernel}

        state_bytes = 24 * N * 8
        h2d = state_bytes + self.actions.nbytes + H * S * 2 * 8 * 8 * len(env.primitives)
        d2h = (H + 1) * 64 + self.grad_out.nbytes + 8
        senv = self.senv
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
                "config": {"workload": self.wname, "description": w["desc"], "n_particles": N, "n_grid": env.simulator.n_grid,
                           "substeps_per_env_step": S, "env_steps": H, "particle_substeps_per_step": units_per_step,
                           "parallelism": "single GPU" if world == 1 else (
                               f"{world} slabs along grid axis 0, {senv.halo_w}-plane halo zones; " +
                               ("active zone blocks pushed into the neighbour's inbox over NVLink peer memory (CUDA IPC) inside the "
                                "captured env-step graphs, once per substep fwd and once bwd" if senv.peer else
                                "zones summed over NCCL send/recv driven from the host, once per substep fwd and once bwd") +
                               f"; loss scalars / pose gradients all-reduced over NCCL; bounds {senv.bounds}"),
                           "n_particles_global": self.N_global,
                           "kernel_switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("PLB_")} or "defaults",
                           "l2": "inputs larger than L2: every substep reads a different trajectory frame "
                                 f"({(H * S + 1) * 96 * N / 1e9:.1f} GB trajectory per episode and GPU)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": ms_e2e / steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
        self.close()
        return line


def slab_parity(args, wname, rank, world, local_rank, dist):
    """Inside the N > 1 bench line (the driver's GPU-test box has one GPU): the slab-decomposed engine against the single-GPU
    engine on a small instance of the same scene (50k particles per GPU, one env step), float64 and the benchmarked dtype."""
    import torch
    from plasticinelab_b200.engine.sharded import ShardedEnv
    from plasticinelab_b200.engine.taichi_env import TaichiEnv
    from plasticinelab_b200.optimizer.solver import Solver
    w = dict(WORKLOADS[wname], n=50_000, horizon=1)
    out = {"what": f"{world}-slab engine vs the single-GPU engine, same scene at 50k particles per GPU, 1 env step, through the bench's episode calls"}
    for dtype in ("float64", args.dtype):
        cfg, S = build_cfg(w, world)
        mats = split_materials if w.get("materials") else None
        senv = ShardedEnv(cfg, dtype=dtype, device=local_rank, halo_w=int(os.environ.get("PLB_BENCH_HALO_W", w.get("halo_w", 8))), materials=mats)
        senv.env.loss.set_weights(10, 10, 1, False)
        A = senv.env.primitives.action_dim
        acts = actions_for(w, A, 1)
        senv.begin_episode(666.0)
        senv.step(acts[0])
        senv.compute_loss()
        grad = senv.backward()
        loss = senv.loss_value()
        senv.close()
        if rank == 0:
            ref = TaichiEnv(build_cfg(w, world)[0], dtype=dtype, device=local_rank)
            ref.initialize()
            if mats:
                ref.simulator.set_materials(*mats(ref.init_particles))
            ref.loss.set_weights(10, 10, 1, False)
            solver = Solver(ref, None, None, n_iters=1, softness=666.0, horizon=1)
            solver.total_steps = 0
            rloss, rgrad = solver.forward(ref.get_state()["state"], acts)
            ref.engine.close()
            key = "f64" if dtype == "float64" else "f32"
            out[f"loss_rel_{key}"] = float(abs(loss - rloss) / abs(rloss))
            if A:
                out[f"grad_rel_{key}"] = float(np.linalg.norm(grad - rgrad) / max(np.linalg.norm(rgrad), 1e-300))
        dist.barrier()
        torch.cuda.synchronize()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the extra single-GPU workloads of the N = 1 line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    also = args.workload is None and world == 1 and not args.no_also
    if args.workload is None:
        args.workload = DEFAULT_WORKLOAD           # the same workload for every N (weak scaling)
    w = WORKLOADS[args.workload]
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, w, rank)
        return

    import torch
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
        if w["scene"] not in ("slab", "block", "torus.yml"):
            raise SystemExit(f"workload {args.workload} has no slab decomposition; use slab1m, torus4m or block16m with --gpus > 1")
    torch.cuda.set_device(local_rank)

    job = Job(args, args.workload, rank, world, local_rank, dist)
    cfg, S, N = job.cfg, job.S, job.N
    line = job.run(primary=True)
    par = None
    if not args.no_parity:
        if world == 1:
            par = job.parity()
        else:
            try:
                par = slab_parity(args, args.workload, rank, world, local_rank, dist)
            except Exception as e:  # noqa: BLE001  (the throughput line is still worth printing; the failure is part of it)
                par = {"error": f"{type(e).__name__}: {e}"}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    if world == 1:
        line["parity"] = par
    else:
        line["slab_parity"] = par

    line["cpu_baseline"] = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        n_sub = 4 if N >= 1_000_000 else 20
        v, dt, kind, used = oracle_sample(cfg, w, S, n_sub, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                                "sample": f"{n_sub} fwd+bwd substeps of the same scene ({kind}; {dt:.1f} s)"}
    if also:
        # the other single-GPU configurations the contract names, same rules, compact records
        extra = {}
        for name in ("move1m", "move100k", "rope1m"):
            t0 = time.perf_counter()
            j = Job(args, name, 0, 1, local_rank, None)
            ln = j.run(primary=False)
            rec = {"value": ln["value"], "unit": UNIT, "ms_per_step": ln["ms_per_step"], "steps": ln["steps"], "e2e": ln["e2e"]["value"],
                   "config": {k: ln["config"][k] for k in ("description", "n_particles", "n_grid", "substeps_per_env_step", "env_steps")},
                   "gpu_launches": ln["gpu_launches"],
                   "roofline": {k: ln["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "avg_launch_us", "algorithmic_bytes_per_launch",
                                                                "share_of_kernel_time", "n_active_nodes", "traffic")},
                   "fused_substep_frac": ln["roofline"]["fused_substep"]["frac"]}
            if not args.no_parity:
                rec["parity"] = {k: v for k, v in j.parity().items() if k not in ("what", "chain")}
            rec["wall_s"] = time.perf_counter() - t0
            extra[name] = rec
        line["also"] = extra
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
