#!/usr/bin/env python
"""Benchmark of the hot path: differentiable MPM substeps, forward + backward, particle-substeps / s.

  python bench.py --gpus N --steps K --warmup W [--workload move100k] [--impl reference]

One "step" = one fwd+bwd episode of the workload (H env steps x S substeps forward with the loss after every env
step, then the full adjoint back to the action gradient) = N_particles * H * S particle-substeps.

  value  whole-job particle-substeps/s with the particle state already resident in HBM (device-timed, CUDA events);
  e2e    the same episode through the public API (`Solver.forward`): host float64 state -> device, per-env-step loss
         read-back, action gradient read-back;
  roofline      dominant kernel: algorithmic bytes per launch / mean CUDA-event duration, against MEASURED_PEAKS.json;
  cpu_baseline  the float64 oracle (a restatement of the reference's Taichi kernels; Taichi itself is not
                installable here) timed on this box's host cores on a bounded sample of the same workload.

--impl reference times that oracle alone (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-substeps/sec fwd+bwd"
UNIT = "particle-substeps/s"

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on for one GPU
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
    # multi-GPU weak scaling (north_star / BASELINE configs[4]-style): an elastic-plastic bar along the slab axis,
    # 1M particles and 0.1 of the domain (25.6 planes of 256) per GPU; --gpus N decomposes it into N slabs
    "slab1m": dict(scene="slab", n=1_000_000, quality=4, horizon=2,
                   desc="bar (0.109*N) x 0.1 x 0.1 on the ground pressed by two spheres, 1M particles per GPU, 256^3 grid, "
                        "2 env steps x 79 substeps, fwd+bwd"),
    "rope1m": dict(scene="rope.yml", n=1_000_000, quality=4, horizon=2,
                   desc="Rope-v1 geometry, 1M particles, 256^3 grid, 2 env steps x 79 substeps, fwd+bwd"),
}


def build_cfg(w, world=1):
    from plasticinelab_b200.envs.scene import load_variants
    from plasticinelab_b200 import _capi
    if w["scene"] == "slab":
        from plasticinelab_b200.config import load_dict
        L = 0.109375 * world          # 28 planes of 256 per GPU: slab boundaries fall on 4-plane block boundaries
        tree = dict(SIMULATOR=dict(quality=w["quality"], yield_stress=50.0, ground_friction=0.3),
                    SHAPES=[dict(shape="box", width=(L, 0.1, 0.1), init_pos=(0.5, 0.06, 0.5), n_particles=w["n"] * world)],
                    PRIMITIVES=[dict(shape="Sphere", radius=0.03, init_pos=(0.5 - 0.3 * L, 0.145, 0.5), friction=0.9,
                                     action=dict(dim=3, scale=(0.01, 0.01, 0.01))),
                                dict(shape="Sphere", radius=0.03, init_pos=(0.5 + 0.3 * L, 0.145, 0.5), friction=0.9,
                                     action=dict(dim=3, scale=(0.01, 0.01, 0.01)))])
        cfg = load_dict(tree)
        cfg.ENV.loss.target_path = "envs/assets/Rope3D-v1.npy"
    else:
        cfg = load_variants(w["scene"], 1)
        cfg.SIMULATOR.quality = w["quality"]
        cfg.SHAPES[0]["n_particles"] = w["n"]
    S = _capi.sim_constants(dict(cfg.SIMULATOR))["substeps"]
    cfg.SIMULATOR.max_steps = (S + 1 + w["horizon"] + 2) if w.get("checkpoint") else (w["horizon"] * S + 2)
    return cfg, S


def actions_for(w, A):
    a = np.random.RandomState(0).uniform(-0.01, 0.01, (w["horizon"], A))
    if w["scene"] == "slab":
        a[:, 1::3] = -0.5          # press the spheres into the bar
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


# algorithmic bytes per launch of each kernel: (bytes per particle, bytes per active node), float32 scalars
# (DESIGN.md "Kernels"; the sums are 204 N + 56 A forward and the engine's own 420 N + 128 A backward)
KERNEL_BYTES = {
    "p2g": (132, 16), "grid_fwd": (0, 28), "g2p": (72, 12), "p2g_recompute": (96, 16), "grid_fwd_recompute": (0, 28),
    "g2p_bwd": (84, 24), "grid_bwd": (0, 44), "p2g_bwd": (240, 16),
}


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the `ncu --set full` capture of the final kernels
# at 1M particles / 128^3 (profiles/r1_final_move1m_ncu_full_summary.md); reported as `roofline.traffic` only for
# workloads of that size, null otherwise.
NCU_TRAFFIC_1M = {"p2g": 111.8e6, "g2p": 20.9e6, "g2p_bwd": 80.7e6, "p2g_bwd": 206.8e6}


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
                f"C/OpenMP float64 port of the reference kernels, {port.threads} of {threads} threads (fastest of a thread-count sweep)", port.threads)
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
    return len(x0) * n_sub / dt, dt, f"float64 torch-CPU oracle, {threads} threads", threads


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
    sample = f"{n_sub} fwd+bwd substeps of the workload scene per step ({kind})"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": w["desc"], "n_particles": w["n"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:      # one GPU: BASELINE configs[1]; several GPUs: the slab-decomposed weak-scaling bar
        args.workload = "move100k" if world == 1 else "slab1m"
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
    from plasticinelab_b200 import _capi
    from plasticinelab_b200.engine.taichi_env import TaichiEnv
    from plasticinelab_b200.optimizer.solver import Solver

    slab = world > 1 and w["scene"] == "slab"
    cfg, S = build_cfg(w, world if slab else 1)
    torch.cuda.set_device(local_rank)
    senv = None
    n1_reference = None
    if slab:
        # scaling reference: the per-GPU share of this workload on ONE GPU (rank 0, regular graph path), same process
        if rank == 0:
            cfg1, _ = build_cfg(w, 1)
            env1 = TaichiEnv(cfg1, dtype=args.dtype, device=local_rank)
            env1.initialize()
            env1.loss.set_weights(10, 10, 1, False)
            a1 = actions_for(w, env1.primitives.action_dim)
            g1 = np.zeros((w["horizon"], env1.primitives.action_dim))
            e1 = env1.engine

            def episode1():          # device-resident episode, identical to episode_device() below
                env1.simulator.cur = 0
                env1._is_copy = False
                for p in env1.primitives:
                    p.set_state(0, p.init_state)
                e1.call("plb_zero_grads")
                for i in range(w["horizon"]):
                    e1.call("plb_set_action", i, S, _capi.dptr(np.ascontiguousarray(a1[i])), a1.shape[1])
                    e1.call("plb_kinematics", i * S, S)
                    e1.call("plb_step_fwd", i * S, i * S, S)
                    e1.call("plb_loss_fwd", (i + 1) * S, (i + 1) * S, None)
                for i in reversed(range(w["horizon"])):
                    e1.call("plb_loss_bwd", (i + 1) * S, (i + 1) * S)
                    e1.call("plb_step_bwd", i * S, i * S, S)
                e1.call("plb_get_action_grad", w["horizon"], S, _capi.dptr(g1))

            env1.primitives.set_softness(666.0)
            for _ in range(3):
                episode1()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                episode1()
            torch.cuda.synchronize()
            n1_reference = {"n_gpus": 1, "value": env1.n_particles * w["horizon"] * S * args.steps / (time.perf_counter() - t0), "unit": UNIT,
                            "how": "the per-GPU share of this workload on one GPU (rank 0, regular single-GPU path, state resident in HBM, "
                                   "same episode structure), measured in this process before the multi-GPU run"}
            env1.engine.close()
            del env1
        dist.barrier()
        from plasticinelab_b200.engine.sharded import ShardedEnv
        senv = ShardedEnv(cfg, dtype=args.dtype, device=local_rank, halo_w=8)
        env = senv.env
    else:
        env = TaichiEnv(cfg, dtype=args.dtype, device=local_rank, max_prim_frames=w["horizon"] * S + 2)
        env.initialize()
    env.loss.set_weights(10, 10, 1, False)
    eng = env.engine
    ckpt = None
    if w.get("checkpoint"):
        from plasticinelab_b200.engine.checkpoint import CheckpointedEpisode
        ckpt = CheckpointedEpisode(env, w["horizon"])
    eng.call("plb_set_stream", C.c_void_p(torch.cuda.current_stream().cuda_stream))
    N, H = env.n_particles, w["horizon"]
    N_global = senv.n_global if slab else N * world
    A = env.primitives.action_dim
    actions = actions_for(w, A)
    pinned_actions = torch.from_numpy(actions).pin_memory()
    # e2e inputs live in pinned host memory (particle state as float64 like the reference's get_state(), actions)
    host_state = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in env.get_state()["state"]]
    solver = Solver(env, None, None, n_iters=1, softness=666.0, horizon=H)
    solver.total_steps = 0
    grad_out = np.zeros((H, max(A, 1)))

    def episode_slab(load_state=False):
        if load_state:
            env.simulator.set_state(0, host_state)
        senv.begin_episode(666.0)
        a = pinned_actions.numpy()
        for i in range(H):
            senv.step(a[i])
            senv.compute_loss(sync=load_state)
        grad_out[:, :A] = senv.backward()

    def episode_device():
        """state resident in HBM (frame 0), no host read-back except the final action gradient"""
        if slab:
            return episode_slab(False)
        if ckpt is not None:
            grad_out[:, :A] = ckpt.forward_backward(pinned_actions.numpy())[1]
            return
        env.simulator.cur = 0
        env._is_copy = False
        for p in env.primitives:
            p.set_state(0, p.init_state)
        eng.call("plb_zero_grads")
        a = pinned_actions.numpy()
        for i in range(H):
            eng.call("plb_set_action", i, S, _capi.dptr(np.ascontiguousarray(a[i])), A)
            eng.call("plb_kinematics", i * S, S)
            eng.call("plb_step_fwd", i * S, i * S, S)
            eng.call("plb_loss_fwd", (i + 1) * S, (i + 1) * S, None)
        for i in reversed(range(H)):
            eng.call("plb_loss_bwd", (i + 1) * S, (i + 1) * S)
            eng.call("plb_step_bwd", i * S, i * S, S)
        eng.call("plb_get_action_grad", H, S, _capi.dptr(grad_out))

    def episode_e2e():
        if slab:
            return episode_slab(True)
        if ckpt is not None:
            env.set_state(host_state, 666.0, False)
            return ckpt.forward_backward(pinned_actions.numpy(), sync_losses=True)
        return solver.forward(host_state, pinned_actions.numpy())

    def timed(fn, k):
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

    env.primitives.set_softness(666.0)
    for _ in range(args.warmup):
        episode_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.lib.plb_launch_count(eng.h)
    ms_dev = timed(episode_device, args.steps)
    launches = eng.lib.plb_launch_count(eng.h) - l0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel device times: one more episode of the same workload with a CUDA-event pair around every launch
    # (the engine then launches kernel by kernel instead of replaying its per-env-step CUDA graphs)
    eng.call("plb_profile_enable", 1)
    episode_device()
    kms = np.zeros(16)
    kcnt = (C.c_longlong * 16)()
    nk = eng.lib.plb_profile_read(eng.h, 16, _capi.dptr(kms), kcnt)
    eng.call("plb_profile_enable", 0)

    # end-to-end through the public API
    episode_e2e()
    ms_e2e = timed(episode_e2e, args.steps)

    units_per_step = N_global * H * S if slab else N * H * S
    mult = 1 if slab else world
    value = mult * units_per_step * args.steps / (ms_dev * 1e-3)
    e2e_value = mult * units_per_step * args.steps / (ms_e2e * 1e-3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + of the fused substep
    na = C.c_longlong()
    eng.call("plb_count_active", (S // 2) if ckpt is not None else (H // 2) * S, C.byref(na))
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
    fused_bytes = sc * (504 * (N_global if slab else N * world) + 168 * n_active * world)
    fused_gbs = fused_bytes * (H * S * args.steps) / (ms_dev * 1e-3) / 1e9
    peak_job = peak * world
    kernel_ms_total = sum(v["total_ms"] for v in per_kernel.values())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (NCU_TRAFFIC_1M.get(dom) if (N == 1_000_000 and args.dtype == "float32") else None),
                "traffic_source": "ncu --set full capture at 1M particles, profiles/r1_final_move1m_ncu_full_summary.md (null for other sizes)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_us": sub[dom]["avg_us"], "share_of_kernel_time": sub[dom]["total_ms"] / max(kernel_ms_total, 1e-9),
                "n_active_nodes": n_active,
                "timing": "CUDA events around every launch of one extra episode run inside bench.py right after the timed "
                          "region (graphs off for that episode); `value` itself is timed with graphs on",
                "fused_substep": {"algorithmic_bytes": fused_bytes, "achieved": fused_gbs, "frac": fused_gbs / peak_job, "peak": peak_job,
                                  "formula": "(504 N + 168 N_active) B per fwd+bwd substep (SURVEY.md 8d)"},
                "kernels": per_kernel}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        n_sub = 4 if N >= 1_000_000 else 20
        v, dt, kind, used = oracle_sample(cfg, w, S, n_sub, threads)
        cpu = {"value": v, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"{n_sub} fwd+bwd substeps of the same scene ({kind}; {dt:.1f} s)"}

    state_bytes = 24 * N * 8
    h2d = state_bytes + actions.nbytes + H * S * 2 * 8 * 8 * len(env.primitives)
    d2h = (H + 1) * 64 + grad_out.nbytes + 8
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": w["desc"], "n_particles": N, "n_grid": env.simulator.n_grid,
                       "substeps_per_env_step": S, "env_steps": H, "particle_substeps_per_step": units_per_step,
                       "parallelism": "single GPU" if world == 1 else (
                           (f"{world} slabs along grid axis 0, {senv.halo_w}-plane halo zones; " +
                            ("active zone blocks pushed into the neighbour's inbox over NVLink peer memory (CUDA IPC) inside the "
                             "captured env-step graphs, once per substep fwd and once bwd" if senv.peer else
                             "zones summed over NCCL send/recv driven from the host, once per substep fwd and once bwd") +
                            f"; loss scalars / pose gradients all-reduced over NCCL; bounds {senv.bounds}") if slab
                           else f"{world} independent replicas (one env per GPU)"),
                       "n_particles_global": N_global, "scaling_reference": n1_reference,
                       "kernel_switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("PLB_")} or "defaults",
                       "l2": "inputs larger than L2: every substep reads a different trajectory frame "
                             f"({(H * S + 1) * 96 * N / 1e9:.1f} GB trajectory per episode)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
