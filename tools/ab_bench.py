#!/usr/bin/env python
"""In-process A/B of the engine's kernel variants (environment switches read at plb_create): runs bench.main() once per
variant in ONE python process so the torch / CUDA start-up is paid once.  Usage: tools/ab_bench.py OUTDIR BUDGET_S"""
import contextlib
import io
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEYS = ["PLB_ENV_LIST", "PLB_SVD_STORE", "PLB_FLUSH_PAIRS", "PLB_GRID_BWD_V2", "PLB_FLUSH_RUNS", "PLB_BWD_OVERLAP", "PLB_GRID_SCAN", "PLB_FWD_PLANE", "PLB_FWD_MINB", "PLB_BWD_PLANE", "PLB_BWD_MINB", "PLB_CTA"]
VARIANTS = [     # the queue for the next GPU session: switches built and CPU-emulated after the round-1 GPU budget was spent
    ("def100k", "move100k", dict()),
    ("svd100k", "move100k", dict(PLB_SVD_STORE=1)),
    ("svd4_100k", "move100k", dict(PLB_SVD_STORE=1, PLB_BWD_MINB=4)),
    ("pairs100k", "move100k", dict(PLB_FLUSH_PAIRS=1)),
    ("svd4pairs", "move100k", dict(PLB_SVD_STORE=1, PLB_BWD_MINB=4, PLB_FLUSH_PAIRS=1)),
    ("envlist100k", "move100k", dict(PLB_ENV_LIST=1)),
    ("all100k", "move100k", dict(PLB_ENV_LIST=1, PLB_SVD_STORE=1, PLB_BWD_MINB=4, PLB_FLUSH_PAIRS=1)),
    ("def1m", "move1m", dict()),
    ("svd1m", "move1m", dict(PLB_SVD_STORE=1)),
    ("svd4_1m", "move1m", dict(PLB_SVD_STORE=1, PLB_BWD_MINB=4)),
    ("pairs1m", "move1m", dict(PLB_FLUSH_PAIRS=1)),
    ("all1m", "move1m", dict(PLB_ENV_LIST=1, PLB_SVD_STORE=1, PLB_BWD_MINB=4)),
]


def main():
    out = sys.argv[1]
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 1e9
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    os.makedirs(out, exist_ok=True)
    t0 = time.time()
    for name, wl, env in VARIANTS:
        if only and name not in only:
            continue
        if time.time() - t0 > budget:
            print(f"[ab] budget exhausted before {name}", flush=True)
            break
        for k in KEYS:
            os.environ.pop(k, None)
        for k, v in env.items():
            os.environ[k] = str(v)
        sys.argv = ["bench.py", "--workload", wl, "--steps", "2", "--warmup", "3", "--no-cpu-baseline"]
        buf = io.StringIO()
        t1 = time.time()
        try:
            with contextlib.redirect_stdout(buf):
                bench.main()
            line = buf.getvalue().strip().splitlines()[-1]
            d = json.loads(line)
            d["variant"] = {"name": name, "env": env}
            with open(os.path.join(out, f"bench_{name}.json"), "w") as f:
                f.write(json.dumps(d) + "\n")
            k = d["roofline"]["kernels"]
            print(f"[ab {time.time() - t0:6.1f}s] {name:10s} {wl:8s} value {d['value']:.4g} e2e {d['e2e']['value']:.4g} ms/step {d['ms_per_step']:.2f} "
                  f"fused-frac {d['roofline']['fused_substep']['frac']:.4f} launches {d['gpu_launches']} ({time.time() - t1:.1f}s) | "
                  + " ".join(f"{n}={v['avg_us']:.1f}" for n, v in k.items()), flush=True)
        except Exception:  # noqa: BLE001
            print(f"[ab] {name} FAILED\n{traceback.format_exc()}\n{buf.getvalue()[-2000:]}", flush=True)


if __name__ == "__main__":
    main()
