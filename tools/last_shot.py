#!/usr/bin/env python
"""One short GPU process: parity of the register-form grid adjoint (PLB_GRID_BWD_V2=1) against the float64 oracle, then an
in-process A/B bench of it.  Everything in one interpreter (one torch / CUDA start-up)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.makedirs(os.path.join(ROOT, "gpurun_out", "last"), exist_ok=True)
t0 = time.time()
os.environ["PLB_GRID_BWD_V2"] = "1"
import pytest  # noqa: E402

rc = pytest.main(["-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_gpu_parity.py"),
                  "-k", "substep_parity or episode_loss or rotating"])
with open(os.path.join(ROOT, "gpurun_out", "last", "parity_v2.txt"), "w") as f:
    f.write(f"PLB_GRID_BWD_V2=1 parity subset exit code {int(rc)} after {time.time() - t0:.1f} s\n")
print(f"[last] parity exit {int(rc)} at {time.time() - t0:.1f}s", flush=True)
sys.argv = ["ab_bench.py", os.path.join(ROOT, "gpurun_out", "last"), "40"]
import ab_bench  # noqa: E402  (tools/ is on sys.path via this file's directory)

ab_bench.main()
