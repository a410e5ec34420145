#!/bin/bash
# One-off probe of the GPU box: is there any way to run the reference's Taichi 0.7.x there?  Answer is recorded in DESIGN.md §2.
OUT=gpurun_out/probe; mkdir -p $OUT
{
echo "== import taichi"; python -c "import taichi; print(taichi.__version__)" 2>&1 | tail -1
echo "== pip download taichi==0.7.14"; timeout 40 python -m pip download --no-deps -d /tmp/ti taichi==0.7.14 2>&1 | tail -2
echo "== pip download taichi (any), wheelhouse"; timeout 40 python -m pip download --no-index --find-links /opt/wheelhouse --no-deps -d /tmp/ti taichi 2>&1 | tail -2
echo "== baseline/_ref"; ls -la baseline/_ref 2>&1 | head
echo "== find taichi"; find / -iname "taichi*" -not -path "/proc/*" 2>/dev/null | head
echo "== network"; timeout 8 python - <<'PY'
import socket
try:
    socket.create_connection(("pypi.org", 443), timeout=5); print("network: reachable")
except Exception as e: print("network: unreachable:", e)
PY
echo "== nproc"; nproc; lscpu | grep -E "Model name|Socket|Core|Thread" 
} > $OUT/taichi_probe.txt 2>&1
cat $OUT/taichi_probe.txt
