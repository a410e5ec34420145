#!/bin/bash
# ncu evidence for the current defaults (run under gpurun): launch list + one --set full capture of the hot kernels.
# Usage: tools/gpu_prof.sh [workload] [skip] [count]
set +e
WL=${1:-move100k}
SKIP=${2:-3840}
COUNT=${3:-24}
OUT=gpurun_out/prof
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
if [ -x tools/ubench/ffma2 ]; then stamp "ffma2 ubench"; tools/ubench/ffma2 | tee $OUT/ffma2.txt; fi
stamp "launch list ($WL)"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file $OUT/launches_$WL.csv \
    python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > $OUT/launches_$WL.log 2>&1
stamp "-> exit $?"
stamp "ncu --set full ($WL, skip $SKIP count $COUNT)"
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"k_g2p_p2g_warp|k_p2g_bwd_g2p_bwd_warp|k_grid_bwd_sparse|k_grid_fwd_sparse|k_grid_fwd_scan|k_restore_blocks" \
    --launch-skip $SKIP --launch-count $COUNT -f -o $OUT/full_$WL python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > $OUT/full_$WL.log 2>&1
stamp "-> exit $?"
ncu -i $OUT/full_$WL.ncu-rep --page raw --csv > $OUT/full_${WL}_raw.csv 2>/dev/null
ncu -i $OUT/full_$WL.ncu-rep --page source --csv --print-source sass > $OUT/full_${WL}_source.csv 2>/dev/null
ls -la $OUT | tee -a $OUT/timeline.txt
stamp done
