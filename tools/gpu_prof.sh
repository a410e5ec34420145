#!/bin/bash
# ncu evidence for the current defaults (run under gpurun).  Keeps gpurun_out small: the .ncu-rep is exported to gzipped CSV
# pages (raw metrics, per-CUDA-line and per-SASS stall samples) and deleted.
# Usage: tools/gpu_prof.sh full|list [workload] [skip] [count] [kernel-regex]
#   e.g. the grid adjoint alone: tools/gpu_prof.sh full move100k 1950 4 "k_grid_bwd_sparse"
set +e
MODE=${1:-full}
WL=${2:-move100k}
SKIP=${3:-1897}
COUNT=${4:-6}
REGEX=${5:-"k_g2p_p2g_warp|k_p2g_bwd_g2p_bwd_warp"}
OUT=gpurun_out/prof
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
if [ "$MODE" = "list" ]; then
    stamp "launch list ($WL)"
    timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file $OUT/launches_$WL.csv \
        python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > $OUT/launches_$WL.log 2>&1
    stamp "-> exit $?"
else
    stamp "ncu --set full ($WL, skip $SKIP count $COUNT)"
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" \
        --launch-skip $SKIP --launch-count $COUNT -f -o /tmp/full_$WL python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > $OUT/full_$WL.log 2>&1
    stamp "-> exit $?"
    ncu -i /tmp/full_$WL.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/full_${WL}_raw.csv.gz
    ncu -i /tmp/full_$WL.ncu-rep --page source --csv --print-source cuda 2>/dev/null | gzip > $OUT/full_${WL}_source_cuda.csv.gz
    ncu -i /tmp/full_$WL.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/full_${WL}_source_sass.csv.gz
    ncu -i /tmp/full_$WL.ncu-rep --page details 2>/dev/null | gzip > $OUT/full_${WL}_details.txt.gz
fi
du -sh gpurun_out | tee -a $OUT/timeline.txt
ls -la $OUT | tee -a $OUT/timeline.txt
stamp done
