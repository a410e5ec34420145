#!/bin/bash
# ncu --set full of the sparse grid kernels (forward operator, adjoint) at move100k, where they are a third of the substep.
set +e
TAG=${TAG:-r2grid}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
BENCH_FLAGS="--steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-also"
prof_full() {
    stamp "ncu --set full $1 $2 (skip $3)"
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 --launch-count 2 -f -o /tmp/full_$4 \
        python bench.py --workload $1 $BENCH_FLAGS > $OUT/full_$4.log 2>&1
    stamp "-> exit $?"
    ncu -i /tmp/full_$4.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/full_$4_raw.csv.gz
    ncu -i /tmp/full_$4.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/full_$4_source_sass.csv.gz
    ncu -i /tmp/full_$4.ncu-rep --page details 2>/dev/null | gzip > $OUT/full_$4_details.txt.gz
}
prof_full ${WL:-move100k} "k_grid_fwd_sparse" ${SKIP_FWD:-6000} grid_fwd
prof_full ${WL:-move100k} "k_grid_bwd_sparse_v2" ${SKIP_BWD:-5950} grid_bwd
stamp "launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LIST_SKIP:-28000} -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --workload ${WL:-move100k} $BENCH_FLAGS > $OUT/launches.log 2>&1
stamp "-> exit $?"
stamp done
