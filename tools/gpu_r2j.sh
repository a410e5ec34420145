#!/bin/bash
# Round 2 final: ncu launch list of the default bench command and ncu --set full of the two fused kernels, final tree.
set +e
TAG=${TAG:-r2j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
QUICK="--steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-also"
stamp "ncu launch list (default workload, defaults incl. PDL)"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1400 --csv --log-file $OUT/launches.csv python bench.py $QUICK > $OUT/launches.log 2>&1
RC=$?
stamp "-> exit $RC"
if [ $RC -ne 0 ]; then
    stamp "again with PLB_PDL=0"
    PLB_PDL=0 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1400 --csv --log-file $OUT/launches.csv python bench.py $QUICK > $OUT/launches.log 2>&1
    stamp "-> exit $?"
fi
prof_full() {
    stamp "ncu --set full slab1m $1 (skip 520)"
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$1" --launch-skip 520 --launch-count 1 -f -o /tmp/full_$2 python bench.py $QUICK > $OUT/full_$2.log 2>&1
    stamp "-> exit $?"
    ncu -i /tmp/full_$2.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/full_$2_raw.csv.gz
    ncu -i /tmp/full_$2.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/full_$2_source_sass.csv.gz
    ncu -i /tmp/full_$2.ncu-rep --page details 2>/dev/null | gzip > $OUT/full_$2_details.txt.gz
}
prof_full "k_fwd_chunk" fwd
prof_full "k_p2g_bwd_g2p_bwd_warp" bwd
stamp done
