#!/bin/bash
# A/B of kernel switches on one B200 (run under gpurun).  CONFIGS="name:ENV=V,ENV2=V name2:..."  WORKLOADS="slab1m move100k"
# Outputs: gpurun_out/$TAG/bench_<workload>_<name>.json + one summary line each in timeline.txt
set +e
TAG=${TAG:-ab2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
QUICK="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-also"
if [ -n "$TESTS" ]; then
    stamp "pytest $TESTS"
    PLB_PARITY_LOG=$OUT/parity.jsonl timeout 900 python -m pytest $TESTS -x -q -s > $OUT/pytest.log 2>&1
    stamp "-> exit $? $(tail -1 $OUT/pytest.log)"
fi
for WL in ${WORKLOADS:-slab1m}; do
    for CFG in ${CONFIGS:-def:}; do
        NAME=${CFG%%:*}
        ENVS=$(echo "${CFG#*:}" | tr ',' ' ')
        stamp "bench $WL $NAME [$ENVS]"
        env $ENVS timeout 300 python bench.py --workload $WL $QUICK > $OUT/bench_${WL}_$NAME.json 2> $OUT/bench_${WL}_$NAME.err
        stamp "-> exit $? $(python -c "
import json,sys
d=json.load(open('$OUT/bench_${WL}_$NAME.json'))
k=d['roofline']['kernels']
print('value %.4g e2e %.4g fused-frac %.4f | ' % (d['value'], d['e2e']['value'], d['roofline']['fused_substep']['frac']) + ' '.join('%s=%.1f' % (n, v['avg_us']) for n, v in k.items()))
" 2>&1 | tail -1)"
    done
done
if [ -n "$PROF_WL" ]; then
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
    prof_full $PROF_WL "k_fwd_chunk" ${PROF_SKIP:-520} ${PROF_WL}_fwd
    prof_full $PROF_WL "k_bwd_chunk" ${PROF_SKIP:-520} ${PROF_WL}_bwd
fi
du -sh gpurun_out | tee -a $OUT/timeline.txt
stamp done
