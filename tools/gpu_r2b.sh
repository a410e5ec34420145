#!/bin/bash
# Round 2: validate the chunked TMA-window kernels (plb_tile.cuh) on a B200, A/B them against the per-thread-gather kernels
# (PLB_TILE=0), capture them with ncu.  Outputs: gpurun_out/$TAG/.
set +e
TAG=${TAG:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
QUICK="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-also"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
    stamp "pytest variants + parity"
    PLB_PARITY_LOG=$OUT/parity.jsonl timeout 900 python -m pytest ${TESTS:-tests/test_gpu_variants.py tests/test_gpu_parity.py tests/test_gpu_large.py} -x -q -s > $OUT/pytest.log 2>&1
    stamp "-> exit $? $(tail -1 $OUT/pytest.log)"
fi
for WL in ${WORKLOADS:-slab1m move100k move1m rope1m}; do
    for TILE in ${TILES:-1 0}; do
        stamp "bench $WL PLB_TILE=$TILE"
        PLB_TILE=$TILE timeout 300 python bench.py --workload $WL $QUICK > $OUT/bench_${WL}_tile$TILE.json 2> $OUT/bench_${WL}_tile$TILE.err
        stamp "-> exit $? $(python -c "
import json,sys
d=json.load(open('$OUT/bench_${WL}_tile$TILE.json'))
k=d['roofline']['kernels']
print('value %.4g e2e %.4g fused-frac %.4f | ' % (d['value'], d['e2e']['value'], d['roofline']['fused_substep']['frac']) + ' '.join('%s=%.1f' % (n, v['avg_us']) for n, v in k.items()))
" 2>&1 | tail -1)"
    done
done
if [ "${SKIP_PROF:-0}" != "1" ]; then
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
    # (k_fwd_chunk: mode 3 = fused G2P+P2G; the first/last kernels of a graph are modes 2/1 -> skip counts include them)
    prof_full ${PROF_WL:-slab1m} "k_fwd_chunk" ${PROF_SKIP:-520} ${PROF_WL:-slab1m}_fwd
    prof_full ${PROF_WL:-slab1m} "k_bwd_chunk" ${PROF_SKIP:-520} ${PROF_WL:-slab1m}_bwd
fi
du -sh gpurun_out | tee -a $OUT/timeline.txt
stamp done
