#!/bin/bash
# Round 2: A/B of the programmatic-dependent-launch chain (PLB_PDL) and the run flush of the per-warp kernels (PLB_FLUSH_MODE),
# variants test first.  Outputs: gpurun_out/$TAG/.
set +e
TAG=${TAG:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
QUICK="--steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-also"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
    stamp "pytest variants + parity"
    PLB_PARITY_LOG=$OUT/parity.jsonl timeout 900 python -m pytest ${TESTS:-tests/test_gpu_variants.py tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_policy.py} -x -q > $OUT/pytest.log 2>&1
    stamp "-> exit $? $(tail -1 $OUT/pytest.log)"
fi
for WL in ${WORKLOADS:-slab1m move100k}; do
    for CFG in ${CONFIGS:-default: nopdl:PLB_PDL=0 groups:PLB_FLUSH_MODE=0 nopdl_groups:PLB_PDL=0,PLB_FLUSH_MODE=0}; do
        NAME=${CFG%%:*}
        ENVS=$(echo "${CFG#*:}" | tr ',' ' ')
        stamp "bench $WL $NAME [$ENVS]"
        env $ENVS timeout 300 python bench.py --workload $WL $QUICK > $OUT/bench_${WL}_$NAME.json 2> $OUT/bench_${WL}_$NAME.err
        stamp "-> exit $? $(python -c "
import json,sys
d=json.loads([l for l in open('$OUT/bench_${WL}_$NAME.json') if l.startswith('{')][-1])
k=d['roofline']['kernels']
print('value %.4g e2e %.4g fused-frac %.4f ms %.2f | ' % (d['value'], d['e2e']['value'], d['roofline']['fused_substep']['frac'], d['ms_per_step']) + ' '.join('%s=%.1f' % (n, v['avg_us']) for n, v in k.items()))
" 2>&1 | tail -1)"
        tail -2 $OUT/bench_${WL}_$NAME.err >> $OUT/timeline.txt
    done
done
stamp done
