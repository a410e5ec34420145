#!/bin/bash
# Round 2: env-step re-sort (PLB_RESORT=1): variant agreement test, then A/B on a translating body and on the standard workloads.
set +e
TAG=${TAG:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
stamp "pytest variants"
timeout 200 python -m pytest tests/test_gpu_variants.py ${EXTRA_TESTS} -x -q > $OUT/pytest.log 2>&1
stamp "-> exit $? $(tail -1 $OUT/pytest.log)"
grep -E "AssertionError|Error|assert " $OUT/pytest.log | head -5 | tee -a $OUT/timeline.txt
for CFG in ${CONFIGS:-fly1m:resort:PLB_RESORT=1 fly1m:default: move100k:resort:PLB_RESORT=1 slab1m:resort:PLB_RESORT=1}; do
    WL=$(echo $CFG | cut -d: -f1); NAME=$(echo $CFG | cut -d: -f2); ENVS=$(echo $CFG | cut -d: -f3- | tr ',' ' ')
    stamp "bench $WL $NAME [$ENVS]"
    env $ENVS timeout 200 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-also > $OUT/bench_${WL}_$NAME.json 2> $OUT/bench_${WL}_$NAME.err
    stamp "-> exit $? $(python -c "
import json
d=json.loads([l for l in open('$OUT/bench_${WL}_$NAME.json') if l.startswith('{')][-1])
k=d['roofline']['kernels']
print('value %.4g e2e %.4g frac %.4f | ' % (d['value'], d['e2e']['value'], d['roofline']['fused_substep']['frac']) + ' '.join('%s=%.1f' % (n, v['avg_us']) for n, v in k.items() if n in ('g2p_p2g','p2g_bwd_g2p_bwd')) + ' | parity ' + json.dumps({k2: d['parity'][k2] for k2 in ('grad_rel_err_f32_vs_f64','loss_rel_err','dx_cells')}))
" 2>&1 | tail -1)"
    tail -2 $OUT/bench_${WL}_$NAME.err | cut -c1-300 >> $OUT/timeline.txt
done
stamp done
