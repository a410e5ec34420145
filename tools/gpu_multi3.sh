#!/bin/bash
# Short multi-GPU check (run under gpurun --gpus N): a subset of the 2-rank parity tests, then bench lines for a list of
# configurations, every step under its own short timeout.  Outputs: gpurun_out/$TAG/.
set +e
N=${N:-2}
TAG=${TAG:-multi3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
if [ -n "$TESTS_K" ]; then
    stamp "pytest tests/test_gpu_multi.py -k '$TESTS_K'"
    timeout ${TEST_TIMEOUT:-150} python -m pytest tests/test_gpu_multi.py -x -q -s -k "$TESTS_K" > $OUT/pytest_multi.log 2>&1
    stamp "-> exit $? $(tail -1 $OUT/pytest_multi.log)"
    grep "slab parity" $OUT/pytest_multi.log | tee -a $OUT/timeline.txt
fi
PORT=29800
for CFG in ${CONFIGS:-default:slab1m:}; do
    NAME=$(echo $CFG | cut -d: -f1); WL=$(echo $CFG | cut -d: -f2); ENVS=$(echo $CFG | cut -d: -f3- | tr ',' ' ')
    PORT=$((PORT + 1))
    stamp "bench $WL x$N $NAME [$ENVS]"
    env $ENVS timeout ${BENCH_TIMEOUT:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $N --workload $WL --steps ${STEPS:-3} --warmup 3 ${BENCH_EXTRA} > $OUT/bench_${WL}_n${N}_$NAME.json 2> $OUT/bench_${WL}_n${N}_$NAME.err
    stamp "-> exit $? $(python -c "
import json,sys
d=json.loads([l for l in open('$OUT/bench_${WL}_n${N}_$NAME.json') if l.startswith('{')][-1])
k=d['roofline']['kernels']
print('value %.4g e2e %.4g fused-frac %.4f ms %.2f | ' % (d['value'], d['e2e']['value'], d['roofline']['fused_substep']['frac'], d['ms_per_step']) + ' '.join('%s=%.1f' % (n, v['avg_us']) for n, v in k.items()) + ' | ' + json.dumps(d.get('slab_parity'))[:300])
" 2>&1 | tail -1)"
    grep -h "Error\|error" $OUT/bench_${WL}_n${N}_$NAME.err | head -3 >> $OUT/timeline.txt
done
if [ "${WITH_N1:-0}" = "1" ]; then
    stamp "bench slab1m x1"
    timeout 120 python bench.py --no-also --no-parity --no-cpu-baseline > $OUT/bench_slab1m_n1.json 2> $OUT/bench_slab1m_n1.err
    stamp "-> exit $? $(python -c "
import json
d=json.load(open('$OUT/bench_slab1m_n1.json'))
print('value %.4g e2e %.4g ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))")"
fi
stamp done
