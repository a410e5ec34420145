#!/bin/bash
# Multi-GPU checks (run under gpurun --gpus N): 2-rank parity tests, then the bench at N ranks with the fused / per-substep halo.
set +e
N=${N:-2}
TAG=${TAG:-multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
if [ "${SKIP_TESTS:-0}" != "1" ]; then
    stamp "pytest tests/test_gpu_multi.py"
    timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > $OUT/pytest_multi.log 2>&1
    stamp "-> exit $? $(tail -1 $OUT/pytest_multi.log)"
    grep "slab parity" $OUT/pytest_multi.log | tee -a $OUT/timeline.txt
fi
PORT=29700
for WL in ${WORKLOADS:-slab1m}; do
    for CFG in ${CONFIGS:-fused: chain:PLB_SLAB_FUSED=0}; do
        NAME=${CFG%%:*}
        ENVS=$(echo "${CFG#*:}" | tr ',' ' ')
        PORT=$((PORT + 1))
        stamp "bench $WL x$N $NAME [$ENVS]"
        env $ENVS timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
            bench.py --gpus $N --workload $WL --steps ${STEPS:-3} --warmup 3 ${BENCH_EXTRA} > $OUT/bench_${WL}_n${N}_$NAME.json 2> $OUT/bench_${WL}_n${N}_$NAME.err
        stamp "-> exit $? $(python -c "
import json,sys
d=json.loads([l for l in open('$OUT/bench_${WL}_n${N}_$NAME.json') if l.startswith('{')][-1])
k=d['roofline']['kernels']
print('value %.4g e2e %.4g fused-frac %.4f ms %.2f | ' % (d['value'], d['e2e']['value'], d['roofline']['fused_substep']['frac'], d['ms_per_step']) + ' '.join('%s=%.1f' % (n, v['avg_us']) for n, v in k.items()) + ' | ' + json.dumps(d.get('slab_parity')))
" 2>&1 | tail -1)"
        tail -3 $OUT/bench_${WL}_n${N}_$NAME.err >> $OUT/timeline.txt
    done
done
if [ "${WITH_N1:-0}" = "1" ]; then
    stamp "bench slab1m x1"
    timeout 300 python bench.py --no-also --no-parity --no-cpu-baseline > $OUT/bench_slab1m_n1.json 2> $OUT/bench_slab1m_n1.err
    stamp "-> exit $? $(python -c "
import json
d=json.load(open('$OUT/bench_slab1m_n1.json'))
print('value %.4g e2e %.4g ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))")"
fi
stamp done
