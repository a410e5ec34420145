#!/bin/bash
# Round 2: variants/parity tests with the new switches, moving-body A/B of the window origins, RL vector-env throughput,
# launch list at move100k, ncu --set full of the fused kernels at rope1m (BASELINE configs[2]).  Outputs: gpurun_out/$TAG/.
set +e
TAG=${TAG:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
QUICK="--steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-also"
stamp "pytest variants + parity"
PLB_PARITY_LOG=$OUT/parity.jsonl timeout 300 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -x -q > $OUT/pytest.log 2>&1
stamp "-> exit $? $(tail -1 $OUT/pytest.log)"
for CFG in follow: fixed:PLB_WINDOW_FOLLOW=0; do
    NAME=${CFG%%:*}; ENVS=$(echo "${CFG#*:}" | tr ',' ' ')
    stamp "bench fly1m $NAME [$ENVS]"
    env $ENVS timeout 200 python bench.py --workload fly1m --steps 3 --warmup 3 --no-cpu-baseline --no-also > $OUT/bench_fly1m_$NAME.json 2> $OUT/bench_fly1m_$NAME.err
    stamp "-> exit $? $(python -c "
import json
d=json.loads([l for l in open('$OUT/bench_fly1m_$NAME.json') if l.startswith('{')][-1])
k=d['roofline']['kernels']
print('value %.4g frac %.4f | ' % (d['value'], d['roofline']['fused_substep']['frac']) + ' '.join('%s=%.1f' % (n, v['avg_us']) for n, v in k.items() if n in ('g2p_p2g','p2g_bwd_g2p_bwd')) + ' | parity ' + json.dumps({k2: d['parity'][k2] for k2 in ('grad_rel_err_f32_vs_f64','loss_rel_err','dx_cells')}))
" 2>&1 | tail -1)"
done
stamp "bench_rl"
timeout 240 python tools/bench_rl.py --envs 1,8,32 > $OUT/bench_rl.jsonl 2> $OUT/bench_rl.err
stamp "-> exit $? $(cut -c1-400 $OUT/bench_rl.jsonl | tr '\n' ' ' | cut -c1-900)"
stamp "ncu launch list move100k (PLB_PDL=0: kernels are serialised under ncu anyway)"
PLB_PDL=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LIST_SKIP:-40000} -c ${LIST_COUNT:-1200} --csv --log-file $OUT/launches_move100k.csv \
    python bench.py --workload move100k --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-also > $OUT/launches_move100k.log 2>&1
stamp "-> exit $?"
prof_full() {
    stamp "ncu --set full $1 $2 (skip $3)"
    PLB_PDL=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 --launch-count 1 -f -o /tmp/full_$4 \
        python bench.py --workload $1 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-also > $OUT/full_$4.log 2>&1
    stamp "-> exit $?"
    ncu -i /tmp/full_$4.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/full_$4_raw.csv.gz
    ncu -i /tmp/full_$4.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/full_$4_source_sass.csv.gz
    ncu -i /tmp/full_$4.ncu-rep --page details 2>/dev/null | gzip > $OUT/full_$4_details.txt.gz
}
prof_full rope1m "k_fwd_chunk" 520 rope1m_fwd
prof_full rope1m "k_p2g_bwd_g2p_bwd_warp" 520 rope1m_bwd
stamp done
