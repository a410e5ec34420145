#!/bin/bash
# Round 2 evidence on one B200 (run under gpurun): smoke, full GPU suite, the driver's default bench line (+ reference arm),
# ncu launch list of the default bench command, ncu --set full of the fused particle kernels of the timed path.
# Outputs: gpurun_out/$TAG/.   Sections: SECTIONS="smoke tests bench ref list full" (default: all)
set +e
TAG=${TAG:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SECTIONS=${SECTIONS:-smoke tests bench ref list full}
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
has() { [[ " $SECTIONS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/gpu.txt 2>&1
if has smoke; then
    stamp smoke
    timeout 200 python -c "import __graft_entry__ as e; e.smoke()" > $OUT/smoke.log 2>&1
    stamp "-> exit $? $(tail -1 $OUT/smoke.log)"
fi
if has tests; then
    stamp "pytest -m gpu"
    PLB_PARITY_LOG=$OUT/parity.jsonl timeout 1200 python -m pytest tests -m gpu -x -q -rs --durations=15 > $OUT/pytest_gpu.log 2>&1
    stamp "-> exit $? $(tail -1 $OUT/pytest_gpu.log)"
fi
if has bench; then
    stamp "bench.py (defaults)"
    timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
    stamp "-> exit $? $(python - <<PY 2>&1 | tail -1
import json
d = json.loads([l for l in open('$OUT/bench_default.json') if l.startswith('{')][-1])
k = d['roofline']['kernels']
s = 'value %.4g e2e %.4g fused-frac %.4f dom %s frac %.3f | ' % (d['value'], d['e2e']['value'], d['roofline']['fused_substep']['frac'], d['roofline']['kernel'], d['roofline']['frac'])
s += ' '.join('%s=%.1f' % (n, v['avg_us']) for n, v in k.items())
for n, r in (d.get('also') or {}).items():
    s += ' || %s %.4g frac %.3f' % (n, r['value'], r['fused_substep_frac'])
print(s)
PY
)"
fi
if has ref; then
    stamp "bench.py --impl reference"
    timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
    stamp "-> exit $? $(cut -c1-200 $OUT/bench_reference.json)"
fi
QUICK="--steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-also"
if has list; then
    stamp "ncu launch list (default workload)"
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LIST_SKIP:-1400} -c ${LIST_COUNT:-800} --csv --log-file $OUT/launches.csv \
        python bench.py $QUICK > $OUT/launches.log 2>&1
    stamp "-> exit $?"
fi
if has full; then
    prof_full() {
        stamp "ncu --set full $1 $2 (skip $3)"
        timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 --launch-count 2 -f -o /tmp/full_$4 \
            python bench.py --workload $1 $QUICK > $OUT/full_$4.log 2>&1
        stamp "-> exit $?"
        ncu -i /tmp/full_$4.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/full_$4_raw.csv.gz
        ncu -i /tmp/full_$4.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/full_$4_source_sass.csv.gz
        ncu -i /tmp/full_$4.ncu-rep --page details 2>/dev/null | gzip > $OUT/full_$4_details.txt.gz
    }
    for WL in ${PROF_WLS:-slab1m}; do
        prof_full $WL "${FWD_REGEX:-k_fwd_chunk}" ${PROF_SKIP:-520} ${WL}_fwd
        prof_full $WL "${BWD_REGEX:-k_p2g_bwd_g2p_bwd_warp|k_bwd_chunk}" ${PROF_SKIP:-520} ${WL}_bwd
    done
fi
du -sh gpurun_out | tee -a $OUT/timeline.txt
stamp done
