#!/bin/bash
# Round 2, first GPU call (run under gpurun): GPU test suite with the measured parity errors logged, the default bench line
# (slab1m + `also` workloads + parity + cpu baseline), the reference arm, the launch list of the default command and
# `ncu --set full` captures of the two fused particle kernels at slab1m / move1m / rope1m.  Outputs: gpurun_out/$TAG/.
set +e
TAG=${TAG:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
BENCH_FLAGS="--steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-also"

if [ "${SKIP_TESTS:-0}" != "1" ]; then
    stamp "pytest -m gpu"
    PLB_PARITY_LOG=$OUT/parity.jsonl timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
    stamp "-> exit $? $(tail -1 $OUT/pytest_gpu.log)"
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
    stamp "bench.py (defaults)"
    timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
    stamp "-> exit $?"
    stamp "bench.py --impl reference"
    timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
    stamp "-> exit $?"
fi

# prof_full WORKLOAD KERNEL-REGEX SKIP NAME
prof_full() {
    stamp "ncu --set full $1 $2 (skip $3)"
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 --launch-count 2 -f -o /tmp/full_$4 \
        python bench.py --workload $1 $BENCH_FLAGS > $OUT/full_$4.log 2>&1
    stamp "-> exit $?"
    ncu -i /tmp/full_$4.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/full_$4_raw.csv.gz
    ncu -i /tmp/full_$4.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/full_$4_source_sass.csv.gz
    ncu -i /tmp/full_$4.ncu-rep --page details 2>/dev/null | gzip > $OUT/full_$4_details.txt.gz
}
if [ "${SKIP_PROF:-0}" != "1" ]; then
    stamp "launch list (default workload)"
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 700 --csv --log-file $OUT/launches_slab1m.csv \
        python bench.py $BENCH_FLAGS > $OUT/launches_slab1m.log 2>&1
    stamp "-> exit $?"
    FWD=${FWD_KERNEL:-k_g2p_p2g_warp}
    BWD=${BWD_KERNEL:-k_p2g_bwd_g2p_bwd_warp}
    prof_full slab1m "$FWD" 508 slab1m_fwd
    prof_full slab1m "$BWD" 508 slab1m_bwd
    prof_full move1m "$FWD" 1180 move1m_fwd
    prof_full move1m "$BWD" 1180 move1m_bwd
    prof_full rope1m "$FWD" 508 rope1m_fwd
    prof_full rope1m "$BWD" 508 rope1m_bwd
fi
du -sh gpurun_out | tee -a $OUT/timeline.txt
stamp done
