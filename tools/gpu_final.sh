#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): smoke, default bench lines (own arm + reference arm), 1M-particle line,
# ncu launch list of the default bench command.  Outputs: gpurun_out/final/.
set +e
OUT=gpurun_out/final
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
stamp smoke
timeout 120 python -c "import __graft_entry__ as e; e.smoke()" > $OUT/smoke.log 2>&1
stamp "-> exit $? $(tail -1 $OUT/smoke.log)"
stamp "bench.py (defaults: move100k, with cpu_baseline)"
timeout 300 python bench.py > $OUT/bench_move100k.json 2> $OUT/bench_move100k.err
stamp "-> exit $?"
stamp "bench.py --impl reference"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
stamp "-> exit $?"
stamp "bench.py --workload move1m"
timeout 200 python bench.py --workload move1m --no-cpu-baseline > $OUT/bench_move1m.json 2> $OUT/bench_move1m.err
stamp "-> exit $?"
if [ "${WITH_CKPT:-0}" = "1" ]; then
    stamp "bench.py --workload move1m_ckpt"
    timeout 200 python bench.py --workload move1m_ckpt --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_move1m_ckpt.json 2> $OUT/bench_move1m_ckpt.err
    stamp "-> exit $?"
fi
stamp "ncu launch list (move100k)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 6300 -c 700 --csv --log-file $OUT/launches_move100k.csv \
    python bench.py --workload move100k --steps 1 --warmup 1 --no-cpu-baseline > $OUT/launches_move100k.log 2>&1
stamp "-> exit $?"
du -sh gpurun_out | tee -a $OUT/timeline.txt
stamp done
