#!/bin/bash
# One-shot GPU validation + A/B of the kernel variants (run under gpurun; logs go to gpurun_out/ab/).
set +e
OUT=gpurun_out/ab
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }

stamp "variants agree (one process, all switches)"
PLB_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q -s -p no:cacheprovider > $OUT/pytest_variants.log 2>&1
stamp "-> exit $? $(tail -1 $OUT/pytest_variants.log)"
grep "variants float" $OUT/pytest_variants.log | tee -a $OUT/timeline.txt

stamp "policy path parity (first GPU run)"
PLB_TEST_POLICY=1 timeout 200 python -m pytest tests/test_gpu_policy.py -m gpu -q -p no:cacheprovider > $OUT/pytest_policy.log 2>&1
stamp "-> exit $? $(tail -1 $OUT/pytest_policy.log)"

stamp "in-process bench A/B"
timeout ${AB_TIMEOUT:-420} python tools/ab_bench.py $OUT ${AB_BUDGET:-330} 2>&1 | tee -a $OUT/timeline.txt
stamp "RL vector env throughput (BASELINE configs[0])"
timeout 200 python tools/bench_rl.py --envs 1,4,16 > $OUT/bench_rl.jsonl 2> $OUT/bench_rl.err
stamp "-> exit $? $(tail -1 $OUT/bench_rl.jsonl | cut -c1-200)"
stamp "full-size property tests (1M particles)"
PLB_TEST_LARGE=1 PLB_TEST_UNVALIDATED=1 timeout 400 python -m pytest tests/test_gpu_large.py -m gpu -q -p no:cacheprovider > $OUT/pytest_large.log 2>&1
stamp "-> exit $? $(tail -1 $OUT/pytest_large.log)"
stamp done
