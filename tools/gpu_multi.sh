#!/bin/bash
# 2-GPU checks (run under gpurun --gpus 2): slab parity vs the single-GPU engine, then one short weak-scaling bench line.
set +e
OUT=gpurun_out/multi
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $OUT/timeline.txt; }
stamp "slab parity (2 ranks)"
timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > $OUT/pytest_multi.log 2>&1
stamp "-> exit $? $(tail -1 $OUT/pytest_multi.log)"
stamp "bench --gpus 2"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 2 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
stamp "-> exit $?"
stamp done
