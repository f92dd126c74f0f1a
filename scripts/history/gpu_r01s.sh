#!/bin/bash
TAG=${1:-r01s}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench (default flags)"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== trace of one pass (1M reads)"
EDLIB_B200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline --no-sweep-sample > $OUT/trace_$TAG.txt 2>&1; tail -32 $OUT/trace_$TAG.txt | cut -c1-160
