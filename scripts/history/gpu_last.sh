#!/bin/bash
TAG=${1:-r01z}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench (default flags)"; timeout 300 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"; cut -c1-2300 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== A/B: three seed levels"
EDLIB_B200_FILTER_SEED_LEVELS=3 timeout 200 python bench.py --no-cpu-baseline --no-sweep-sample 2>/dev/null | tee $OUT/bench_${TAG}_levels3.json | cut -c1-1500
