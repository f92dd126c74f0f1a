#!/bin/bash
# Two-stage filter + add-with-carry shifts: parity, bench, A/B against the single 64-row stage.
TAG=${1:-r01n}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench (default flags)"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== A/B: stage 1 off (64-row stage only, K0=12)"
EDLIB_B200_FILTER_K1=0 EDLIB_B200_FILTER_K0=12 timeout 600 python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_${TAG}_stage2only.json
echo "== A/B: K1=9"
EDLIB_B200_FILTER_K1=9 timeout 600 python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_${TAG}_k1_9.json
echo "== trace of one pass (262144 reads)"
EDLIB_B200_TRACE=1 timeout 600 python bench.py --reads 262144 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline 2>&1 | grep -v "^\[edlib_b200\] [a-z:]* *[0-9.]* ms" | tail -30 > $OUT/trace_$TAG.txt; grep "filter stage" $OUT/trace_$TAG.txt | tail -4
