#!/bin/bash
TAG=${1:-tr}
OUT=gpurun_out; mkdir -p $OUT
EDLIB_B200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 2 --e2e-steps 2 --no-cpu-baseline --no-sweep-sample > $OUT/trace_$TAG.txt 2>&1; tail -34 $OUT/trace_$TAG.txt | cut -c1-200
