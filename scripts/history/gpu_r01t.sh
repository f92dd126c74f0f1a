#!/bin/bash
# Evidence run: parity, default bench (both arms), launch list of the bench command, full ncu captures of the
# two dominant kernels (window sweep k1w, seed planning), DRAM traffic of every kernel of a step.
TAG=${1:-r01t}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench (default flags)"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"; cut -c1-1600 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== trace of one pass (1M reads)"
EDLIB_B200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 2 --e2e-steps 1 --no-cpu-baseline --no-sweep-sample > $OUT/trace_$TAG.txt 2>&1; tail -22 $OUT/trace_$TAG.txt | cut -c1-160
echo "== ncu launch list of the bench command (1 step, 1 warm-up)"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sweep-sample > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu --set full: k1w"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1w_kernel -s 0 -c 1 -o $OUT/k1w_$TAG -f \
    python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline --no-sweep-sample > $OUT/ncu_k1w_$TAG.log 2>&1; echo "ncu2 rc=$?"
echo "== ncu --set full: seed_plan"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seed_plan_kernel -s 0 -c 1 -o $OUT/seedplan_$TAG -f \
    python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline --no-sweep-sample > $OUT/ncu_seedplan_$TAG.log 2>&1; echo "ncu3 rc=$?"
ls -la $OUT | tail -12
