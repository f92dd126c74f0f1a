#!/bin/bash
TAG=${1:-r01y}
OUT=gpurun_out; mkdir -p $OUT
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench (default flags)"; timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"; cut -c1-2400 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== ncu launch list of the bench command (1 step, 1 warm-up)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 140 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sweep-sample > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu rc=$?"
python scripts/step_traffic.py $OUT/launches_$TAG.csv 1000000 $OUT/step_traffic_$TAG.json
echo "== trace"; EDLIB_B200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 2 --e2e-steps 2 --no-cpu-baseline --no-sweep-sample > $OUT/trace_$TAG.txt 2>&1; grep "stage\|K1 groups" $OUT/trace_$TAG.txt | tail -5
echo "== stress"; EDLIB_B200_FILTER_MIN_TARGET=128 EDLIB_B200_K1_MIN_GROUP=4 timeout 300 python scripts/stress.py 0.5 2>&1 | tail -2 | tee $OUT/stress_$TAG.txt
