#!/bin/bash
# Final evidence of a round: smoke, parity, default bench, A/B of the seed-length rule, ncu launch list (time + DRAM
# bytes per launch) of the bench command, optional reference arm and stress.   Usage: bash scripts/gpu_final.sh <tag> [stress minutes]
TAG=${1:-r01x}
OUT=gpurun_out; mkdir -p $OUT
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench (default flags)"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"; cut -c1-2600 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== A/B: seed length with sigma^L >= n (slack 1)"
EDLIB_B200_FILTER_SEED_SLACK=1 timeout 600 python bench.py --no-cpu-baseline --no-sweep-sample 2>/dev/null | tee $OUT/bench_${TAG}_slack1.json | cut -c1-1400
echo "== ncu launch list of the bench command (1 step, 1 warm-up)"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sweep-sample > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu rc=$?"
python scripts/step_traffic.py $OUT/launches_$TAG.csv 1000000 $OUT/step_traffic_$TAG.json
if [ -n "$2" ]; then
  echo "== bench --impl reference"; timeout 900 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; cut -c1-300 $OUT/bench_ref_$TAG.json
  echo "== stress (filter forced on for small targets)"
  EDLIB_B200_FILTER_MIN_TARGET=128 EDLIB_B200_K1_MIN_GROUP=4 timeout 600 python scripts/stress.py $2 2>&1 | tail -2 | tee $OUT/stress_$TAG.txt
fi
