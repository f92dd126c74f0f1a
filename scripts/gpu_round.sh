#!/bin/bash
# One evidence run on a B200 (gpurun -- 'bash scripts/gpu_round.sh <tag> [stages]'); stages: smoke tests bench trace launches ncu ref stress memcheck
TAG=${1:-r02a}
STAGES=${2:-"smoke tests bench trace launches ncu"}
OUT=gpurun_out; mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has smoke; then echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; fi
if has tests; then echo "== pytest gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6; fi
if has bench; then
  echo "== bench (default flags)"; timeout 1200 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"; cut -c1-3000 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
fi
if has trace; then
  echo "== host-phase trace"
  EDLIB_B200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 2 --e2e-steps 2 --no-cpu-baseline --no-sweep-sample --no-extras > $OUT/trace_$TAG.txt 2>&1; grep -v "^{" $OUT/trace_$TAG.txt | tail -60 | cut -c1-160
fi
if has launches; then
  echo "== ncu launch list of the bench command (1 step, 1 warm-up)"
  timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sweep-sample --no-extras > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu rc=$?"
  python scripts/step_traffic.py $OUT/launches_$TAG.csv 1000000 $OUT/step_traffic_$TAG.json
fi
if has ncu; then
  for K in ${NCU_KERNELS:-k1w_kernel seed_plan_kernel}; do
    echo "== ncu --set full: $K"
    timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-1} -f -o $OUT/ncu_${TAG}_$K \
        python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-sweep-sample --no-extras > $OUT/ncu_${TAG}_$K.log 2>&1; echo "rc=$?"
    ncu -i $OUT/ncu_${TAG}_$K.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_${K}_raw.csv 2>/dev/null
    # gpurun brings back at most 64 MiB: a report that would break that is dropped (its CSV stays)
    if [ $(du -sm $OUT | cut -f1) -gt 48 ]; then ncu -i $OUT/ncu_${TAG}_$K.ncu-rep --page source --csv > $OUT/ncu_${TAG}_${K}_source.csv 2>/dev/null; rm -f $OUT/ncu_${TAG}_$K.ncu-rep; fi
  done
fi
if has cfgtrace; then
  echo "== host-phase trace of configs 3 and 4 (end to end)"
  EDLIB_B200_TRACE=1 timeout 900 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-sweep-sample --extras config4,config3 > $OUT/cfgtrace_$TAG.txt 2>&1
  grep -v "^{" $OUT/cfgtrace_$TAG.txt | tail -150 | cut -c1-170
fi
if has ncu2; then
  for KW in "band_kernel band" "lane_kernel path" "traceback_kernel path" "res_kernel path" "w_kernel long"; do
    set -- $KW; K=$1; W=$2
    echo "== ncu --set full: $K ($W)"
    timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K -s ${NCU2_SKIP:-1} -c 1 -f -o $OUT/ncu_${TAG}_$K \
        python scripts/ncu_targets.py $W > $OUT/ncu_${TAG}_$K.log 2>&1; echo "rc=$?"
    ncu -i $OUT/ncu_${TAG}_$K.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_${K}_raw.csv 2>/dev/null
    rm -f $OUT/ncu_${TAG}_$K.ncu-rep
  done
fi
if has longtrace; then
  echo "== host-phase trace of single long HW calls"
  EDLIB_B200_TRACE=1 timeout 600 python scripts/ncu_targets.py long1 2>&1 | tail -120 | cut -c1-170
  echo "== every long read, one call each"
  timeout 600 python scripts/ncu_targets.py long 2>&1 | tail -45
fi
if has h2d; then echo "== h2d microbenchmark"; ./scripts/microbench/h2d 2048 2>&1 | tail -12; fi
if has ref; then echo "== bench --impl reference"; timeout 900 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; cut -c1-400 $OUT/bench_ref_$TAG.json; fi
if has stress; then
  echo "== stress (filter forced on for small targets)"
  EDLIB_B200_FILTER_MIN_TARGET=128 EDLIB_B200_FILTER_MIN_LEVEL_READS=0 EDLIB_B200_K1_MIN_GROUP=4 EDLIB_B200_STREAM_MIN_PAIRS=64 EDLIB_B200_LONG_HW_MIN_TARGET=2000 \
      timeout 900 python scripts/stress.py ${STRESS_MIN:-3} 2>&1 | tail -2 | tee $OUT/stress_$TAG.txt
fi
if has memcheck; then
  # compute-sanitizer over a short stress plan of the real kernels (not run in round 2: no GPU minutes were left for it)
  for TOOL in memcheck racecheck synccheck; do
    echo "== compute-sanitizer --tool $TOOL"
    EDLIB_B200_FILTER_MIN_TARGET=128 EDLIB_B200_FILTER_MIN_LEVEL_READS=0 EDLIB_B200_K1_MIN_GROUP=4 EDLIB_B200_STREAM_MIN_PAIRS=64 EDLIB_B200_LONG_HW_MIN_TARGET=2000 \
        timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 python scripts/stress.py 0.2 > $OUT/sanitizer_${TOOL}_$TAG.txt 2>&1
    echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|stress ok" $OUT/sanitizer_${TOOL}_$TAG.txt | tail -3
  done
fi
