#!/bin/bash
TAG=${1:-r01h}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
echo "== full bench (default flags)"
EDLIB_B200_TRACE=1 timeout 1500 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; grep edlib_b200 $OUT/bench_$TAG.err | tail -16
echo "== reference arm"
timeout 900 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; cat $OUT/bench_ref_$TAG.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --reads 131072 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full (range kernel = first k1 launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 0 -c 1 -o $OUT/k1range_$TAG -f \
    python bench.py --reads 131072 --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu rc=$?"
echo "== configs"
timeout 900 python scripts/config_runs.py --pairs3 20000 --reads4 100000 > $OUT/configs_$TAG.json 2> $OUT/configs_$TAG.err; cat $OUT/configs_$TAG.json; tail -3 $OUT/configs_$TAG.err
