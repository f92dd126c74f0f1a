#!/bin/bash
# Final evidence of the round: parity, default bench (both arms), launch list of the bench command,
# full ncu capture of the dominant kernel, the other configs.   Usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-r01m}
OUT=gpurun_out; mkdir -p $OUT
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "== bench (default flags)"; timeout 1500 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"; cat $OUT/bench_$TAG.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; cat $OUT/bench_ref_$TAG.json
echo "== ncu launch list of the bench command (1 step, 1 warm-up)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu --set full, prefix sweep (first k1 launch), 131072 reads"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 0 -c 1 -o $OUT/k1range_$TAG -f \
    python bench.py --reads 131072 --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu2 rc=$?"
echo "== configs"; timeout 900 python scripts/config_runs.py --pairs3 20000 --reads4 100000 2>/dev/null | tee $OUT/configs_$TAG.json
