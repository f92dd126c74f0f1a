#!/bin/bash
TAG=${1:-r01u}
OUT=gpurun_out; mkdir -p $OUT
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench (default flags)"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "rc=$?"; cut -c1-1500 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; cut -c1-400 $OUT/bench_ref_$TAG.json
echo "== trace of one pass (1M reads)"
EDLIB_B200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 2 --e2e-steps 2 --no-cpu-baseline --no-sweep-sample > $OUT/trace_$TAG.txt 2>&1; tail -30 $OUT/trace_$TAG.txt | cut -c1-160
echo "== configs"; timeout 900 python scripts/config_runs.py --pairs3 20000 --reads4 100000 2>/dev/null | tee $OUT/configs_$TAG.json | cut -c1-1500
