#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, a bench run, ncu launch list + one full capture.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [reads] [tag]
READS=${1:-262144}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke_$TAG.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu_$TAG.log
echo "== bench"; timeout 1200 python bench.py --reads $READS --steps 2 --warmup 3 --cpu-seconds 10 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --reads 32768 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full (k1)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 1 -c 1 -o $OUT/k1_$TAG -f \
    python bench.py --reads 32768 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu2 rc=$?"
ls -la $OUT
