#!/bin/bash
# A/B of kernel variants + full default bench.  Usage: bash scripts/gpu_ab.sh <tag>
TAG=${1:-r01b}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log
for V in 0 1; do
  echo "== variant FMA_SHIFT=$V"
  EDLIB_B200_K1_FMA_SHIFT=$V timeout 600 python bench.py --reads 262144 --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/ab_${TAG}_v$V.json 2> $OUT/ab_${TAG}_v$V.err
  python -c "import json;d=json.load(open('$OUT/ab_${TAG}_v$V.json'));print('value',round(d['value']),'ms/step',round(d['ms_per_step']),'k1ms',round(d['roofline']['kernel_ms']),'launches',d['gpu_launches'])" || tail -5 $OUT/ab_${TAG}_v$V.err
done
echo "== full bench (default flags)"
timeout 1500 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
