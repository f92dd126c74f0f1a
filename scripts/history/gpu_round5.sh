#!/bin/bash
TAG=${1:-r01f}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
for V in 0 3; do
  echo "== SHIFT=$V filter on (1M reads, 1 step)"
  EDLIB_B200_K1_SHIFT=$V timeout 600 python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/v_${TAG}_$V.json 2> $OUT/v_${TAG}_$V.err
  python -c "import json;d=json.load(open('$OUT/v_${TAG}_$V.json'));print('value',round(d['value']),'ms/step',round(d['ms_per_step']),'k1ms',round(d['roofline']['kernel_ms']),'launches',d['gpu_launches'],d['filter'])" || tail -5 $OUT/v_${TAG}_$V.err
done
echo "== trace (1M, e2e)"; EDLIB_B200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $OUT/t_$TAG.json 2> $OUT/t_$TAG.err; tail -40 $OUT/t_$TAG.err
