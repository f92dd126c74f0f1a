#!/bin/bash
TAG=${1:-r01e}
OUT=gpurun_out; mkdir -p $OUT
for V in 0 1; do
  echo "== FMA_SHIFT=$V filter on (1M reads, 1 step)"
  EDLIB_B200_K1_FMA_SHIFT=$V timeout 600 python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/v_${TAG}_$V.json 2> $OUT/v_${TAG}_$V.err
  python -c "import json;d=json.load(open('$OUT/v_${TAG}_$V.json'));print('value',round(d['value']),'ms/step',round(d['ms_per_step']),'k1ms',round(d['roofline']['kernel_ms']),'launches',d['gpu_launches'],d['filter'])" || tail -5 $OUT/v_${TAG}_$V.err
done
echo "== ncu full (range kernel = first k1 launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 0 -c 1 -o $OUT/k1range_$TAG -f \
    python bench.py --reads 131072 --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu rc=$?"
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
