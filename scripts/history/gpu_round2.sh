#!/bin/bash
TAG=${1:-r01c}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu_$TAG.log
echo "== config runs"; timeout 1500 python scripts/config_runs.py > $OUT/configs_$TAG.json 2> $OUT/configs_$TAG.err; echo "rc=$?"; cat $OUT/configs_$TAG.json; tail -5 $OUT/configs_$TAG.err
echo "== bench 262144"; timeout 600 python bench.py --reads 262144 --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/b262k_$TAG.json 2> $OUT/b262k_$TAG.err; python -c "import json;d=json.load(open('$OUT/b262k_$TAG.json'));print('value',round(d['value']),'ms/step',round(d['ms_per_step']),'k1ms',round(d['roofline']['kernel_ms']),'launches',d['gpu_launches'])" || tail -5 $OUT/b262k_$TAG.err
