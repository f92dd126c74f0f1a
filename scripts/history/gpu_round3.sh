#!/bin/bash
TAG=${1:-r01d}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu_$TAG.log
for K0 in 0 12; do
  echo "== filter K0=$K0 (262144 reads)"
  EDLIB_B200_FILTER_K0=$K0 timeout 600 python bench.py --reads 262144 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $OUT/f_${TAG}_k$K0.json 2> $OUT/f_${TAG}_k$K0.err
  python -c "import json;d=json.load(open('$OUT/f_${TAG}_k$K0.json'));print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step']),'k1ms',round(d['roofline']['kernel_ms']),'kernel_ms',round(d['kernel_ms_per_step']),'launches',d['gpu_launches'],d['filter'])" || tail -5 $OUT/f_${TAG}_k$K0.err
done
echo "== full bench (default flags)"
timeout 1500 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --reads 65536 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full (k1 range kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 2 -c 1 -o $OUT/k1range_$TAG -f \
    python bench.py --reads 65536 --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu2 rc=$?"
echo "== single-call latency / configs"
timeout 900 python scripts/config_runs.py --pairs3 20000 --reads4 100000 > $OUT/configs_$TAG.json 2> $OUT/configs_$TAG.err; cat $OUT/configs_$TAG.json; tail -3 $OUT/configs_$TAG.err
