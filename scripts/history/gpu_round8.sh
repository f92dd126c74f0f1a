#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== configs (steady state) with trace"
EDLIB_B200_TRACE=1 timeout 900 python scripts/config_runs.py --pairs3 20000 --reads4 100000 > $OUT/configs_r01j.json 2> $OUT/configs_r01j.err; cat $OUT/configs_r01j.json; grep -n "edlib_b200" $OUT/configs_r01j.err | sed -n '36,60p'
echo "== bench 1M quick"; timeout 600 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $OUT/q_r01j.json 2> $OUT/q_r01j.err; python -c "import json;d=json.load(open('$OUT/q_r01j.json'));print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step']),'k1ms',round(d['roofline']['kernel_ms']),'d2h',d['e2e']['d2h_bytes_per_step'],d['filter'])"
