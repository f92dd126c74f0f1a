#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu (python mirror)"; timeout 600 python -m pytest tests/test_python_mirror.py -x -q -m gpu 2>&1 | tail -3
echo "== configs (steady state) with trace"
EDLIB_B200_TRACE=1 timeout 900 python scripts/config_runs.py --pairs3 20000 --reads4 100000 > $OUT/configs_r01i.json 2> $OUT/configs_r01i.err; cat $OUT/configs_r01i.json; grep -n "edlib_b200" $OUT/configs_r01i.err | sed -n '15,32p'
