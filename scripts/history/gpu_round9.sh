#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== config 4 at 1M reads (HW PATH), trace"
EDLIB_B200_TRACE=1 timeout 900 python scripts/config_runs.py --pairs3 2000 --reads4 1000000 > $OUT/configs_1M_r01.json 2> $OUT/configs_1M_r01.err; grep config4 $OUT/configs_1M_r01.json; grep -n "edlib_b200" $OUT/configs_1M_r01.err | tail -45 | head -30
