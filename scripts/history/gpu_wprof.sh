#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
cd scripts
timeout 900 ncu --set full --clock-control none --import-source on -k regex:w_kernel -s 1 -c 1 -o ../$OUT/wk_r01 -f python w_probe.py 4000 > ../$OUT/wk_r01.log 2>&1; echo rc=$?; tail -3 ../$OUT/wk_r01.log
