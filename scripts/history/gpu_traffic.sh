#!/bin/bash
# DRAM traffic of the dominant kernel at the bench size (1 M reads): two light ncu passes.
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k1_kernel -s 0 -c 1 --csv \
    --log-file $OUT/traffic_r01.csv python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline > $OUT/traffic_r01.log 2>&1
echo rc=$?; tail -4 $OUT/traffic_r01.csv
