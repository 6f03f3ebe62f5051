#!/bin/bash
O=gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $O/r02_launches_bench.csv python bench.py --no-e2e --no-cpu-baseline > $O/r2c44_ncu_bench.log 2>&1; tail -1 $O/r2c44_ncu_bench.log | cut -c1-120
