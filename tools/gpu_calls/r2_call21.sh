#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r2c21_tests.log 2>&1; tail -3 $O/r2c21_tests.log
python tools/sweep.py latency > $O/r2c21_latency.log 2>&1; cat $O/r2c21_latency.log
python bench.py > $O/r2c21_bench.json 2> $O/r2c21_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --no-e2e --no-cpu-baseline > $O/r2c21_ncu_bench.log 2>&1
tail -1 $O/r2c21_bench.err | cut -c1-300
timeout 500 python tools/n1_sweep.py PNPAPVmodel case300 15 16 > $O/r2c21_n1_case300_gpu.log 2>&1; tail -8 $O/r2c21_n1_case300_gpu.log
